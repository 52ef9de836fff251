"""CPU restatement of the reference's training hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import anything in
this package; nothing under ``freerl_b200/`` does, and the product path has no CPU fallback (it raises when the CUDA library is
missing).  Parity status: PINNED — the reference ships no tests or golden vectors of its own, so every module
here is checked against outputs of the UNMODIFIED reference generated in the build container by the committed
``oracle/make_golden*.py`` scripts (fixtures under ``tests/golden/``, replayed by ``tests/test_oracle_*.py`` and
``tests/test_vecloop.py``).  The one exception is documented where it occurs: ``PPO_file/PPO_with_tricks.py`` cannot execute its own
``learn()`` upstream (``np.zeros(..., dtype=torch.float32)``), so ``make_golden_ppo_tricks.py`` rebinds that single call.

Modules: ``buffers`` (ring replay, sum-tree / PER, n-step), ``algos`` (DQN, SAC, TD3, DDPG, PPO and its plain-Adam / tricks siblings),
``dqn_tricks`` / ``rainbow`` (DQN_with_tricks), ``marl`` (MADDPG / MATD3 / MAPPO / IPPO / HAPPO), ``vecloop`` (train-loop helpers),
``refload`` (imports the reference classes with gymnasium / pettingzoo stubbed; fixture generation only).
"""
