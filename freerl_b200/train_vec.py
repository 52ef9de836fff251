"""Vectorised train loops (off-policy SAC / TD3 / DQN / Rainbow, on-policy PPO / MAPPO): N gymnasium or PettingZoo-MPE envs stepped on the host cores, everything else
on the device.

    python -m freerl_b200.train_vec --algo SAC --env_name HalfCheetah-v4 --n_envs 256 --total_steps 1000000 --device cuda
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 -m freerl_b200.train_vec --algo SAC --n_envs 256 …     # 8 x 256 envs, replica sync

The same loop as the reference mains (``SAC_file/SAC.py:497-590``, ``TD3_file/TD3.py``, ``DQN_file/DQN.py:287-349``) with a leading env
axis — one ``select_action`` (batched inference kernel) per vector step, one ``add`` of N transitions, then ``n_envs *
updates_per_step`` sequential learns fused into ONE persistent launch (``learn(..., n_updates=K)``, update-to-data ratio 1 like the
reference's one learn per env step).  The per-step helpers of the reference mains run as device ops over the N envs
(``freerl_b200.vecloop``: OU / Gaussian exploration, ε-greedy, observation normalisation).  Uses the real ``gymnasium`` when it is
importable, else the synthetic shape-only shim (``freerl_b200.envshim``) — which is also what ``bench.py`` steps.

``--algo PPO`` is the on-policy loop of ``PPO_file/PPO.py:388-447``: rollouts of ``horizon`` vector steps ([horizon, N] rows in the
rollout store, GAE scanned per env column), then ``K_epochs`` of minibatches in one persistent launch.

This is the B200-native counterpart of the reference ``__main__`` blocks for users who want vectorised envs; the unchanged reference
scripts themselves run through ``freerl_b200.launcher``.
"""
import argparse
import os
import time

import numpy as np
import torch

from . import vecloop


def _gym():
    try:
        import gymnasium
        if hasattr(gymnasium, "make"):
            return gymnasium
    except Exception:
        pass
    from . import envshim
    return envshim.as_module()


def build_policy(algo, obs_dim, act_dim, n_actions, args, device):
    if algo == "SAC":
        from .SAC import SAC
        return SAC([obs_dim, act_dim], True, args.actor_lr, args.critic_lr, args.buffer_size, device,
                   trick={"ObsNorm": False, "Batch_ObsNorm": False, "OUNoise": False, "GaussNoise": False}, mode=args.mode)
    if algo == "TD3":
        from .TD3 import TD3
        return TD3([obs_dim, act_dim], True, args.actor_lr, args.critic_lr, args.buffer_size, device, trick=None,
                   realize={"clip_double": True, "policy_noise": True, "twin_delay": True}, mode=args.mode)
    if algo == "DQN":
        from .DQN import DQN
        return DQN([obs_dim, n_actions], False, args.actor_lr, args.buffer_size, device, mode=args.mode)
    if algo == "RAINBOW":     # DQN_with_tricks.py with every trick on (PER + Noisy + C51 + N-step + Double + Dueling), BASELINE config C4
        from .DQN_with_tricks import DQN
        trick = {"Double": True, "Dueling": True, "PER": True, "Noisy": True, "N_Step": True, "Categorical": True}
        return DQN([obs_dim, n_actions], False, args.actor_lr, args.buffer_size, device, trick=trick, gamma=args.gamma,
                   batch_size=args.batch_size, mode=args.mode)
    raise ValueError("algo must be SAC, TD3, DQN or RAINBOW")


def _save_obs_norm(args, norm):
    """The reference mains store the running observation statistics next to the model (``SAC.py:583-584``,
    ``PPO_with_tricks.py:581-582``: ``np.array([mean, std])``) — evaluate.py and a resumed run need them."""
    if norm is not None:
        ms = norm.running_ms
        np.save(os.path.join(args.save_dir, "%s_running_mean_std.npy" % args.algo), np.array([ms.mean.cpu().numpy(), ms.std.cpu().numpy()]))


def _ppo_loop(args, envs, obs_dim, action_dim, discrete, device, world=1, rank=0):
    from .PPO import PPO
    N, T = args.n_envs, args.horizon
    max_action = None if discrete else float(envs[0].action_space.high[0])
    policy = PPO([obs_dim, action_dim], not discrete, args.actor_lr, args.critic_lr, T * N, device, trick={"adv_norm": False}, mode=args.mode)
    if world > 1:      # synchronous data parallel: every minibatch step all-reduces ONE flat gradient buffer, replicas stay bit-identical
        policy.enable_data_parallel()
    norm = vecloop.Normalization(obs_dim, device) if args.obs_norm else None
    observe = (lambda rows: norm(np.stack(rows).astype(np.float32)).cpu().numpy()) if norm is not None else (lambda rows: np.stack(rows).astype(np.float32))
    obs = observe([e.reset(seed=args.seed + i)[0] for i, e in enumerate(envs)])
    ep_ret, returns, steps, n_learn, t0 = np.zeros(N), [], 0, 0, time.perf_counter()
    while steps < args.total_steps:
        for _ in range(T):
            action, logp = policy.select_action(obs)                                     # [N] ids or [N, act] in (-1, 1)
            out = [e.step(int(a) if discrete else np.clip(a * max_action, -max_action, max_action)) for e, a in zip(envs, action)]
            next_raw = [o[0] for o in out]
            reward = np.array([o[1] for o in out], dtype=np.float64)
            terminated = np.array([o[2] for o in out], dtype=bool)
            done = terminated | np.array([o[3] for o in out], dtype=bool)
            ep_ret += reward
            for i in np.nonzero(done)[0]:
                returns.append(ep_ret[i]); ep_ret[i] = 0.0
            next_obs = observe(next_raw)
            policy.add(obs, action, reward, next_obs, terminated, logp, done)
            obs = next_obs
            if done.any():
                idx = np.nonzero(done)[0]
                obs = next_obs.copy()
                obs[idx] = observe([envs[i].reset(seed=args.seed + int(i))[0] for i in idx])
            steps += N
        policy.learn(min(args.minibatch_size, T * N), args.gamma, args.lmbda, args.clip_param, args.K_epochs, args.entropy_coefficient)
        n_learn += 1
        if args.log_every:
            print("steps %d  rollouts %d  %.0f env-steps/s  mean return(last 20) %s" % (
                steps, n_learn, steps / (time.perf_counter() - t0), "%.2f" % np.mean(returns[-20:]) if returns else "n/a"), flush=True)
    if args.save_dir and rank == 0:
        os.makedirs(args.save_dir, exist_ok=True)
        policy.save(args.save_dir)
        _save_obs_norm(args, norm)
    return {"policy": policy, "steps": steps, "learns": n_learn, "returns": returns, "rank": rank, "world": world}


def _mpe(name, **kw):
    try:
        import importlib
        mod = importlib.import_module("pettingzoo.mpe." + name)
        if hasattr(mod, "parallel_env") and not getattr(importlib.import_module("pettingzoo"), "__freerl_b200_shim__", False):
            return mod.parallel_env(**kw)
    except Exception:
        pass
    from . import envshim
    return envshim.mpe_modules()["pettingzoo.mpe." + name].parallel_env(**kw)


def _mappo_loop(args, device, world=1, rank=0):
    """``MAPPO_file/MAPPO.py:640-742`` over N parallel envs (BASELINE config C5: simple_spread_v3, 3 agents, continuous actions)."""
    from .MAPPO import MAPPO
    N, T = args.n_envs, args.horizon
    envs = [_mpe(args.env_name, max_cycles=25, continuous_actions=True, **({"N": args.n_agents} if args.n_agents else {})) for _ in range(N)]
    first = [e.reset(seed=args.seed + i)[0] for i, e in enumerate(envs)]
    ids = list(envs[0].agents)
    dim_info = {a: [envs[0].observation_space(a).shape[0], envs[0].action_space(a).shape[0]] for a in ids}
    trick = {'adv_norm': True, 'ObsNorm': False, 'reward_norm': False, 'reward_scaling': False, 'orthogonal_init': True, 'adam_eps': True,
             'lr_decay': False, 'ValueClip': False, 'huber_loss': False, 'LayerNorm': True, 'feature_norm': True}
    policy = MAPPO(dim_info, True, args.actor_lr, args.critic_lr, T * N, device, trick, mode=args.mode)
    if world > 1:      # synchronous data parallel (BASELINE config 5 at 8 GPUs): every rank steps its own N envs; per optimiser step the flat
        policy.enable_data_parallel()       # gradient of the agent is summed over the ranks, advantages are normalised over the union rollout
    stack = lambda dicts: {a: np.stack([d[a] for d in dicts]).astype(np.float32) for a in ids}
    obs = stack(first)
    ep_ret, returns, steps, n_learn, t0 = np.zeros(N), [], 0, 0, time.perf_counter()
    while steps < args.total_steps:
        for _ in range(T):
            action, logp = policy.select_action(obs)                                     # per agent [N, act] in (-1, 1)
            act_env = {a: (np.clip(action[a], -1.0, 1.0).astype(np.float32) + 1) / 2 for a in ids}          # MAPPO.py:682-683: -> [0, 1]
            out = [e.step({a: act_env[a][i] for a in ids}) for i, e in enumerate(envs)]
            next_obs = stack([o[0] for o in out])
            reward = {a: np.array([o[1][a] for o in out], dtype=np.float64) for a in ids}
            term = {a: np.array([o[2][a] for o in out], dtype=bool) for a in ids}
            done = {a: term[a] | np.array([o[3][a] for o in out], dtype=bool) for a in ids}
            policy.add(obs, action, reward, next_obs, term, logp, done)
            ep_ret += sum(reward[a] for a in ids)
            obs = next_obs
            over = np.array([not e.agents for e in envs]) | np.any([done[a] for a in ids], axis=0)
            if over.any():
                obs = {a: next_obs[a].copy() for a in ids}
                for i in np.nonzero(over)[0]:
                    returns.append(ep_ret[i]); ep_ret[i] = 0.0
                    o0 = envs[i].reset(seed=args.seed + int(i))[0]
                    for a in ids:
                        obs[a][i] = o0[a]
            steps += N
        policy.learn(min(args.minibatch_size, T * N), args.gamma, args.lmbda, args.clip_param, args.K_epochs, args.entropy_coefficient)
        n_learn += 1
        if args.log_every:
            print("steps %d  rollouts %d  %.0f env-steps/s  mean return(last 20) %s" % (
                steps, n_learn, steps / (time.perf_counter() - t0), "%.2f" % np.mean(returns[-20:]) if returns else "n/a"), flush=True)
    if args.save_dir and rank == 0:
        os.makedirs(args.save_dir, exist_ok=True)
        policy.save(args.save_dir)
    return {"policy": policy, "steps": steps, "learns": n_learn, "returns": returns, "rank": rank, "world": world}


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--algo", default="SAC", choices=["SAC", "TD3", "DQN", "RAINBOW", "PPO", "MAPPO"])
    ap.add_argument("--env_name", default="HalfCheetah-v4")
    ap.add_argument("--n_envs", type=int, default=256)
    ap.add_argument("--total_steps", type=int, default=100_000, help="env steps summed over the envs")
    ap.add_argument("--random_steps", type=int, default=5_000)
    ap.add_argument("--start_steps", type=int, default=5_000, help="env steps collected before the first learn")
    ap.add_argument("--updates_per_step", type=float, default=1.0, help="learns per env step (the reference: 1)")
    ap.add_argument("--batch_size", type=int, default=256)
    ap.add_argument("--buffer_size", type=int, default=1_000_000)
    ap.add_argument("--gamma", type=float, default=0.99)
    ap.add_argument("--tau", type=float, default=0.01)
    ap.add_argument("--actor_lr", type=float, default=1e-3)
    ap.add_argument("--critic_lr", type=float, default=1e-3)
    ap.add_argument("--epsilon", type=float, default=0.1, help="DQN epsilon-greedy")
    ap.add_argument("--gauss_sigma", type=float, default=0.1, help="TD3 exploration noise (DDPG.py:522 form)")
    ap.add_argument("--policy_noise", type=float, default=0.1, help="TD3 target smoothing (TD3.py:343-345)")
    ap.add_argument("--noise_clip", type=float, default=0.5)
    ap.add_argument("--policy_freq", type=int, default=2)
    ap.add_argument("--horizon", type=int, default=128, help="PPO: vector steps per rollout (rollout rows = horizon * n_envs)")
    ap.add_argument("--minibatch_size", type=int, default=8192)
    ap.add_argument("--K_epochs", type=int, default=10)
    ap.add_argument("--lmbda", type=float, default=0.95)
    ap.add_argument("--clip_param", type=float, default=0.2)
    ap.add_argument("--entropy_coefficient", type=float, default=0.01)
    ap.add_argument("--n_agents", type=int, default=0, help="MAPPO: N of the MPE env (0: its default)")
    ap.add_argument("--obs_norm", action="store_true", help="running observation normalisation over all envs (vecloop.Normalization)")
    ap.add_argument("--mode", default="fast", choices=["fast", "parity"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--save_dir", default=None)
    ap.add_argument("--log_every", type=int, default=50, help="vector steps between progress lines (0: quiet)")
    args = ap.parse_args(argv)

    # one process per GPU (torchrun): every rank steps its own n_envs envs into its own replay shard (PER: its own sum-tree) — no
    # data-path collective — and the SAC / TD3 / DQN / Rainbow replicas are kept one policy by a parameter average per vector step
    # (sync_replicas); PPO / MAPPO train synchronously data parallel (gradient sum per optimiser step)
    world, rank, dist = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), None
    if world > 1:
        if args.obs_norm:
            # per-rank running statistics would feed differently normalised observations to replicas that are then averaged /
            # gradient-summed as one policy; a merged (Chan) statistic per vector step is not implemented
            raise ValueError("--obs_norm is single-process only (the running statistics are per process)")
        import torch.distributed as dist
        if args.device.startswith("cuda"):
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
            args.device = "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0"))
        if not dist.is_initialized():
            dist.init_process_group("nccl" if args.device.startswith("cuda") else "gloo")
        args.seed += 1000 * rank
        args.log_every = args.log_every if rank == 0 else 0
    device = torch.device(args.device)
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    if args.algo == "MAPPO":
        return _mappo_loop(args, device, world, rank)
    gym = _gym()
    N = args.n_envs
    envs = [gym.make(args.env_name) for _ in range(N)]
    space = envs[0].action_space
    obs_dim = envs[0].observation_space.shape[0]
    discrete = not hasattr(space, "high")
    if args.algo != "PPO" and discrete != (args.algo in ("DQN", "RAINBOW")):
        raise ValueError("%s needs a %s action space (%s has the other kind)" % (args.algo, "discrete" if args.algo in ("DQN", "RAINBOW") else "continuous", args.env_name))
    act_dim = 1 if discrete else space.shape[0]
    n_actions = space.n if discrete else 0
    if args.algo == "PPO":
        return _ppo_loop(args, envs, obs_dim, space.n if discrete else space.shape[0], discrete, device, world, rank)
    max_action = None if discrete else float(space.high[0])
    for i, e in enumerate(envs):
        e.action_space.seed(seed=args.seed + i)
    policy = build_policy(args.algo, obs_dim, act_dim, n_actions, args, device)
    if world > 1:
        policy.enable_replica_sync()                                                     # start every replica from rank 0's parameters
    norm = vecloop.Normalization(obs_dim, device) if args.obs_norm else None

    def observe(rows):
        rows = np.stack(rows).astype(np.float32)
        return norm(rows).cpu().numpy() if norm is not None else rows

    obs = observe([e.reset(seed=args.seed + i)[0] for i, e in enumerate(envs)])
    ep_ret, returns = np.zeros(N), []
    steps, vec_step, n_learn, t0 = 0, 0, 0, time.perf_counter()
    carry = 0.0
    while steps < args.total_steps:
        if steps < args.random_steps:
            if discrete:
                action = np.array([e.action_space.sample() for e in envs], dtype=np.int64)
            else:
                action = np.stack([e.action_space.sample() for e in envs]) / max_action
        else:
            action = policy.select_action(obs)                                           # one batched inference launch for the N envs
            if args.algo == "DQN":                                                       # Rainbow explores through its NoisyLinear layers
                action = vecloop.epsilon_greedy(action, n_actions, args.epsilon, device=device, mode=args.mode, seed=args.seed,
                                                counter=vec_step).cpu().numpy()
        if discrete:
            action_ = action
        elif args.algo == "TD3" and steps >= args.random_steps:
            action_ = vecloop.explore_gauss(action, max_action, 1.0, args.gauss_sigma, device=device, mode=args.mode, seed=args.seed,
                                            counter=vec_step).cpu().numpy()
        else:
            action_ = np.clip(action * max_action, -max_action, max_action)
        out = [e.step(a if not discrete else int(a)) for e, a in zip(envs, action_)]
        next_raw = [o[0] for o in out]
        reward = np.array([o[1] for o in out], dtype=np.float64)
        terminated = np.array([o[2] for o in out], dtype=bool)
        done = terminated | np.array([o[3] for o in out], dtype=bool)
        ep_ret += reward
        for i in np.nonzero(done)[0]:                                                    # the stored next_obs is the terminal one; reset after
            returns.append(ep_ret[i]); ep_ret[i] = 0.0
        next_obs = observe(next_raw)
        policy.add(obs, np.asarray(action).reshape(N, act_dim), reward, next_obs, terminated)
        obs = next_obs
        if done.any():                                                                   # only the reset envs' first observations are new rows
            idx = np.nonzero(done)[0]
            obs = next_obs.copy()
            obs[idx] = observe([envs[i].reset(seed=args.seed + int(i))[0] for i in idx])
        steps += N
        vec_step += 1
        if steps >= args.start_steps:
            carry += N * args.updates_per_step
            k = int(carry)
            if k > 0:
                if args.algo == "TD3":                                                   # k sequential learns, one persistent launch
                    policy.learn(args.batch_size, args.gamma, args.tau, args.policy_noise, args.noise_clip, max_action, args.policy_freq,
                                 1.0, n_updates=k)
                elif args.algo == "RAINBOW":                                             # PER priorities feed back between learns: one launch each
                    for _ in range(k):
                        policy.learn(args.batch_size, args.gamma, args.tau)
                else:
                    policy.learn(args.batch_size, args.gamma, args.tau, n_updates=k)
                carry -= k
                n_learn += k
                if world > 1:
                    policy.sync_replicas()
        if args.log_every and vec_step % args.log_every == 0:
            dt = time.perf_counter() - t0
            print("steps %d  learns %d  %.0f env-steps/s  mean return(last 20) %s" % (
                steps, n_learn, steps / dt, "%.2f" % np.mean(returns[-20:]) if returns else "n/a"), flush=True)
    if args.save_dir and rank == 0:
        os.makedirs(args.save_dir, exist_ok=True)
        policy.save(args.save_dir)
        _save_obs_norm(args, norm)
        np.save(os.path.join(args.save_dir, "%s_seed_%d.npy" % (args.algo, args.seed)), np.array(returns))
    return {"policy": policy, "steps": steps, "learns": n_learn, "returns": returns, "rank": rank, "world": world}


if __name__ == "__main__":
    main()
