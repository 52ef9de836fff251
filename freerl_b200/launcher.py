"""Run an UNCHANGED reference train script on top of freerl_b200 (SURVEY §8b "binding mechanism").

    python -m freerl_b200.launcher /path/to/FreeRL/SAC_file/SAC.py --env_name HalfCheetah-v4 --device cuda ...

What it does, in order: (1) pre-seeds ``sys.modules['Buffer']`` with our device buffers (``sys.modules`` beats the
script directory on ``sys.path``); (2) if ``gymnasium`` is not importable, injects the synthetic shape-only shim
(``freerl_b200.envshim``); (3) executes the script's top-level statements (imports, class definitions, ``get_env`` …)
unmodified; (4) rebinds the algorithm class name in the script namespace to ours; (5) executes the script's
``if __name__ == '__main__':`` block.  ``__file__`` points into ``--results-root`` (default ``./freerl_runs``) so
``make_dir`` writes ``results/`` there instead of next to a possibly read-only script.
"""
import ast
import importlib
import os
import sys
import types

_ALGOS = {            # script file name -> (class name in the script, our module, our class)
    "DQN.py": ("DQN", "freerl_b200.DQN", "DQN"),
    "DQN_with_tricks.py": ("DQN", "freerl_b200.DQN_with_tricks", "DQN"),
    "SAC.py": ("SAC", "freerl_b200.SAC", "SAC"),
    "SAC_add_discrete.py": ("SAC", "freerl_b200.SAC_add_discrete", "SAC"),
    "TD3.py": ("TD3", "freerl_b200.TD3", "TD3"),
    "DDPG.py": ("DDPG", "freerl_b200.DDPG", "DDPG"),
    "PPO.py": ("PPO", "freerl_b200.PPO", "PPO"),
    "PPO_advance/PPO.py": ("PPO", "freerl_b200.PPO_advance", "PPO"),      # keyed by <dir>/<file> where names collide
    "PPO_cc.py": ("PPO", "freerl_b200.PPO_advance", "PPO"),               # same classes as PPO_advance/PPO.py (only the train loop differs)
    "PPO_with_tricks.py": ("PPO", "freerl_b200.PPO_with_tricks", "PPO"),  # PPO_file/ and PPO_advance/ hold the same classes
    "MADDPG.py": ("MADDPG", "freerl_b200.MADDPG", "MADDPG"),
    "MADDPG_simple.py": ("MADDPG", "freerl_b200.MADDPG_simple", "MADDPG"),
    "MATD3_simple.py": ("MATD3", "freerl_b200.MATD3_simple", "MATD3"),
    "DDPG_simple.py": ("DDPG", "freerl_b200.DDPG_simple", "DDPG"),
    "MAPPO.py": ("MAPPO", "freerl_b200.MAPPO", "MAPPO"),
    "MAPPO_discrete.py": ("MAPPO", "freerl_b200.MAPPO_discrete", "MAPPO"),  # shared nets + episode ReplayBuffer (--policy_name MAPPO_simple)
    "IPPO.py": ("IPPO", "freerl_b200.IPPO", "IPPO"),
    "HAPPO.py": ("HAPPO", "freerl_b200.HAPPO", "HAPPO"),
}


def buffer_module():
    from . import Buffer as B, MAPPO_discrete, per
    m = types.ModuleType("Buffer")
    m.ReplayBuffer = MAPPO_discrete.ReplayBuffer
    m.Buffer, m.Buffer_for_PPO = B.Buffer, B.Buffer_for_PPO
    m.SumTree, m.PER_Buffer = per.SumTree, per.PER_Buffer
    m.N_Step_Buffer, m.N_Step_PER_Buffer = per.N_Step_Buffer, per.N_Step_PER_Buffer
    return m


def _is_main_guard(node):
    if not isinstance(node, ast.If) or not isinstance(node.test, ast.Compare):
        return False
    t = node.test
    return (isinstance(t.left, ast.Name) and t.left.id == "__name__" and len(t.comparators) == 1
            and isinstance(t.comparators[0], ast.Constant) and t.comparators[0].value == "__main__")


def run_reference_script(script_path, argv=(), results_root="./freerl_runs", extra_rebinds=None):
    script_path = os.path.abspath(script_path)
    fname = os.path.basename(script_path)
    qual = os.path.basename(os.path.dirname(script_path)) + "/" + fname
    if qual not in _ALGOS and fname not in _ALGOS:
        raise ValueError("no freerl_b200 class for %s (supported: %s)" % (fname, sorted(_ALGOS)))
    cls_name, mod_name, our_name = _ALGOS.get(qual) or _ALGOS[fname]
    sys.modules["Buffer"] = buffer_module()
    def usable(name, attr):      # importable AND a real package (an empty placeholder left in sys.modules by other tooling is not)
        try:
            return hasattr(importlib.import_module(name), attr)
        except Exception:
            return False
    if not usable("gymnasium", "make"):
        from . import envshim
        sys.modules["gymnasium"] = envshim.as_module()
    if not usable("pettingzoo.mpe", "__path__"):
        from . import envshim
        sys.modules.update(envshim.mpe_modules())
    for helper in ("Noisy_net", "normalization", "c_adamw", "util"):      # same module names, different contents per directory
        sys.modules.pop(helper, None)
    sdir = os.path.dirname(script_path)
    if sdir not in sys.path:
        sys.path.insert(0, sdir)                 # sibling helpers (Noisy_net, c_adamw, normalization) stay the reference's own
    run_dir = os.path.abspath(os.path.join(results_root, os.path.basename(sdir)))
    os.makedirs(run_dir, exist_ok=True)
    with open(script_path, "r", encoding="utf-8") as f:
        tree = ast.parse(f.read(), filename=script_path)
    head = [n for n in tree.body if not _is_main_guard(n)]
    mains = [n for n in tree.body if _is_main_guard(n)]
    ns = {"__name__": "freerl_b200_reference_script", "__file__": os.path.join(run_dir, fname), "__builtins__": __builtins__}
    exec(compile(ast.Module(body=head, type_ignores=[]), script_path, "exec"), ns)
    ns[cls_name] = getattr(importlib.import_module(mod_name), our_name)
    for k, v in (extra_rebinds or {}).items():
        ns[k] = v
    old_argv = sys.argv
    sys.argv = [script_path] + list(argv)
    try:
        for guard in mains:
            exec(compile(ast.Module(body=guard.body, type_ignores=[]), script_path, "exec"), ns)
    finally:
        sys.argv = old_argv
    return ns


def main():
    if len(sys.argv) < 2:
        print(__doc__)
        raise SystemExit(2)
    root = "./freerl_runs"
    args = sys.argv[2:]
    if "--results-root" in args:
        i = args.index("--results-root")
        root = args[i + 1]
        del args[i:i + 2]
    run_reference_script(sys.argv[1], args, results_root=root)


if __name__ == "__main__":
    main()
