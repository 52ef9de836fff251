"""Resume checkpoints: the reference only saves actor / Q ``state_dict``s (``DQN.py:131-138``, ``SAC.py:274-282`` …), so a
run cannot be continued.  ``save_checkpoint(policy, path)`` stores EVERYTHING the next ``learn()`` depends on — parameter blocks
with their Adam moments and step counters, target networks, the device replay / rollout stores with ring positions, the PER
sum-tree and beta, n-step windows, SAC's temperature state, Batch_ObsNorm statistics, sampler counters and the numpy / torch RNG
states — and ``load_checkpoint(policy, path)`` restores it IN PLACE into a policy constructed with the same arguments (device
tensors are ``copy_``-ed, so the raw pointers cached in the C descriptors stay valid).  A restored run continues bit-identically.

The state is found by walking the policy's object graph (every ``freerl_b200`` object, dict / list / tuple / deque containers);
``nn.Module`` shims are skipped (their parameters alias the device blocks), as are transient members (``last_*``, keep-alive refs).
"""
import collections

import numpy as np
import torch

# transient members, by EXACT name (a prefix rule would silently drop any future `_c…` / `_keep…` state from checkpoints):
# results of the last call, keep-alive references of in-flight launches, cached C descriptors, process-group handles, and the
# scratch of the small-batch actor-critic schedule (rewritten by every launch)
_SKIP = frozenset(("last_adv", "last_error", "last_factor", "last_indices", "last_metrics", "last_path", "last_v_target", "_last_eps",
                   "_keep", "_keep_noise", "_keepalive", "_c", "_dp", "_dp_peers", "_rs", "_rs_peers", "_fx_ws", "_fx_sync",
                   "_host",         # pinned staging tensors of MAPPO_discrete.ReplayBuffer: `buffer` (numpy views of them) is the state
                   "_stage_buf", "_stage_idx"))     # per-learn staging replay of the discrete SAC's Batch_ObsNorm (refilled by every learn)
_SCALARS = (int, float, bool, str, type(None), np.integer, np.floating, np.bool_)


def _walk(obj, path, out, seen, frozen=False):
    if id(obj) in seen:
        return
    if isinstance(obj, torch.Tensor):
        out[path] = obj
        return
    if isinstance(obj, _SCALARS) or isinstance(obj, np.ndarray):
        if not frozen:                       # members of a tuple are configuration (immutable), not state
            out[path] = obj
        return
    if isinstance(obj, torch.nn.Module) or callable(obj) and not hasattr(obj, "__dict__"):
        return
    if isinstance(obj, collections.deque):
        out[path] = obj                      # n-step windows: numpy payloads, stored whole (aliases share one object, see save)
        return
    if isinstance(obj, dict):
        seen.add(id(obj))
        for k, v in obj.items():
            if isinstance(k, (str, int)):
                _walk(v, "%s[%r]" % (path, k), out, seen)
        return
    if isinstance(obj, (list, tuple)):
        seen.add(id(obj))
        for i, v in enumerate(obj):
            _walk(v, "%s[%d]" % (path, i), out, seen, frozen=isinstance(obj, tuple))
        return
    mod = type(obj).__module__ or ""
    if not mod.startswith("freerl_b200") or not hasattr(obj, "__dict__"):
        return
    seen.add(id(obj))
    for k, v in vars(obj).items():
        if k in _SKIP:
            continue
        _walk(v, "%s.%s" % (path, k), out, seen)


def state_of(policy):
    out = {}
    _walk(policy, "policy", out, set())
    return out


def save_checkpoint(policy, path):
    st = state_of(policy)
    blob = {"format": "freerl_b200.checkpoint/1", "class": "%s.%s" % (type(policy).__module__, type(policy).__name__), "state": {}}
    deque_ids = {}
    for k, v in st.items():
        if isinstance(v, torch.Tensor):
            blob["state"][k] = ("tensor", v.detach().cpu().clone())
        elif isinstance(v, collections.deque):
            first = deque_ids.setdefault(id(v), k)
            blob["state"][k] = ("deque", (list(v), v.maxlen)) if first == k else ("alias", first)
        else:
            blob["state"][k] = ("value", v)
    blob["rng"] = {"numpy": np.random.get_state(), "torch": torch.get_rng_state(),
                   "cuda": torch.cuda.get_rng_state_all() if torch.cuda.is_available() else None}
    torch.save(blob, path)


def _assign(policy, path, value):
    """set `policy<path> = value` for a path made of .attr and [key] steps (keys are str / int literals)"""
    import re
    from ast import literal_eval as _lit
    steps = re.findall(r"\.([A-Za-z_]\w*)|\[([^\]]+)\]", path[len("policy"):])
    obj = policy
    for attr, key in steps[:-1]:
        obj = getattr(obj, attr) if attr else obj[_lit(key)]
    attr, key = steps[-1]
    if attr:
        setattr(obj, attr, value)
    elif isinstance(obj, list) and _lit(key) == len(obj):
        obj.append(value)                    # e.g. one n-step window per vectorised env, created lazily
    else:
        obj[_lit(key)] = value


def load_checkpoint(policy, path, restore_rng=True):
    """Checkpoints are TRUSTED files (pickled numpy payloads and RNG states): only load what this package wrote."""
    blob = torch.load(path, weights_only=False)
    if blob.get("format") != "freerl_b200.checkpoint/1":
        raise ValueError("not a freerl_b200 checkpoint: %s" % path)
    want = "%s.%s" % (type(policy).__module__, type(policy).__name__)
    if blob["class"] != want:
        raise ValueError("checkpoint of %s cannot be loaded into %s" % (blob["class"], want))
    cur = state_of(policy)
    missing = [k for k, (kind, _) in blob["state"].items() if kind == "tensor" and k not in cur]
    if missing:
        raise ValueError("policy was constructed differently from the checkpointed one (no %s)" % missing[0])
    restored = {}
    for k, (kind, v) in blob["state"].items():
        if kind == "tensor":
            dst = cur[k]
            if not isinstance(dst, torch.Tensor):        # lazily allocated scratch (None in a fresh policy)
                _assign(policy, k, v.to(policy.device))
                continue
            if tuple(dst.shape) != tuple(v.shape):
                raise ValueError("shape mismatch at %s: %s vs %s" % (k, tuple(dst.shape), tuple(v.shape)))
            dst.copy_(v.to(dst.device))
        elif kind == "deque":
            restored[k] = collections.deque(v[0], maxlen=v[1])
            _assign(policy, k, restored[k])
        elif kind == "alias":                # a second name of a deque saved above: the same object again
            _assign(policy, k, restored[v])
        else:
            _assign(policy, k, v)
    if restore_rng:
        np.random.set_state(blob["rng"]["numpy"])
        torch.set_rng_state(blob["rng"]["torch"])
        if blob["rng"]["cuda"] is not None and torch.cuda.is_available():
            torch.cuda.set_rng_state_all(blob["rng"]["cuda"])
    return policy
