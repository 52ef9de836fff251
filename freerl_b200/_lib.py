"""ctypes binding of ``libfreerl_b200.so`` (the C ABI declared in ``include/freerl_b200.h``).

This is the reference-side binding a maintainer adds: plain ``ctypes``, raw device pointers
(``tensor.data_ptr()``) and the current CUDA stream.  There is NO fallback: if the shared library is missing
or reports a different ABI the import of any compute class raises.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
FRL_MAX_LAYERS = 6
FRL_MAX_AGENTS = 6
ABI_VERSION = 10


class Layer(C.Structure):
    _fields_ = [("in_", C.c_int), ("out", C.c_int), ("in_pad", C.c_int), ("out_pad", C.c_int),
                ("w_off", C.c_int), ("b_off", C.c_int), ("wt_off", C.c_int)]


class Net(C.Structure):
    _fields_ = [("p", C.c_void_p), ("pt", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("g", C.c_void_p),
                ("n_p", C.c_int), ("n_pt", C.c_int), ("n_layers", C.c_int), ("x_off", C.c_int), ("x_len", C.c_int),
                ("L", Layer * FRL_MAX_LAYERS)]


class Replay(C.Structure):
    _fields_ = [("storage", C.c_void_p), ("capacity", C.c_int64), ("row_floats", C.c_int), ("obs_dim", C.c_int),
                ("act_dim", C.c_int)]


class DqnArgs(C.Structure):
    _fields_ = [("q", Net), ("q_target", Net), ("replay", Replay), ("indices", C.c_void_p), ("B", C.c_int),
                ("n_updates", C.c_int), ("gamma", C.c_float), ("tau", C.c_float), ("lr", C.c_double),
                ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double), ("step0", C.c_int64),
                ("gpart", C.c_void_p), ("stats", C.c_void_p), ("out", C.c_void_p),
                ("double_q", C.c_int), ("dueling", C.c_int), ("is_weight", C.c_void_p), ("td_error", C.c_void_p)]


class AcArgs(C.Structure):
    _fields_ = [("actor", Net), ("actor_target", Net), ("critic", Net), ("critic_target", Net),
                ("n_heads", C.c_int), ("actor_kind", C.c_int), ("replay", Replay), ("indices", C.c_void_p),
                ("B", C.c_int), ("n_updates", C.c_int), ("noise_next", C.c_void_p), ("noise_new", C.c_void_p),
                ("seed", C.c_uint64), ("gamma", C.c_float), ("tau", C.c_float),
                ("lr_actor", C.c_double), ("lr_critic", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double),
                ("eps", C.c_double), ("wd_critic", C.c_double), ("max_norm", C.c_float),
                ("step_actor0", C.c_int64), ("step_critic0", C.c_int64), ("total_it0", C.c_int64),
                ("policy_freq", C.c_int), ("target_smoothing", C.c_int), ("policy_noise", C.c_float),
                ("noise_clip", C.c_float), ("max_action", C.c_float), ("policy_noise_scale", C.c_float),
                ("alpha_state", C.c_void_p), ("adaptive_alpha", C.c_int), ("alpha_lr", C.c_double),
                ("target_entropy", C.c_float), ("step_alpha0", C.c_int64),
                ("gpart", C.c_void_p), ("sumsq", C.c_void_p), ("stats", C.c_void_p), ("out", C.c_void_p),
                ("n_agents", C.c_int), ("agent_index", C.c_int), ("ma_replay", Replay * FRL_MAX_AGENTS),
                ("ma_actor_target", Net * FRL_MAX_AGENTS), ("defer_polyak", C.c_int), ("xchg", C.c_void_p),
                ("obs_norm", C.c_void_p * FRL_MAX_AGENTS), ("obs_norm_n0", C.c_int64),
                ("ma_noise_next", C.c_void_p * FRL_MAX_AGENTS), ("ws", C.c_void_p), ("sync", C.c_void_p)]


class InferArgs(C.Structure):
    _fields_ = [("net", Net), ("l0", C.c_int), ("nl", C.c_int), ("obs", C.c_void_p), ("n", C.c_int), ("obs_dim", C.c_int), ("mode", C.c_int),
                ("noise", C.c_void_p), ("seed", C.c_uint64), ("counter", C.c_uint32), ("out", C.c_void_p),
                ("out_cols", C.c_int), ("layer_norm", C.c_int), ("obs_norm", C.c_void_p), ("hidden_tanh", C.c_int)]


FRL_DP_MAX_RANKS = 8


class DpPeers(C.Structure):
    _fields_ = [("g", C.c_void_p * FRL_DP_MAX_RANKS), ("flags", C.c_void_p * FRL_DP_MAX_RANKS), ("rank", C.c_int), ("world", C.c_int),
                ("epoch0", C.c_uint)]


class PpoArgs(C.Structure):
    _fields_ = [("net", Net), ("continuous", C.c_int), ("obs", C.c_void_p), ("action", C.c_void_p),
                ("logp_old", C.c_void_p), ("adv", C.c_void_p), ("v_target", C.c_void_p),
                ("M", C.c_int), ("obs_dim", C.c_int), ("act_cols", C.c_int), ("logp_cols", C.c_int), ("n_adv", C.c_int),
                ("indices", C.c_void_p), ("mb_rows", C.c_void_p), ("mb", C.c_int), ("n_updates", C.c_int),
                ("clip_param", C.c_float), ("entropy_coef", C.c_float), ("max_norm_actor", C.c_float),
                ("max_norm_critic", C.c_float), ("optimizer", C.c_int), ("lr", C.c_double), ("beta1", C.c_double),
                ("beta2", C.c_double), ("eps", C.c_double), ("step0", C.c_int64),
                ("layer_norm", C.c_int), ("critic_obs", C.c_void_p), ("critic_obs_dim", C.c_int), ("value_loss", C.c_int),
                ("huber_delta", C.c_float), ("stage_lo", C.c_int), ("stage_hi", C.c_int), ("grad_scale", C.c_float),
                ("gpart", C.c_void_p),
                ("sumsq", C.c_void_p), ("segcnt", C.c_void_p), ("stats", C.c_void_p), ("out", C.c_void_p),
                ("lr_critic", C.c_double), ("hidden_tanh", C.c_int), ("umma_ws", C.c_void_p), ("dp", DpPeers),
                ("max_norm_joint", C.c_float), ("opt_repeat", C.c_int), ("v_old", C.c_void_p),
                ("group_rows", C.c_int), ("group_norm", C.c_int), ("group_prepass", C.c_int)]


FRL_RA_MAX_TENSORS = 8


class ReplicaAvgArgs(C.Structure):
    _fields_ = [("dp", DpPeers), ("tensor", C.c_void_p * FRL_RA_MAX_TENSORS), ("n", C.c_int * FRL_RA_MAX_TENSORS), ("n_tensors", C.c_int),
                ("block_floats", C.c_longlong), ("status", C.c_void_p)]


class SacdArgs(C.Structure):
    _fields_ = [("actor", Net), ("actor_target", Net), ("critic", Net), ("critic_target", Net), ("replay", Replay),
                ("indices", C.c_void_p), ("B", C.c_int), ("n_updates", C.c_int), ("gamma", C.c_float), ("tau", C.c_float),
                ("lr_actor", C.c_double), ("lr_critic", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
                ("max_norm", C.c_float), ("step_actor0", C.c_int64), ("step_critic0", C.c_int64), ("alpha_state", C.c_void_p),
                ("adaptive_alpha", C.c_int), ("alpha_lr", C.c_double), ("target_entropy", C.c_float), ("step_alpha0", C.c_int64),
                ("gpart", C.c_void_p), ("sumsq", C.c_void_p), ("stats", C.c_void_p), ("out", C.c_void_p)]


class NoisyMap(C.Structure):
    _fields_ = [("mu_w", C.c_int), ("sg_w", C.c_int), ("mu_b", C.c_int), ("sg_b", C.c_int), ("row0", C.c_int),
                ("eps_in", C.c_int), ("eps_out", C.c_int)]


class ExploreArgs(C.Structure):
    _fields_ = [("kind", C.c_int), ("N", C.c_int), ("A", C.c_int), ("action", C.c_void_p), ("ou_state", C.c_void_p), ("z", C.c_void_p),
                ("seed", C.c_uint64), ("counter", C.c_uint64), ("mu", C.c_double), ("theta", C.c_double), ("sigma", C.c_double),
                ("dt", C.c_double), ("scale", C.c_double), ("gauss_scale", C.c_double), ("gauss_sigma", C.c_double),
                ("max_action", C.c_double), ("clip", C.c_int), ("out64", C.c_void_p), ("out", C.c_void_p)]


class RainbowArgs(C.Structure):
    _fields_ = [("p", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("p_target", C.c_void_p), ("n_train", C.c_int),
                ("eff", Net * 3), ("map", NoisyMap * FRL_MAX_LAYERS), ("eps", C.c_void_p), ("eps_len", C.c_int),
                ("n_actions", C.c_int), ("n_atoms", C.c_int), ("z", C.c_void_p), ("v_min", C.c_float), ("v_max", C.c_float),
                ("delta_z", C.c_float), ("double_q", C.c_int), ("replay", Replay), ("indices", C.c_void_p),
                ("is_weight", C.c_void_p), ("B", C.c_int), ("gamma", C.c_float), ("tau", C.c_float), ("lr", C.c_double),
                ("beta1", C.c_double), ("beta2", C.c_double), ("eps_adam", C.c_double), ("step0", C.c_int64),
                ("gpart", C.c_void_p), ("stats", C.c_void_p), ("error_out", C.c_void_p), ("out", C.c_void_p),
                ("noise_gen", C.c_int), ("noise_seed", C.c_uint64), ("noise_counter", C.c_uint64)]


OPT_CAUTIOUS_ADAMW, OPT_ADAM = 0, 1
NSEG = 2 * FRL_MAX_LAYERS + 1
ACTOR_TANH, ACTOR_SAC = 0, 1
INFER_ARGMAX, INFER_TANH, INFER_SAC_SAMPLE, INFER_SAC_MEAN, INFER_RAW, INFER_PPO_GAUSS, INFER_PPO_CAT = 0, 1, 2, 3, 4, 5, 6
INFER_ARGMAX_DUELING = 7

_lib = None


def _declare(lib):
    vp, i64, u64, ci = C.c_void_p, C.c_int64, C.c_uint64, C.c_int
    lib.frl_last_error.restype = C.c_char_p
    lib.frl_replay_add_batch.argtypes = [C.POINTER(Replay), i64, vp, vp, vp, vp, vp, ci, vp]
    lib.frl_replay_gather.argtypes = [C.POINTER(Replay), vp, ci, vp, vp, vp, vp, vp, vp]
    lib.frl_sample_uniform.argtypes = [vp, i64, ci, ci, u64, u64, vp]
    lib.frl_net_sync_mirror.argtypes = [C.POINTER(Net), vp]
    lib.frl_polyak.argtypes = [C.POINTER(Net), C.POINTER(Net), C.c_float, vp]
    lib.frl_polyak.restype = ci
    lib.frl_dqn_learn.argtypes = [C.POINTER(DqnArgs), vp]
    lib.frl_ac_learn.argtypes = [C.POINTER(AcArgs), vp]
    lib.frl_ac_ws_floats.argtypes = [C.POINTER(AcArgs)]
    lib.frl_ac_ws_floats.restype = C.c_longlong
    lib.frl_ac_path.argtypes = [C.POINTER(AcArgs)]
    lib.frl_ac_path.restype = ci
    lib.frl_policy_infer.argtypes = [C.POINTER(InferArgs), vp]
    lib.frl_gae.argtypes = [vp, vp, vp, vp, vp, ci, ci, C.c_double, C.c_double, vp, vp, vp]
    lib.frl_ppo_update.argtypes = [C.POINTER(PpoArgs), vp]
    lib.frl_sumtree_update.argtypes = [vp, i64, vp, vp, vp, C.c_double, i64, ci, ci, vp, vp]
    lib.frl_sumtree_update_td.argtypes = [vp, i64, vp, vp, C.c_float, C.c_float, ci, vp, vp]
    lib.frl_sumtree_sample.argtypes = [vp, i64, vp, u64, u64, ci, i64, C.c_double, C.c_double, vp, vp, vp, vp]
    lib.frl_sumtree_max.argtypes = [vp, i64, vp, ci, vp, vp]
    lib.frl_per_priorities.argtypes = [vp, ci, C.c_float, C.c_float, vp, vp]
    lib.frl_rainbow_learn.argtypes = [C.POINTER(RainbowArgs), vp]
    lib.frl_rainbow_act.argtypes = [C.POINTER(RainbowArgs), vp, ci, vp, vp]
    lib.frl_adv_norm.argtypes = [vp, ci, C.c_float, vp, vp]
    lib.frl_vecnorm.argtypes = [vp, i64, vp, ci, ci, ci, ci, vp, vp, vp]
    lib.frl_reward_scaling.argtypes = [vp, i64, vp, vp, ci, C.c_double, ci, vp, vp, vp]
    lib.frl_explore.argtypes = [C.POINTER(ExploreArgs), vp]
    lib.frl_masked_reset.argtypes = [vp, vp, ci, ci, C.c_double, vp]
    lib.frl_epsilon_greedy.argtypes = [vp, ci, ci, C.c_double, vp, vp, u64, u64, vp, vp]
    lib.frl_dis_to_con.argtypes = [vp, ci, ci, ci, ci, vp, vp, vp, vp, vp]
    for name in ("frl_vecnorm", "frl_reward_scaling", "frl_explore", "frl_masked_reset", "frl_epsilon_greedy", "frl_dis_to_con"):
        getattr(lib, name).restype = ci
    lib.frl_wt_ld.argtypes = [ci]
    lib.frl_ppo_umma_ws_floats.restype = C.c_longlong
    lib.frl_launch_count.restype = C.c_longlong
    lib.frl_debug_randn.argtypes = [u64, C.c_uint32, C.c_uint32, C.c_longlong, vp, vp]
    lib.frl_debug_randn.restype = ci
    lib.frl_dp_alloc.argtypes = [C.c_longlong, C.POINTER(C.c_void_p), C.c_char_p]
    lib.frl_dp_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.frl_dp_close.argtypes = [vp]
    lib.frl_dp_free.argtypes = [vp]
    for name in ("frl_dp_alloc", "frl_dp_open", "frl_dp_close", "frl_dp_free"):
        getattr(lib, name).restype = ci
    lib.frl_launch_count.argtypes = []
    lib.frl_ppo_umma_ws_floats.argtypes = []
    lib.frl_adv_norm.restype = ci
    lib.frl_rainbow_learn.restype = ci
    lib.frl_sacd_learn.argtypes = [C.POINTER(SacdArgs), vp]
    lib.frl_sacd_learn.restype = ci
    lib.frl_replica_average.argtypes = [C.POINTER(ReplicaAvgArgs), vp]
    lib.frl_replica_average.restype = ci
    lib.frl_rainbow_act.restype = ci
    for name in ("frl_replay_add_batch", "frl_replay_gather", "frl_sample_uniform", "frl_net_sync_mirror",
                 "frl_dqn_learn", "frl_ac_learn", "frl_policy_infer", "frl_gae", "frl_ppo_update", "frl_sumtree_update", "frl_sumtree_update_td", "frl_sumtree_sample", "frl_sumtree_max",
                 "frl_per_priorities", "frl_is_emulation", "frl_device_sm_count", "frl_wt_ld",
                 "frl_abi_version"):
        getattr(lib, name).restype = ci


def library_path():
    return os.environ.get("FREERL_B200_LIB", os.path.join(_HERE, "libfreerl_b200.so"))


def lib():
    """Load (once) and return the shared library.  Raises if it is absent — there is no CPU/eager fallback."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise RuntimeError(
                "freerl_b200: CUDA extension %s not found. Build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). There is no CPU fallback." % path)
        l = C.CDLL(path)
        _declare(l)
        if l.frl_abi_version() != ABI_VERSION:
            raise RuntimeError("freerl_b200: ABI mismatch in %s (library %d, binding %d): rebuild the extension"
                               % (path, l.frl_abi_version(), ABI_VERSION))
        l.frl_struct_size.restype = C.c_int
        l.frl_struct_size.argtypes = [C.c_int]
        for which, mirror in enumerate((Layer, Net, Replay, DqnArgs, AcArgs, InferArgs, PpoArgs, NoisyMap, RainbowArgs, ExploreArgs, SacdArgs, ReplicaAvgArgs)):
            if l.frl_struct_size(which) != C.sizeof(mirror):
                raise RuntimeError("freerl_b200: ctypes mirror %s is %d bytes, the library's struct is %d"
                                   % (mirror.__name__, C.sizeof(mirror), l.frl_struct_size(which)))
        _lib = l
    return _lib


def is_emulation():
    return bool(lib().frl_is_emulation())


def check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (what, rc, lib().frl_last_error().decode()))


def require_device(device):
    """The product path runs on CUDA only.  (The host-emulation library used by the CPU unit tests reports
    ``frl_is_emulation() == 1`` and is the only thing that may be driven with CPU tensors.)"""
    device = torch.device(device)
    if device.type != "cuda" and not is_emulation():
        raise RuntimeError("freerl_b200 runs on CUDA devices only (got device=%s); there is no CPU path" % device)
    return device


def stream_ptr(device):
    """Stream every library call launches on.  The library launches on the CURRENT CUDA device, so a policy that lives on another device
    than the current one (cuda:1 without torch.cuda.set_device(1), or one process driving two GPUs) first makes its device current —
    otherwise its kernels would run on the wrong GPU with pointers of another one."""
    if device.type == "cuda":
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if torch.cuda.current_device() != idx:
            torch.cuda.set_device(idx)
        return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    return C.c_void_p(0)


def ptr(t):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


_sm_count = None


def sm_count():
    global _sm_count
    if _sm_count is None:
        _sm_count = int(lib().frl_device_sm_count())
    return _sm_count
