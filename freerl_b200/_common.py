"""Host-side helpers shared by the algorithm classes (index generation, scratch, batched inference)."""
import ctypes
import os

import numpy as np
import torch

from . import _lib


def resolve_mode(mode):
    """'parity' (default): indices from the numpy legacy global stream exactly like the reference
    (``np.random.choice(total, B, replace=False)``) and noise tensors drawn with torch on ``device`` in the
    reference's call order -> identical results on identical seeds.
    'fast': on-device Philox sampling / noise (different random streams, same distributions) so a whole batch of
    learn() calls is one kernel launch with no host work."""
    mode = mode or os.environ.get("FREERL_B200_MODE", "parity")
    if mode not in ("parity", "fast"):
        raise ValueError("mode must be 'parity' or 'fast'")
    return mode


class DeviceScratch:
    """Per-policy scratch the fused kernels need (gradient partials per CTA, norm partials, metrics)."""

    def __init__(self, device, n_p_max):
        sm = _lib.sm_count()
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=device)
        self.gpart = z(sm, n_p_max)
        self.sumsq = z(sm)
        self.stats = z(sm, 8)
        self._out = None
        self._xchg = None

    def xchg(self, B, device):
        if self._xchg is None or self._xchg.numel() < 3 * B:
            self._xchg = torch.zeros(3 * B, dtype=torch.float32, device=device)
        return self._xchg

    def out(self, n_updates, device):
        if self._out is None or self._out.shape[0] < n_updates:
            self._out = torch.zeros((n_updates, 8), dtype=torch.float32, device=device)
        return self._out


_fast_seed_counter = [0]


def default_seed():
    """Seed for the on-device generators, derived from torch's CPU generator state so `torch.manual_seed`
    still controls fast-mode runs."""
    _fast_seed_counter[0] += 1
    return (int(torch.initial_seed()) * 0x9E3779B97F4A7C15 + _fast_seed_counter[0]) & 0xFFFFFFFFFFFFFFFF


def make_indices(mode, total_size, batch_size, n_updates, device, seed, counter):
    """[n_updates, B] int64 device tensor of sampled transition indices (without replacement per update)."""
    if mode == "parity":
        idx = np.stack([np.random.choice(total_size, batch_size, replace=False) for _ in range(n_updates)])
        return torch.from_numpy(idx.astype(np.int64, copy=False)).to(device)
    out = torch.empty((n_updates, batch_size), dtype=torch.int64, device=device)
    _lib.check(_lib.lib().frl_sample_uniform(_lib.ptr(out), int(total_size), int(batch_size), int(n_updates),
                                             ctypes.c_uint64(seed), ctypes.c_uint64(counter), _lib.stream_ptr(device)),
               "frl_sample_uniform")
    return out


def reference_randn(shape, device):
    """What ``torch.distributions.Normal.rsample`` / ``torch.randn_like`` draw in the reference: a standard normal
    tensor created directly on ``device`` (``torch.empty(shape).normal_()``)."""
    return torch.empty(shape, dtype=torch.float32, device=device).normal_()


def infer(net, obs, mode, device, out_cols, noise=None, seed=0, counter=0, l0=0, nl=0, layer_norm=False, obs_norm=None, hidden_tanh=False):
    """Batched policy inference.  ``obs``: numpy / tensor [n, obs_dim] -> device tensor [n, out_cols]."""
    if isinstance(obs, torch.Tensor):
        x = obs.to(device=device, dtype=torch.float32).contiguous()
    else:
        x = torch.from_numpy(np.ascontiguousarray(obs, dtype=np.float32)).to(device)
    n, obs_dim = x.shape
    out = torch.empty((n, out_cols), dtype=torch.float32, device=device)
    a = _lib.InferArgs()
    a.net = net.c_struct()
    a.l0, a.nl = l0, nl
    a.layer_norm = int(layer_norm)
    a.hidden_tanh = int(hidden_tanh)
    a.obs, a.n, a.obs_dim, a.mode = x.data_ptr(), n, obs_dim, mode
    a.noise = noise.data_ptr() if noise is not None else None
    a.seed, a.counter = seed, counter & 0xFFFFFFFF
    a.out, a.out_cols = out.data_ptr(), out_cols
    a.obs_norm = obs_norm.data_ptr() if obs_norm is not None else None      # Batch_ObsNorm, update=False
    _lib.check(_lib.lib().frl_policy_infer(ctypes.byref(a), _lib.stream_ptr(device)), "frl_policy_infer")
    return out


def as_obs_batch(obs, obs_dim):
    """Reference ``select_action`` does ``reshape(1, -1)``; we accept [obs_dim] (single env, reference semantics)
    or [N, obs_dim] (vectorised extension).  Returns (array [n, obs_dim], single: bool)."""
    if isinstance(obs, torch.Tensor):
        single = obs.dim() == 1
        return obs.reshape(-1, obs_dim), single
    arr = np.asarray(obs, dtype=np.float32)
    single = arr.ndim <= 1
    return arr.reshape(-1, obs_dim), single


def device_minibatch_plan(H, minibatch_size, K_epochs, device, seed):
    """fast mode: K_epochs random permutations of range(H) cut into minibatches, built ON the device (no host loops, no H2D of
    the K * H index matrix).  Returns (indices int64 [K * nmb, mb], valid rows int32 [K * nmb], n_updates) like the host plan
    of the reference loop (``np.random.permutation(horizon)`` sliced every ``minibatch_size``, PPO.py:247-248)."""
    nmb = (H + minibatch_size - 1) // minibatch_size
    g = torch.Generator(device=device)
    g.manual_seed(int(seed) & 0x7FFFFFFF)
    perms = torch.stack([torch.randperm(H, device=device, generator=g) for _ in range(K_epochs)])
    pad = nmb * minibatch_size - H
    if pad:
        perms = torch.cat([perms, torch.zeros((K_epochs, pad), dtype=torch.int64, device=device)], dim=1)
    idx = perms.reshape(K_epochs * nmb, minibatch_size).contiguous()
    last = H - (nmb - 1) * minibatch_size
    rows = torch.tensor(([minibatch_size] * (nmb - 1) + [last]) * K_epochs, dtype=torch.int32).to(device)
    return idx, rows, K_epochs * nmb


_UMMA_WS = {}


def umma_ws_ptr(device, minibatch_size):
    """Device scratch of the tensor-core PPO / MAPPO path (``frl_ppo_args_t.umma_ws``, csrc/algo_ppo_umma.cuh): split weights +
    per-CTA activation scratch, one allocation per device shared by every policy (launches on a device are stream-ordered).
    0 when the minibatch is below the 1024-row threshold (the library then takes the FFMA tile kernels)."""
    from . import _lib
    if minibatch_size < 1024 or os.environ.get("FREERL_B200_NO_UMMA"):      # the switch exists for A/B timing (tools/configbench.py)
        return 0
    n = int(_lib.lib().frl_ppo_umma_ws_floats())
    if n <= 0:
        return 0
    key = str(device)
    if key not in _UMMA_WS:
        _UMMA_WS[key] = torch.zeros(n, dtype=torch.float32, device=device)
    return _UMMA_WS[key].data_ptr()


class DpPeers:
    """Peer-memory blocks of the in-kernel data-parallel gradient exchange (``frl_dp_peers_t``, include/freerl_b200.h): every rank
    allocates ``[flags 256 B | g: 2 x n_floats]`` with ``frl_dp_alloc``, the CUDA IPC handles travel through ``all_gather_object`` and
    each rank maps the others' blocks (``frl_dp_open``).  One block serves every network of a policy (sized for the largest): the
    exchanges of successive launches are numbered by one epoch counter that all ranks advance identically."""

    def __init__(self, dist, group, device, n_floats):
        import ctypes
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > _lib.FRL_DP_MAX_RANKS:
            raise ValueError("peer-memory data parallel supports up to %d ranks per node" % _lib.FRL_DP_MAX_RANKS)
        self.n_floats = int(n_floats)
        self.bytes = 256 + 2 * self.n_floats * 4
        lib = _lib.lib()
        ptr, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        with torch.cuda.device(device):
            _lib.check(lib.frl_dp_alloc(self.bytes, ctypes.byref(ptr), handle), "frl_dp_alloc")
            handles = [None] * self.world
            dist.all_gather_object(handles, (bytes(handle.raw), self.bytes), group=group)
            self.base = [0] * self.world
            for r, (h, nbytes) in enumerate(handles):
                if nbytes != self.bytes:
                    raise RuntimeError("rank %d allocated a different exchange block" % r)
                if r == self.rank:
                    self.base[r] = ptr.value
                else:
                    q = ctypes.c_void_p()
                    _lib.check(lib.frl_dp_open(h, ctypes.byref(q)), "frl_dp_open")
                    self.base[r] = q.value
        self.epoch = 0
        dist.barrier(group=group)

    def close(self, dist=None, group=None):
        """Unmap the peers' blocks and free this rank's own.  Call it on every rank at the same point (pass ``dist`` to put the barriers
        in: nobody may still be reading a block that is being freed).  Not called implicitly — at interpreter exit the driver
        reclaims the blocks."""
        lib = _lib.lib()
        if dist is not None:
            dist.barrier(group=group)
        for r, base in enumerate(self.base):
            if r != self.rank and base:
                _lib.check(lib.frl_dp_close(base), "frl_dp_close")
        if dist is not None:
            dist.barrier(group=group)
        if self.base[self.rank]:
            _lib.check(lib.frl_dp_free(self.base[self.rank]), "frl_dp_free")
        self.base = [0] * self.world

    def fill(self, a, n_floats, n_updates):
        """Point ``a.dp`` (a ``PpoArgs``) at the blocks for a launch of ``n_updates`` exchanges of ``n_floats`` gradients each."""
        if n_floats > self.n_floats:
            raise ValueError("network larger than the exchange block")
        for r in range(self.world):
            a.dp.flags[r] = self.base[r]
            a.dp.g[r] = self.base[r] + 256
        a.dp.rank, a.dp.world, a.dp.epoch0 = self.rank, self.world, self.epoch & 0xFFFFFFFF
        self.epoch += n_updates


def replica_peers(dist, group, device, tensors):
    """Exchange blocks for the in-kernel replica average (``frl_replica_average``) of these tensors, or None where the peer path does
    not apply (one rank, CPU / the test-only emulation, more ranks than a node holds, ``FREERL_B200_DP_PEER=0``): the callers then keep the
    ``dist.all_reduce`` + divide pairs."""
    world = dist.get_world_size(group)
    if (world < 2 or world > _lib.FRL_DP_MAX_RANKS or len(tensors) > _lib.FRL_RA_MAX_TENSORS or device.type != "cuda"
            or _lib.lib().frl_is_emulation() or os.environ.get("FREERL_B200_DP_PEER", "1") == "0"):
        return None
    total = sum((t.numel() + 3) & ~3 for t in tensors)
    peers = DpPeers(dist, group, device, total)
    peers.status = torch.zeros(1, dtype=torch.float32, device=device)
    return peers


def replica_average(peers, tensors, device):
    """tensors[i] <- mean over the ranks (rank-ordered sum / world), ONE cooperative launch over peer-mapped NVLink memory"""
    import ctypes
    a = _lib.ReplicaAvgArgs()
    peers.fill(a, sum((t.numel() + 3) & ~3 for t in tensors), 1)
    for i, t in enumerate(tensors):
        a.tensor[i], a.n[i] = t.data_ptr(), t.numel()
    a.n_tensors, a.block_floats, a.status = len(tensors), peers.n_floats, peers.status.data_ptr()
    _lib.check(_lib.lib().frl_replica_average(ctypes.byref(a), _lib.stream_ptr(device)), "frl_replica_average")


class ReplicaSyncMixin:
    """Multi-GPU mode of the off-policy value learners (SURVEY §8e "replicas + sharded replay"; DQN / Rainbow = BASELINE config 4):
    one process per GPU, each with its own env shard and its own replay shard — for PER its own sum-tree, priorities never leave the
    GPU that sampled them — and no data-path collective.  The replicas are kept ONE policy by a parameter average over
    ``torch.distributed`` (NCCL over NVLink): ``enable_replica_sync()`` broadcasts rank 0's parameters, ``sync_replicas()`` averages
    online and target parameters (call it every K learns; Adam moments stay local, the standard local-SGD scheme).
    A class provides ``_replica_pairs() -> [(parameter block tensor, refresh callable or None), ...]``."""

    def enable_replica_sync(self, group=None):
        import torch.distributed as dist
        self._rs = (dist, group, dist.get_world_size(group))
        for t, refresh in self._replica_pairs():
            dist.broadcast(t, src=0, group=group)
            if refresh:
                refresh()
        self._rs_peers = replica_peers(dist, group, self.device, [t for t, _ in self._replica_pairs()])
        self.replica_collective = ("in-kernel peer-memory average (frl_replica_average), one launch per sync" if self._rs_peers
                                   else "dist.all_reduce + divide per parameter block")

    def sync_replicas(self):
        rs = getattr(self, "_rs", None)
        if rs is None or rs[2] == 1:
            return
        dist, group, world = rs
        pairs = self._replica_pairs()
        if getattr(self, "_rs_peers", None) is not None:
            replica_average(self._rs_peers, [t for t, _ in pairs], self.device)
            for _, refresh in pairs:
                if refresh:
                    refresh()
            return
        for t, refresh in pairs:
            dist.all_reduce(t, group=group)
            t.div_(world)
            if refresh:
                refresh()
