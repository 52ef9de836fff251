"""Rainbow DQN with the reference's class API (``DQN_file/DQN_with_tricks.py:40-308``) on the fused B200 kernels.

``DQN(dim_info, is_continue, Qnet_lr, buffer_size, device, trick=None, gamma=None, batch_size=None)`` with the six
tricks ``Double / Dueling / PER / Noisy / N_Step / Categorical``.  Two fused paths:

* distributional core (Categorical + Dueling + Noisy — the ``DQN_Rainbow_`` configuration, the script's default trick set,
  SURVEY App. A) with ``Double``, ``PER`` and ``N_Step`` switchable -> ``frl_rainbow_learn``;
* the non-distributional branch (``DQN_with_tricks.py:261-283``): ``MLP`` or ``Dueling`` value net with any of ``Double``,
  ``PER``, ``N_Step`` -> ``frl_dqn_learn`` with its trick flags.

``Noisy`` without ``Categorical`` and ``Categorical`` without ``Dueling`` + ``Noisy`` raise ``NotImplementedError``.

Reference behaviour kept: ``NoisyLinear`` resamples factorised noise from the torch CPU generator on EVERY forward
(V then A, ``randn(in)`` then ``randn(out)``) — three forwards per ``learn`` (online(s') for Double, target(s'),
online(s)) plus one per ``select_action``; ``NoisyLinear.__init__`` ends with ``torch.manual_seed(100)``
(``Noisy_net.py:37``); PER priorities come from ``(m * log p).sum(1)``; the n-step buffer defaults to n = 3 with PER.
"""
import ctypes
import math
import os

import numpy as np
import torch
import torch.nn as nn

from . import _common, _lib
from .Buffer import Buffer
from .nets import DeviceNet, pad4
from .per import N_Step_Buffer, N_Step_PER_Buffer, PER_Buffer

HIDDEN = 128


class _NoisyShim(nn.Module):
    """parameter container with the reference's NoisyLinear names (``Noisy_net.py:24-31``)"""

    def __init__(self, views, out_dim, in_dim, device):
        super().__init__()
        for k in ("weight_mu", "weight_sigma", "bias_mu", "bias_sigma"):
            self.register_parameter(k, nn.Parameter(views[k], requires_grad=False))
        self.register_buffer("weight_epsilon", torch.zeros(out_dim, in_dim, device=device))
        self.register_buffer("bias_epsilon", torch.zeros(out_dim, device=device))
        self.is_train = True


class _LinearShim(nn.Module):
    def __init__(self, w, b):
        super().__init__()
        self.weight = nn.Parameter(w, requires_grad=False)
        self.bias = nn.Parameter(b, requires_grad=False)


class _CategoricalShim(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError("freerl_b200 modules are parameter containers; use select_action()")


def _noisy_reference_init(in_dim, out_dim, sigma_init=0.05):
    """Same torch RNG consumption / side effects as ``NoisyLinear.__init__`` (``Noisy_net.py:17-76``)."""
    mu_range = 1 / math.sqrt(in_dim)
    w_mu = torch.empty(out_dim, in_dim).uniform_(-mu_range, mu_range)
    b_mu = torch.empty(out_dim).uniform_(-mu_range, mu_range)
    w_sg = torch.full((out_dim, in_dim), sigma_init / math.sqrt(in_dim))
    b_sg = torch.full((out_dim,), sigma_init / math.sqrt(out_dim))
    e_in, e_out = torch.randn(in_dim), torch.randn(out_dim)            # reset_noise()
    torch.manual_seed(100)                                              # Noisy_net.py:37
    return w_mu, w_sg, b_mu, b_sg, e_in, e_out


def f_noise(x):
    return x.sign() * torch.sqrt(abs(x))


class _Block:
    """One trainable block (online or target) with torch-shaped views and its parameter-container module."""

    def __init__(self, obs_dim, n_out_a, n_atoms, device):
        ip = pad4(obs_dim)
        self.off = {}
        o = 0
        for name, n in (("l1.w", HIDDEN * ip), ("l1.b", HIDDEN), ("V.wmu", n_atoms * HIDDEN), ("V.wsg", n_atoms * HIDDEN),
                        ("V.bmu", n_atoms), ("V.bsg", n_atoms), ("A.wmu", n_out_a * HIDDEN), ("A.wsg", n_out_a * HIDDEN),
                        ("A.bmu", n_out_a), ("A.bsg", n_out_a)):
            self.off[name] = o
            o += pad4(n)
        self.n = o
        self.p = torch.zeros(o, dtype=torch.float32, device=device)
        v = lambda name, *shape: self.p[self.off[name]:self.off[name] + int(np.prod(shape))].view(*shape)
        self.l1_w = v("l1.w", HIDDEN, ip)[:, :obs_dim]
        self.l1_b = v("l1.b", HIDDEN)
        self.V = dict(weight_mu=v("V.wmu", n_atoms, HIDDEN), weight_sigma=v("V.wsg", n_atoms, HIDDEN),
                      bias_mu=v("V.bmu", n_atoms), bias_sigma=v("V.bsg", n_atoms))
        self.A = dict(weight_mu=v("A.wmu", n_out_a, HIDDEN), weight_sigma=v("A.wsg", n_out_a, HIDDEN),
                      bias_mu=v("A.bmu", n_out_a), bias_sigma=v("A.bsg", n_out_a))
        self.module = _CategoricalShim()
        self.module.l1 = _LinearShim(self.l1_w, self.l1_b)
        self.module.V = _NoisyShim(self.V, n_atoms, HIDDEN, device)
        self.module.A = _NoisyShim(self.A, n_out_a, HIDDEN, device)


class Agent:
    def __init__(self, obs_dim, action_dim, Qnet_lr, device, trick=None, batch_size=None, n_atoms=51):
        if not (trick['Categorical'] and trick['Dueling'] and trick['Noisy']):
            raise NotImplementedError("freerl_b200 fuses the Categorical + Dueling + Noisy network (Rainbow); "
                                      "for plain DQN use freerl_b200.DQN")
        self.n_atoms, self.nA, self.obs_dim = n_atoms, action_dim, obs_dim
        n_out_a = action_dim * n_atoms
        self.online = _Block(obs_dim, n_out_a, n_atoms, device)
        self.target = _Block(obs_dim, n_out_a, n_atoms, device)
        # reference construction order: l1 = nn.Linear, V = NoisyLinear(hidden, atoms), A = NoisyLinear(hidden, A*atoms)
        l1 = nn.Linear(obs_dim, HIDDEN)
        vi = _noisy_reference_init(HIDDEN, n_atoms)
        ai = _noisy_reference_init(HIDDEN, n_out_a)
        with torch.no_grad():
            self.online.l1_w.copy_(l1.weight)
            self.online.l1_b.copy_(l1.bias)
            for blk, init in ((self.online.V, vi), (self.online.A, ai)):
                for k, val in zip(("weight_mu", "weight_sigma", "bias_mu", "bias_sigma"), init[:4]):
                    blk[k].copy_(val)
        self.target.p.copy_(self.online.p)                          # deepcopy(self.Qnet)
        self.m = torch.zeros_like(self.online.p)
        self.v = torch.zeros_like(self.online.p)
        self.Qnet, self.Qnet_target = self.online.module, self.target.module
        # last raw noise of each net (weight_epsilon buffers are materialised lazily for state_dict())
        self._last_eps = {"online": (vi[4], vi[5], ai[4], ai[5]), "target": (vi[4], vi[5], ai[4], ai[5])}
        self._refresh_buffers()
        self.lr, self.step = Qnet_lr, 0
        # effective nets: l1, V, A column blocks of <= 128 outputs
        blocks, c = [], 0
        while c < n_out_a:
            w = min(128, n_out_a - c)
            blocks.append((c, w))
            c += w
        if 2 + len(blocks) > _lib.FRL_MAX_LAYERS:
            raise NotImplementedError("too many actions for the fused Rainbow head (%d column blocks)" % len(blocks))
        self.blocks = blocks
        dims = [(obs_dim, HIDDEN), (HIDDEN, n_atoms)] + [(HIDDEN, w) for _, w in blocks]
        self.eff = [DeviceNet(dims, device, trainable=False) for _ in range(3)]
        self.eps_len = pad4(2 * HIDDEN + n_atoms + n_out_a)
        self.eps_off = dict(V_in=0, V_out=HIDDEN, A_in=HIDDEN + n_atoms, A_out=2 * HIDDEN + n_atoms)

    def _refresh_buffers(self):
        for key, mod in (("online", self.Qnet), ("target", self.Qnet_target)):
            last = self._last_eps[key]
            if len(last) == 2 and isinstance(last[0], str):          # fast mode: ("f", already transformed device slices)
                vin, vout, ain, aout = [x.detach().cpu() for x in last[1]]
            else:
                vin, vout, ain, aout = [f_noise(x) for x in last]
            dev = mod.V.weight_epsilon.device
            mod.V.weight_epsilon.copy_(torch.ger(vout, vin).to(dev)); mod.V.bias_epsilon.copy_(vout.to(dev))
            mod.A.weight_epsilon.copy_(torch.ger(aout, ain).to(dev)); mod.A.bias_epsilon.copy_(aout.to(dev))


TRICKS = ("Double", "Dueling", "PER", "Noisy", "N_Step", "Categorical")


class DQN(_common.ReplicaSyncMixin):
    """Dispatches on the trick set like the reference's ``Agent.__init__`` (``DQN_with_tricks.py:163-172``)."""

    def _replica_pairs(self):
        ag = self.agent
        if hasattr(ag, "online"):                 # Rainbow: flat parameter blocks (the noisy effective weights are rebuilt by every learn)
            return [(ag.online.p, None), (ag.target.p, None)]
        return [(ag._q.p, ag._q.sync_mirror), (ag._qt.p, ag._qt.sync_mirror)]

    def __new__(cls, dim_info=None, is_continue=None, Qnet_lr=None, buffer_size=None, device=None, trick=None, *args, **kw):
        if cls is DQN:
            cls = _RainbowDQN if (trick or {}).get("Categorical") else _ValueDQN
        return object.__new__(cls)

    @staticmethod
    def load(dim_info, is_continue, model_dir, trick=None, device=None, gamma=0.99, batch_size=256):
        """The reference's ``load`` omits gamma/batch_size and crashes for Categorical/N_Step; here they default."""
        device = device if device is not None else torch.device("cuda")
        policy = DQN(dim_info, is_continue, 0, 0, device=device, trick=trick, gamma=gamma, batch_size=batch_size)
        policy.agent.Qnet.load_state_dict(torch.load(os.path.join(model_dir, "DQN.pt"), map_location=device))
        if (trick or {}).get('Noisy'):
            for module in policy.agent.Qnet.children():
                if isinstance(module, _NoisyShim):
                    module.is_train = False
        return policy


class _ValueModule(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError("freerl_b200 modules are parameter containers; use select_action()")


def _value_shim(net, dueling):
    """state_dict schema of the reference's ``MLP`` (l1, l2) / ``Dueling`` (l1, V, A): the dueling head is ONE device layer
    whose rows are [V | A_0..A_{n-1}]"""
    mod = _ValueModule()
    mod.l1 = _LinearShim(net.weight(0), net.bias(0))
    if dueling:
        mod.V = _LinearShim(net.weight(1)[0:1], net.bias(1)[0:1])
        mod.A = _LinearShim(net.weight(1)[1:], net.bias(1)[1:])
    else:
        mod.l2 = _LinearShim(net.weight(1), net.bias(1))
    mod.register_load_state_dict_post_hook(lambda module, incompatible: net.sync_mirror())
    return mod


class _ValueAgent:
    def __init__(self, obs_dim, action_dim, Qnet_lr, device, dueling):
        head = (1 + action_dim) if dueling else action_dim
        dims = [(obs_dim, HIDDEN), (HIDDEN, head)]
        self._q = DeviceNet(dims, device, trainable=True)
        self._qt = DeviceNet(dims, device, trainable=False)
        l1 = nn.Linear(obs_dim, HIDDEN)                         # reference construction order (RNG consumption)
        with torch.no_grad():
            self._q.weight(0).copy_(l1.weight)
            self._q.bias(0).copy_(l1.bias)
            if dueling:
                V, A = nn.Linear(HIDDEN, 1), nn.Linear(HIDDEN, action_dim)
                self._q.weight(1)[0:1].copy_(V.weight); self._q.bias(1)[0:1].copy_(V.bias)
                self._q.weight(1)[1:].copy_(A.weight); self._q.bias(1)[1:].copy_(A.bias)
            else:
                l2 = nn.Linear(HIDDEN, action_dim)
                self._q.weight(1).copy_(l2.weight); self._q.bias(1).copy_(l2.bias)
        self._q.sync_mirror()
        self._qt.copy_from(self._q)                             # deepcopy(self.Qnet)
        self.Qnet, self.Qnet_target = _value_shim(self._q, dueling), _value_shim(self._qt, dueling)
        self.lr, self.step = Qnet_lr, 0


def _make_buffer(trick, buffer_size, obs_dim, act_dim, device, gamma, mode):
    """``DQN_with_tricks.py:186-193``"""
    if trick['PER'] and trick['N_Step']:
        return N_Step_PER_Buffer(buffer_size, obs_dim, act_dim=act_dim, device=device, gamma=gamma, mode=mode)
    if trick['PER']:
        return PER_Buffer(buffer_size, obs_dim, act_dim=act_dim, device=device, mode=mode)
    if trick['N_Step']:
        return N_Step_Buffer(buffer_size, obs_dim, act_dim=act_dim, device=device, gamma=gamma)
    return Buffer(buffer_size, obs_dim, act_dim=act_dim, device=device)


class _ValueDQN(DQN):
    """Non-distributional branch: MLP / Dueling value net + Double / PER / N-step (``DQN_with_tricks.py:261-283``)."""

    def __init__(self, dim_info, is_continue, Qnet_lr, buffer_size, device, trick=None, gamma=None, batch_size=None, mode=None):
        obs_dim, action_dim = dim_info
        self.device = _lib.require_device(device)
        trick = dict({k: False for k in TRICKS}, **(trick or {}))
        if trick['Noisy']:
            raise NotImplementedError("freerl_b200: NoisyLinear is fused only in the Categorical (Rainbow) kernel")
        self.trick = trick
        self.agent = _ValueAgent(obs_dim, action_dim, Qnet_lr, self.device, bool(trick['Dueling']))
        self.buffer = _make_buffer(trick, buffer_size, obs_dim, action_dim if is_continue else 1, self.device, gamma, mode)
        self.is_continue = is_continue
        self.obs_dim, self.action_dim = obs_dim, action_dim
        self.mode = _common.resolve_mode(mode)
        self._scratch = _common.DeviceScratch(self.device, self.agent._q.n_p)
        self._seed = _common.default_seed()
        self._n_learn = 0
        self.last_metrics = None

    def select_action(self, obs):
        if self.is_continue:
            raise RuntimeError("DQN is not suitable for continuous action spaces (use dis_to_con)")
        x, single = _common.as_obs_batch(obs, self.obs_dim)
        mode = _lib.INFER_ARGMAX_DUELING if self.trick['Dueling'] else _lib.INFER_ARGMAX
        a = _common.infer(self.agent._q, x, mode, self.device, 1).reshape(-1).to(torch.int64).cpu().numpy()
        return a[0] if single else a

    def evaluate_action(self, obs):
        return self.select_action(obs)

    def add(self, obs, action, reward, next_obs, done):
        self.buffer.add(obs, action, reward, next_obs, done)

    sample = None      # bound below (same body as the distributional class)

    def learn(self, batch_size, gamma, tau, *, u=None, indices=None):
        total = len(self.buffer)
        B = min(batch_size, total)
        if B <= 0:
            raise RuntimeError("learn() called on an empty replay (with N_Step the first n_step - 1 adds only fill the window)")
        per = bool(self.trick['PER'])
        if per:
            idx, w, _ = self.buffer.sample_device(B, u=u)
        else:
            w = None
            idx = _common.make_indices(self.mode, total, B, 1, self.device, self._seed, self._n_learn).reshape(-1) if indices is None \
                else self.buffer._indices_to_device(indices).reshape(-1)
            B = idx.numel()
        if self.trick['N_Step']:
            gamma = self.buffer.n_step_gamma
        ag = self.agent
        td = torch.empty((B, 1), dtype=torch.float32, device=self.device)
        a = _lib.DqnArgs()
        a.q, a.q_target = ag._q.c_struct(), ag._qt.c_struct()
        a.replay = (self.buffer.buffer if per else self.buffer).c_struct()
        a.indices, a.B, a.n_updates = idx.data_ptr(), B, 1
        a.gamma, a.tau = gamma, tau
        a.lr, a.beta1, a.beta2, a.eps = ag.lr, 0.9, 0.999, 1e-8
        a.step0 = ag.step
        out = self._scratch.out(1, self.device)
        a.gpart, a.stats, a.out = self._scratch.gpart.data_ptr(), self._scratch.stats.data_ptr(), out.data_ptr()
        a.double_q, a.dueling = int(bool(self.trick['Double'])), int(bool(self.trick['Dueling']))
        a.is_weight = w.data_ptr() if w is not None else None
        a.td_error = td.data_ptr()
        _lib.check(_lib.lib().frl_dqn_learn(ctypes.byref(a), _lib.stream_ptr(self.device)), "frl_dqn_learn")
        ag.step += 1
        self._n_learn += 1
        if per:
            self.buffer.update_priorities(idx, td)                  # td_error [B,1] like the reference (:278)
        self.last_metrics = out[0]
        self.last_error, self.last_indices = td, idx

    def update_target(self, tau):
        t, s_ = self.agent._qt, self.agent._q
        t.p.mul_(1.0 - tau).add_(s_.p * tau)
        t.sync_mirror()

    def save(self, model_dir):
        torch.save({k: v.detach().clone().cpu() for k, v in self.agent.Qnet.state_dict().items()}, os.path.join(model_dir, "DQN.pt"))


class _RainbowDQN(DQN):
    def __init__(self, dim_info, is_continue, Qnet_lr, buffer_size, device, trick=None, gamma=None, batch_size=None, mode=None):
        obs_dim, action_dim = dim_info
        self.device = _lib.require_device(device)
        trick = dict({k: False for k in TRICKS}, **(trick or {}))
        self.trick = trick
        self.agent = Agent(obs_dim, action_dim, Qnet_lr, self.device, trick=trick, batch_size=batch_size)
        act_dim = action_dim if is_continue else 1
        if trick['PER'] and trick['N_Step']:
            self.buffer = N_Step_PER_Buffer(buffer_size, obs_dim, act_dim=act_dim, device=self.device, gamma=gamma, mode=mode)
        elif trick['PER']:
            self.buffer = PER_Buffer(buffer_size, obs_dim, act_dim=act_dim, device=self.device, mode=mode)
        elif trick['N_Step']:
            self.buffer = N_Step_Buffer(buffer_size, obs_dim, act_dim=act_dim, device=self.device, gamma=gamma)
        else:
            self.buffer = Buffer(buffer_size, obs_dim, act_dim=act_dim, device=self.device)
        self.is_continue = is_continue
        self.obs_dim, self.action_dim = obs_dim, action_dim
        self.mode = _common.resolve_mode(mode)
        ag = self.agent
        self.v_min, self.v_max = -100.0, 100.0
        self.z = torch.linspace(self.v_min, self.v_max, steps=ag.n_atoms).to(self.device)
        self.delta_z = (self.v_max - self.v_min) / (ag.n_atoms - 1)
        sm = _lib.sm_count()
        self._gpart = torch.zeros((sm, ag.eff[2].n_p), dtype=torch.float32, device=self.device)
        self._stats = torch.zeros((sm, 8), dtype=torch.float32, device=self.device)
        self._out = torch.zeros(8, dtype=torch.float32, device=self.device)
        self._eps = torch.zeros((3, ag.eps_len), dtype=torch.float32, device=self.device)
        self._seed = _common.default_seed()
        self._n_learn = 0
        self.last_metrics = None

    # ---- noise -------------------------------------------------------------------------------------------
    def _draw_forward_noise(self):
        """raw (V_in, V_out, A_in, A_out) of one forward, drawn like ``NoisyLinear.reset_noise`` on the CPU generator"""
        ag = self.agent
        return (torch.randn(HIDDEN), torch.randn(ag.n_atoms), torch.randn(HIDDEN), torch.randn(ag.nA * ag.n_atoms))

    def _device_eps(self):
        """fast mode: the transformed noise f(eps) = sign(eps) sqrt|eps| of all three forwards drawn ON the device (no host
        round trip); returns per-forward 4-tuples of slices for the weight_epsilon bookkeeping"""
        e = self._eps
        e.normal_()
        torch.mul(torch.sign(e), torch.sqrt(torch.abs(e)), out=e)
        return self._eps_slices()

    def _eps_slices(self):
        ag, e, o = self.agent, self._eps, self.agent.eps_off
        return [(e[f, o["V_in"]:o["V_in"] + HIDDEN], e[f, o["V_out"]:o["V_out"] + ag.n_atoms], e[f, o["A_in"]:o["A_in"] + HIDDEN],
                 e[f, o["A_out"]:o["A_out"] + ag.nA * ag.n_atoms]) for f in range(3)]

    def _pack_eps(self, raws):
        """raws: list of 3 (or fewer) 4-tuples -> device [3, eps_len] of transformed noise"""
        ag = self.agent
        host = torch.zeros((3, ag.eps_len), dtype=torch.float32)
        for f, raw in enumerate(raws):
            if raw is None:
                continue
            o = ag.eps_off
            host[f, o["V_in"]:o["V_in"] + HIDDEN] = f_noise(torch.as_tensor(raw[0]))
            host[f, o["V_out"]:o["V_out"] + ag.n_atoms] = f_noise(torch.as_tensor(raw[1]))
            host[f, o["A_in"]:o["A_in"] + HIDDEN] = f_noise(torch.as_tensor(raw[2]))
            host[f, o["A_out"]:o["A_out"] + ag.nA * ag.n_atoms] = f_noise(torch.as_tensor(raw[3]))
        self._eps.copy_(host.to(self.device))

    def _args(self):
        ag = self.agent
        a = _lib.RainbowArgs()
        a.p, a.m, a.v, a.p_target, a.n_train = ag.online.p.data_ptr(), ag.m.data_ptr(), ag.v.data_ptr(), ag.target.p.data_ptr(), ag.online.n
        for f in range(3):
            a.eff[f] = ag.eff[f].c_struct()
        off, eo = ag.online.off, ag.eps_off
        def setmap(i, mu_w, sg_w, mu_b, sg_b, row0, e_in, e_out):
            mp = a.map[i]
            mp.mu_w, mp.sg_w, mp.mu_b, mp.sg_b, mp.row0, mp.eps_in, mp.eps_out = mu_w, sg_w, mu_b, sg_b, row0, e_in, e_out
        setmap(0, off["l1.w"], -1, off["l1.b"], -1, 0, -1, -1)
        setmap(1, off["V.wmu"], off["V.wsg"], off["V.bmu"], off["V.bsg"], 0, eo["V_in"], eo["V_out"])
        for b, (c0, w) in enumerate(ag.blocks):
            setmap(2 + b, off["A.wmu"], off["A.wsg"], off["A.bmu"], off["A.bsg"], c0, eo["A_in"], eo["A_out"])
        a.eps, a.eps_len = self._eps.data_ptr(), ag.eps_len
        a.n_actions, a.n_atoms = ag.nA, ag.n_atoms
        a.z, a.v_min, a.v_max, a.delta_z = self.z.data_ptr(), self.v_min, self.v_max, self.delta_z
        a.double_q = int(bool(self.trick['Double']))
        buf = self.buffer.buffer if self.trick['PER'] else self.buffer
        a.replay = buf.c_struct()
        a.lr, a.beta1, a.beta2, a.eps_adam = ag.lr, 0.9, 0.999, 1e-8
        a.step0 = ag.step
        a.gpart, a.stats, a.out = self._gpart.data_ptr(), self._stats.data_ptr(), self._out.data_ptr()
        return a

    # ---- acting ------------------------------------------------------------------------------------------
    def select_action(self, obs, *, noise=None):
        if self.is_continue:
            raise RuntimeError("DQN is not suitable for continuous action spaces (use dis_to_con)")
        x, single = _common.as_obs_batch(obs, self.obs_dim)
        xd = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(self.device) if not isinstance(x, torch.Tensor) else x.to(self.device).float().contiguous()
        train = getattr(self.agent.Qnet.V, "is_train", True)
        if noise is None and train and self.mode == "fast":
            self.agent._last_eps["online"] = ("f", self._device_eps()[0])
        else:
            raw = noise if noise is not None else (self._draw_forward_noise() if train else tuple(torch.zeros(n) for n in (HIDDEN, self.agent.n_atoms, HIDDEN, self.agent.nA * self.agent.n_atoms)))
            if train:
                self.agent._last_eps["online"] = tuple(torch.as_tensor(r) for r in raw)
            self._pack_eps([raw, None, None])
        out = torch.empty(xd.shape[0], dtype=torch.float32, device=self.device)
        a = self._args()
        _lib.check(_lib.lib().frl_rainbow_act(ctypes.byref(a), _lib.ptr(xd), xd.shape[0], _lib.ptr(out), _lib.stream_ptr(self.device)),
                   "frl_rainbow_act")
        act = out.to(torch.int64).cpu().numpy()
        return act[0] if single else act

    def evaluate_action(self, obs):
        return self.select_action(obs)

    def add(self, obs, action, reward, next_obs, done):
        self.buffer.add(obs, action, reward, next_obs, done)

    def sample(self, batch_size):
        total_size = len(self.buffer)
        batch_size = min(batch_size, total_size)
        if self.trick['PER']:
            indices, is_weight = self.buffer.sample(batch_size)
            return (*self.buffer.buffer.sample(indices), is_weight, indices)
        indices = np.random.choice(total_size, batch_size, replace=False)
        return self.buffer.sample(indices)

    # ---- learning ----------------------------------------------------------------------------------------
    def learn(self, batch_size, gamma, tau, *, u=None, noise=None, indices=None):
        total = len(self.buffer)
        B = min(batch_size, total)
        if B <= 0:
            raise RuntimeError("learn() called on an empty replay (with N_Step the first n_step - 1 adds only fill the window)")
        per = bool(self.trick['PER'])
        if per:
            idx, w, _ = self.buffer.sample_device(B, u=u)
        else:
            w = None
            idx = _common.make_indices(self.mode, total, B, 1, self.device, self._seed, self._n_learn).reshape(-1) if indices is None \
                else self.buffer._indices_to_device(indices).reshape(-1)
        if self.trick['N_Step']:
            gamma = self.buffer.n_step_gamma
        gen = noise is None and self.mode == "fast"
        if gen:                                  # the kernel draws the noise itself (Philox) and leaves it in self._eps
            sl = self._eps_slices()
            self.agent._last_eps["online"], self.agent._last_eps["target"] = ("f", sl[2]), ("f", sl[1])
        else:
            if noise is None:
                noise = [self._draw_forward_noise() if self.trick['Double'] else None, self._draw_forward_noise(), self._draw_forward_noise()]
            self.agent._last_eps["online"] = tuple(torch.as_tensor(r) for r in noise[2])
            self.agent._last_eps["target"] = tuple(torch.as_tensor(r) for r in noise[1])
            self._pack_eps(noise)
        err = torch.empty(B, dtype=torch.float32, device=self.device)
        a = self._args()
        a.indices, a.B = idx.data_ptr(), B
        a.is_weight = w.data_ptr() if w is not None else None
        a.gamma, a.tau = gamma, tau
        a.error_out = err.data_ptr()
        a.noise_gen, a.noise_seed, a.noise_counter = int(gen), self._seed, self._n_learn
        _lib.check(_lib.lib().frl_rainbow_learn(ctypes.byref(a), _lib.stream_ptr(self.device)), "frl_rainbow_learn")
        self.agent.step += 1
        self._n_learn += 1
        if per:
            self.buffer.update_priorities(idx, err)
        self.last_metrics = self._out
        self.last_error, self.last_indices = err, idx

    def update_target(self, tau):
        self.agent.target.p.mul_(1.0 - tau).add_(self.agent.online.p * tau)

    # ---- checkpoint --------------------------------------------------------------------------------------
    def save(self, model_dir):
        self.agent._refresh_buffers()
        torch.save({k: v.detach().clone().cpu() for k, v in self.agent.Qnet.state_dict().items()}, os.path.join(model_dir, "DQN.pt"))


_ValueDQN.sample = _RainbowDQN.sample
