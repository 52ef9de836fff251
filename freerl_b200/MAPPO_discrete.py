"""MAPPO with SHARED networks over homogeneous discrete agents, with the reference's class API
(``MAPPO_file/MAPPO_discrete.py:205-404`` + ``MAPPO_file/Buffer.py:386-431``) on the fused B200 kernels.

``ReplayBuffer(N, obs_dim, state_dim, episode_limit, batch_size, device)`` holds ``batch_size`` EPISODES of ``episode_limit`` steps;
``MAPPO(dim_info, is_continue, actor_lr, critic_lr, horizon, device, trick=None, buffer=None)``; ``select_action(obs_n) -> (actions [N],
log_probs [N])``, ``evaluate_action``, ``get_value(s) -> [N]``, ``add(obs, action, reward, next_obs, done, action_log_pi, adv_dones,
episode_step)``, ``learn(minibatch_size, gamma, lmbda, clip_param, K_epochs, entropy_coefficient, huber_delta=None)``, ``lr_decay``,
``save`` / ``load`` (``MAPPO_discrete.pth`` = the actor's state dict).

Reference behaviour kept:
  * one actor (own observation -> softmax over the actions) and one critic (joint observation) serve all N agents; ONE Adam over both,
    lr = ``actor_lr`` (``critic_lr`` unused), eps 1e-5 with ``adam_eps`` (``:160-166``);
  * ``update_ac`` clips the JOINT gradient norm to 10 and steps, then ``learn`` calls ``ac_optimizer.step()`` a SECOND time on the same
    gradients (``:363,371``) — ``frl_ppo_args_t.max_norm_joint`` / ``opt_repeat = 2``; the Adam step counter advances by two per minibatch;
  * GAE without a done mask on the recursion, zero tail per episode, float32 (``:302-315``) — ``frl_gae`` over [T][B*N] columns with
    ``adv_done = 0`` (float64 per-column scan, rounded once); ``adv_norm`` over the whole [B, T, N] block (``frl_adv_norm``);
  * minibatches are consecutive blocks of ``minibatch_size`` EPISODES in storage order, every epoch the same (``SequentialSampler``, ``:326``);
  * ``ValueClip`` without ``huber_loss``: element-wise ``max((clamp(V - v_old, +-clip) + v_old - v_target)^2, (V - v_target)^2)``
    (``value_loss = 2``); with ``huber_loss`` the squared maximum of the two batch-mean huber SCALARS (``:353-357``; ``value_loss = 3``:
    a forward-only pre-pass launch leaves the two sums, the update launch folds them); ``huber_loss`` alone changes nothing upstream
    (it is only read inside the ``ValueClip`` branch);
  * ``LayerNorm`` / ``feature_norm`` are ``F.layer_norm(x, x.size()[1:])``: per row when acting (2-D inputs) but, inside ``learn``, jointly
    over the (step, agent, feature) axes of each EPISODE (4-D inputs) — the group mode of ``frl_ppo_update`` (csrc/algo_ppo_group.cuh: a CTA
    owns whole episodes and recomputes the forward pass once per statistic).  The discrete actor overwrites its normalised input
    (``:113-115``), so ``feature_norm`` only reaches the critic.  The script's default trick set therefore runs as upstream.
Device layout: TIME-MAJOR rows ``r = (t * B + b) * N + n`` so that one ``frl_gae`` call scans all B*N columns; the host-side staging
arrays keep the reference's ``[B, T, N, ...]`` shapes (``ReplayBuffer.buffer``), uploaded once per ``learn`` like the reference's
``get_training_data``.

Not reproduced (raises ``NotImplementedError``): ``is_continue=True`` — upstream ``learn`` builds ``Categorical(actor(x))`` from the
Gaussian actor's ``(mean, std)`` tuple and fails.
"""
import ctypes
import os

import numpy as np
import torch

from . import _common, _lib
from .MAPPO import Agent as _Agent


class ReplayBuffer:
    """``MAPPO_file/Buffer.py:386-431``.  ``buffer`` is the reference's dict of host arrays (float32 here: the cast the reference applies in
    ``get_training_data``; ``a_n`` int64) in pinned memory; ``get_training_data`` returns the same dict as device tensors."""

    KEYS = ("obs_n", "s", "v_n", "a_n", "a_logprob_n", "r_n", "done_n")

    def __init__(self, N, obs_dim, state_dim, episode_limit, batch_size, device):
        self.N, self.obs_dim, self.state_dim = N, obs_dim, state_dim
        self.episode_limit, self.batch_size = int(episode_limit), int(batch_size)
        self.device = torch.device(device) if not isinstance(device, torch.device) else device
        self.episode_num = 0
        self.buffer = None
        self.reset_buffer()

    def reset_buffer(self):
        B, T, N = self.batch_size, self.episode_limit, self.N
        shapes = {"obs_n": (B, T, N, self.obs_dim), "s": (B, T, self.state_dim), "v_n": (B, T + 1, N), "a_n": (B, T, N),
                  "a_logprob_n": (B, T, N), "r_n": (B, T, N), "done_n": (B, T, N)}
        pin = self.device.type == "cuda"
        self._host = {k: torch.zeros(s, dtype=torch.int64 if k == "a_n" else torch.float32, pin_memory=pin) for k, s in shapes.items()}
        self.buffer = {k: v.numpy() for k, v in self._host.items()}
        self.episode_num = 0

    def store_transition(self, episode_step, obs_n, s, v_n, a_n, a_logprob_n, r_n, done_n):
        e, t, b = self.episode_num, episode_step, self.buffer
        b["obs_n"][e][t] = obs_n
        b["s"][e][t] = s
        b["v_n"][e][t] = v_n
        b["a_n"][e][t] = a_n
        b["a_logprob_n"][e][t] = a_logprob_n
        b["r_n"][e][t] = r_n
        b["done_n"][e][t] = done_n

    def store_last_value(self, episode_step, v_n):
        self.buffer["v_n"][self.episode_num][episode_step] = v_n
        self.episode_num += 1

    def get_training_data(self):
        # read through `buffer` (the reference's attribute): normally views of the pinned staging tensors (asynchronous copies); arrays
        # assigned from outside — a restored checkpoint, a test — are uploaded just the same
        out = {}
        for k in self.KEYS:
            arr, h = self.buffer[k], self._host.get(k)
            if h is not None and arr.shape == tuple(h.shape) and arr.__array_interface__["data"][0] == h.data_ptr():
                out[k] = h.to(self.device, non_blocking=True)       # the pinned tensor itself: its allocator tracks the copy in flight
            else:
                out[k] = torch.from_numpy(np.ascontiguousarray(arr)).to(self.device)
        return out


class MAPPO:
    max_norm_joint = 10.0        # clip_grad_norm_(ac_parameters, 10)   MAPPO_discrete.py:191

    def __init__(self, dim_info, is_continue, actor_lr, critic_lr, horizon, device, trick=None, buffer=None):
        self.device = _lib.require_device(device)
        if is_continue:
            raise NotImplementedError("MAPPO_discrete.learn is Categorical-only upstream (MAPPO_discrete.py:337); use MAPPO.py for Gaussian actors")
        if len({tuple(v) for v in dim_info.values()}) != 1:
            raise ValueError("MAPPO_discrete shares one actor: every agent needs the same (obs_dim, action_dim)")
        self.agent_x = list(dim_info.keys())[0]
        obs_dim, action_dim = dim_info[self.agent_x]
        self.obs_dim, self.action_dim = obs_dim, action_dim
        self.N = len(dim_info)
        self.agent = _Agent(obs_dim, action_dim, dim_info, actor_lr, critic_lr, is_continue, self.device, trick)
        self.buffer = buffer
        self.batch_size = buffer.batch_size
        self.episode_limit = buffer.episode_limit
        self.is_continue = is_continue
        print('actor_type:continue') if self.is_continue else print('actor_type:discrete')
        self.horizon = int(horizon)
        self.trick = trick
        self.actor_lr, self.critic_lr = actor_lr, critic_lr
        ln, fn = bool(trick['LayerNorm']), bool(trick['feature_norm'])
        # acting networks normalise per row (frl_infer_args_t.layer_norm: 1 input + hidden, 2 hidden only, 3 input only)
        self._ln_actor = 2 if ln else 0
        self._ln_critic = (1 if fn else 2) if ln else (3 if fn else 0)
        self._group_norm = (1 if ln else 0) | (2 if fn else 0)
        self._scalar_vloss = bool(trick['ValueClip'] and trick['huber_loss'])
        self.mode = _common.resolve_mode(None)
        self._seed = _common.default_seed()
        self._n_act = 0
        sm = _lib.sm_count()
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=self.device)
        self._gpart, self._sumsq = z(sm, self.agent._net.n_p), z(sm, 2)
        self._segcnt, self._stats = z(sm, _lib.NSEG), z(sm, 8)
        self.last_metrics = None

    # ---- acting ------------------------------------------------------------------------------------
    def select_action(self, obs, *, noise=None):
        """``obs``: N rows of ``obs_dim`` -> (actions int64 [N], log-probs float32 [N]); one ``Categorical.sample()`` over the [N, A] block
        (``:236-239``; torch.multinomial draws q ~ Exp(1) per class and takes argmax(p / q))."""
        x = np.asarray(obs, dtype=np.float32).reshape(-1, self.obs_dim)
        n = x.shape[0]
        self._n_act += 1
        if noise is not None:
            noise = torch.as_tensor(noise, dtype=torch.float32).to(self.device).reshape(n, self.action_dim).contiguous()
        elif self.mode == "parity":
            noise = torch.empty((n, self.action_dim), dtype=torch.float32, device=self.device).exponential_(1)
        out = _common.infer(self.agent._net, x, _lib.INFER_PPO_CAT, self.device, 2, noise=noise, seed=self._seed, counter=self._n_act,
                            l0=0, nl=3, layer_norm=self._ln_actor).cpu().numpy()
        return out[:, 0].astype(np.int64), out[:, 1]

    def evaluate_action(self, obs):
        x = np.asarray(obs, dtype=np.float32).reshape(-1, self.obs_dim)
        return _common.infer(self.agent._net, x, _lib.INFER_ARGMAX, self.device, 1, l0=0, nl=3,
                             layer_norm=self._ln_actor).reshape(-1).to(torch.int64).cpu().numpy()

    # ---- buffer ------------------------------------------------------------------------------------
    def get_value(self, s):
        """every agent sees the same joint state, so the N critic rows of ``:252-260`` are one value repeated"""
        x = np.asarray(s, dtype=np.float32).reshape(1, -1)
        v = _common.infer(self.agent._net, x, _lib.INFER_RAW, self.device, 1, l0=3, nl=3, layer_norm=self._ln_critic).cpu().numpy().reshape(-1)
        return np.repeat(v, self.N)

    def add(self, obs, action, reward, next_obs, done, action_log_pi, adv_dones, episode_step):
        """``:270-283``; ``next_obs`` and ``adv_dones`` are unused upstream too (interface compatibility)."""
        s = np.asarray(obs, dtype=np.float32).flatten()
        v_n = self.get_value(s)
        r_n = [v for v in reward.values()]
        done_n = [d for d in done.values()]
        self.buffer.store_transition(episode_step, obs, s, v_n, action, action_log_pi, r_n, done_n)

    # ---- learning ----------------------------------------------------------------------------------
    def _time_major(self, batch):
        """device tensors of ``get_training_data`` ([B, T, N, ...]) -> time-major row blocks, rows r = (t * B + b) * N + n"""
        B, T, N = self.batch_size, self.episode_limit, self.N
        tm = lambda x: x.transpose(0, 1).contiguous()
        obs = tm(batch["obs_n"]).reshape(T * B * N, self.obs_dim)
        s = tm(batch["s"]).unsqueeze(2).expand(T, B, N, batch["s"].shape[-1]).reshape(T * B * N, -1).contiguous()      # :282 repeat over the agents
        v = tm(batch["v_n"]).reshape(T + 1, B * N)
        act = tm(batch["a_n"]).to(torch.float32).reshape(T * B * N, 1).contiguous()
        logp = tm(batch["a_logprob_n"]).reshape(T * B * N, 1)
        rew, done = tm(batch["r_n"]).reshape(T, B * N), tm(batch["done_n"]).reshape(T, B * N)
        return obs, s, v, act, logp, rew, done

    def compute_advantages(self, v, rew, done, gamma, lmbda):
        T, cols = rew.shape
        adv = torch.empty((T * cols, 1), dtype=torch.float32, device=self.device)
        vt = torch.empty((T * cols, 1), dtype=torch.float32, device=self.device)
        zeros = torch.zeros_like(rew)
        vs, vs_next = v[:T].contiguous(), v[1:].contiguous()
        _lib.check(_lib.lib().frl_gae(_lib.ptr(rew), _lib.ptr(done), _lib.ptr(zeros), _lib.ptr(vs), _lib.ptr(vs_next), T, cols,
                                      float(gamma), float(lmbda), _lib.ptr(adv), _lib.ptr(vt), _lib.stream_ptr(self.device)), "frl_gae")
        if self.trick['adv_norm']:
            _lib.check(_lib.lib().frl_adv_norm(_lib.ptr(adv), adv.numel(), 1e-8, _lib.ptr(adv), _lib.stream_ptr(self.device)), "frl_adv_norm")
        return adv, vt, vs

    def learn(self, minibatch_size, gamma, lmbda, clip_param, K_epochs, entropy_coefficient, huber_delta=None):
        B, T, N = self.batch_size, self.episode_limit, self.N
        obs, s, v, act, logp, rew, done = self._time_major(self.buffer.get_training_data())
        adv, v_target, v_old = self.compute_advantages(v, rew, done, gamma, lmbda)
        self.last_adv, self.last_v_target = adv, v_target
        # one epoch's plan: consecutive blocks of minibatch_size episodes; rows of episode b are {(t * B + b) * N + n}, listed episode by
        # episode (the T * N rows of an episode are one LayerNorm group of the kernel's group mode)
        nmb = (B + minibatch_size - 1) // minibatch_size
        mb = minibatch_size * T * N
        ar = torch.arange
        base = (ar(T, device=self.device).view(1, T, 1) * B) * N + ar(N, device=self.device).view(1, 1, N)          # + b * N
        idx = torch.zeros((nmb, mb), dtype=torch.int64, device=self.device)
        rows = torch.zeros(nmb, dtype=torch.int32, device=self.device)
        for j in range(nmb):
            eps = ar(j * minibatch_size, min((j + 1) * minibatch_size, B), device=self.device)
            r = (base + eps.view(-1, 1, 1) * N).reshape(-1)
            idx[j, :r.numel()] = r
            rows[j] = r.numel()
        idx_d, rows_d = idx.repeat(K_epochs, 1).contiguous(), rows.repeat(K_epochs).contiguous()
        n_updates = K_epochs * nmb
        out = torch.zeros((n_updates, 8), dtype=torch.float32, device=self.device)
        ag = self.agent
        a = _lib.PpoArgs()
        a.net, a.continuous = ag._net.c_struct(), 0
        a.obs, a.action, a.logp_old = obs.data_ptr(), act.data_ptr(), logp.data_ptr()
        a.adv, a.v_target = adv.data_ptr(), v_target.data_ptr()
        a.M, a.obs_dim, a.act_cols, a.logp_cols, a.n_adv = T * B * N, self.obs_dim, 1, 1, 1
        a.indices, a.mb_rows, a.mb, a.n_updates = idx_d.data_ptr(), rows_d.data_ptr(), mb, n_updates
        a.clip_param, a.entropy_coef = clip_param, entropy_coefficient
        a.max_norm_actor = a.max_norm_critic = 0.0
        a.max_norm_joint, a.opt_repeat = self.max_norm_joint, 2
        a.optimizer = _lib.OPT_ADAM
        a.lr, a.beta1, a.beta2, a.eps = ag.lr, 0.9, 0.999, (1e-5 if self.trick['adam_eps'] else 1e-8)
        a.step0 = ag.step
        a.critic_obs, a.critic_obs_dim = s.data_ptr(), s.shape[1]
        if self.trick['ValueClip']:
            a.value_loss, a.v_old = (3 if self._scalar_vloss else 2), v_old.data_ptr()
            a.huber_delta = float(huber_delta) if huber_delta is not None else 0.0
        a.gpart, a.sumsq, a.segcnt = self._gpart.data_ptr(), self._sumsq.data_ptr(), self._segcnt.data_ptr()
        a.stats, a.out = self._stats.data_ptr(), out.data_ptr()
        if self._group_norm or self._scalar_vloss:
            a.group_rows, a.group_norm = T * N, self._group_norm
        else:
            a.umma_ws = _common.umma_ws_ptr(self.device, int(mb))
        launch = lambda: _lib.check(_lib.lib().frl_ppo_update(ctypes.byref(a), _lib.stream_ptr(self.device)), "frl_ppo_update")
        if not self._scalar_vloss:
            launch()                              # every minibatch of every epoch in one persistent launch
        else:
            # the scalar value loss needs the whole minibatch's critic outputs before any gradient: per update, a forward-only
            # pre-pass launch (per-CTA huber sums) and the update launch that folds them
            idx0, rows0, out0, step0 = a.indices, a.mb_rows, a.out, a.step0
            a.n_updates = 1
            for u in range(n_updates):
                a.indices, a.mb_rows, a.out, a.step0 = idx0 + u * mb * 8, rows0 + u * 4, out0 + u * 8 * 4, step0 + 2 * u
                a.group_prepass, a.stage_lo, a.stage_hi = 1, 0, 1
                launch()
                a.group_prepass, a.stage_lo, a.stage_hi = 0, 0, 0
                launch()
        ag.step += 2 * n_updates
        self._keep = (obs, s, v, act, logp, rew, done, adv, v_target, v_old, idx_d, rows_d)
        self.last_metrics = out
        self.buffer.reset_buffer()

    def lr_decay(self, episode_num, max_episodes):
        self.agent.lr = self.actor_lr * (1 - episode_num / max_episodes)                  # :373-387 (the critic rate is commented out upstream)

    def save(self, model_dir):
        torch.save({k: v.detach().clone().cpu() for k, v in self.agent.actor.state_dict().items()}, os.path.join(model_dir, 'MAPPO_discrete.pth'))

    @staticmethod
    def load(dim_info, is_continue, model_dir, trick=None, device=None, buffer=None):
        """the reference's ``load`` builds ``MAPPO(...)`` without a buffer and fails on ``buffer.batch_size`` (``:214``); a one-episode
        buffer is supplied here so that a saved actor can be evaluated"""
        device = device if device is not None else torch.device("cuda")
        obs_dim = list(dim_info.values())[0][0]
        if buffer is None:
            buffer = ReplayBuffer(len(dim_info), obs_dim, obs_dim * len(dim_info), 1, 1, device)
        policy = MAPPO(dim_info, is_continue=is_continue, actor_lr=0, critic_lr=0, horizon=0, device=device, trick=trick, buffer=buffer)
        policy.agent.actor.load_state_dict(torch.load(os.path.join(model_dir, 'MAPPO_discrete.pth'), map_location=device))
        return policy
