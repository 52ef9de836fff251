"""MADDPG with the reference's class API (``MADDPG_file/MADDPG.py:52-253``) on the fused B200 actor-critic kernel.

``MADDPG(dim_info: dict, is_continue, actor_lr, critic_lr, buffer_size, device, trick, supplement)``;
``select_action(obs: dict) -> dict``, ``evaluate_action``, ``add`` (dicts keyed by agent id), ``sample``,
``learn(batch_size, gamma, tau)``, ``update_target``, ``save``/``load``.  Reference behaviour kept: one replay per
agent sharing the sampled indices; a FRESH sample for every agent inside the learn loop (:211); centralised critic on
``cat(all obs, all actions)``; next actions from every agent's TARGET actor; critic Adam with L2 weight_decay 1e-3
and the uniform ``net_init`` when the supplements ask; Polyak of all agents only after the loop.
``Batch_ObsNorm`` (default supplement, MADDPG.py:437): per-agent statistics updated by EVERY agent's sample, in-kernel.
"""
import ctypes
import os

import numpy as np
import torch

from . import _common, _lib
from .Buffer import Buffer
from .DDPG import _reference_net_init
from ._actor_critic import _ActorInit, _CriticInit
from .nets import DeviceNet, alias_module, bind_module
from .normalization import Normalization_batch_size


class Agent:
    """``MADDPG.py:112-136``: per-agent actor (own obs) + centralised critic (all obs + all actions), targets, 2 Adams."""

    def __init__(self, obs_dim, action_dim, dim_info, actor_lr, critic_lr, device, supplement, n_heads=1):
        joint = sum(sum(v) for v in dim_info.values())
        a_dims = [(obs_dim, 128), (128, 128), (128, action_dim)]
        c_dims = [(joint, 128), (128, 128), (128, 1)] * n_heads         # twin: Critic_TD3 l1-l3 / l4-l6 (MATD3_simple.py:87-115)
        names = ("l1", "l2", "l3")
        c_names = names if n_heads == 1 else ("l1", "l2", "l3", "l4", "l5", "l6")
        self._actor, self._critic = DeviceNet(a_dims, device, True), DeviceNet(c_dims, device, True)
        self._actor_t, self._critic_t = DeviceNet(a_dims, device, False), DeviceNet(c_dims, device, False)
        a_init = _ActorInit(obs_dim, action_dim, "l3")
        if supplement['net_init']:
            _reference_net_init(a_init, names)
        c_init = _CriticInit(joint, n_heads)
        if supplement['net_init']:
            _reference_net_init(c_init, c_names)
        self.actor = bind_module(self._actor, a_init, names)
        self.critic = bind_module(self._critic, c_init, c_names)
        self._actor_t.copy_from(self._actor)
        self._critic_t.copy_from(self._critic)
        self.actor_target = alias_module(self._actor_t, names)
        self.critic_target = alias_module(self._critic_t, c_names)
        self.actor_lr, self.critic_lr = actor_lr, critic_lr
        self.weight_decay = 1e-3 if supplement['weight_decay'] else 0.0
        self.actor_step = self.critic_step = 0


class MADDPG:
    n_heads = 1          # MATD3 (clip_double): 2

    def __init__(self, dim_info, is_continue, actor_lr, critic_lr, buffer_size, device, trick, supplement, mode=None):
        self.device = _lib.require_device(device)
        if len(dim_info) > _lib.FRL_MAX_AGENTS:
            raise NotImplementedError("at most %d agents" % _lib.FRL_MAX_AGENTS)
        if not is_continue:
            raise NotImplementedError("the reference implements continuous actions only (MADDPG.py:166)")
        self.agents, self.buffers = {}, {}
        for agent_id, (obs_dim, action_dim) in dim_info.items():
            self.agents[agent_id] = Agent(obs_dim, action_dim, dim_info, actor_lr, critic_lr, self.device, supplement, self.n_heads)
            self.buffers[agent_id] = Buffer(buffer_size, obs_dim, act_dim=action_dim if is_continue else 1, device=self.device)
        self.dim_info = dim_info
        self.is_continue = is_continue
        self._bon = bool(supplement.get('Batch_ObsNorm'))
        if self._bon:
            self.batch_size_obs_norm = {agent_id: Normalization_batch_size(shape=dim_info[agent_id][0], device=self.device)
                                        for agent_id in dim_info.keys()}
        self.agent_x = list(self.agents.keys())[0]
        self.regular = False
        self.supplement = supplement
        self.mode = _common.resolve_mode(mode)
        self._seed = _common.default_seed()
        self._n_learn = 0
        n_p = max(max(a._actor.n_p, a._critic.n_p) for a in self.agents.values())
        self._scratch = _common.DeviceScratch(self.device, n_p)
        self.last_metrics = None

    def select_action(self, obs, evaluate=False):
        actions = {}
        for agent_id, o in obs.items():
            od, ad = self.dim_info[agent_id]
            x, single = _common.as_obs_batch(o, od)
            norm = self.batch_size_obs_norm[agent_id] if (self._bon and not evaluate) else None   # MADDPG.py:162-163 vs :176-180
            a = _common.infer(self.agents[agent_id]._actor, x, _lib.INFER_TANH, self.device, ad, obs_norm=norm).cpu().numpy()
            actions[agent_id] = a[0] if single else a
        return actions

    def evaluate_action(self, obs):
        return self.select_action(obs, evaluate=True)

    def add(self, obs, action, reward, next_obs, done):
        for agent_id, buffer in self.buffers.items():
            buffer.add(obs[agent_id], action[agent_id], reward[agent_id], next_obs[agent_id], done[agent_id])

    def sample(self, batch_size):
        total_size = len(self.buffers[self.agent_x])
        indices = np.random.choice(total_size, batch_size, replace=False)
        obs, action, reward, next_obs, done, next_action = {}, {}, {}, {}, {}, {}
        for agent_id, buffer in self.buffers.items():
            obs[agent_id], action[agent_id], reward[agent_id], next_obs[agent_id], done[agent_id] = buffer.sample(indices)
            od, ad = self.dim_info[agent_id]
            if self._bon:
                obs[agent_id] = self.batch_size_obs_norm[agent_id](obs[agent_id], update=True)
                next_obs[agent_id] = self.batch_size_obs_norm[agent_id](next_obs[agent_id], update=False)
            next_action[agent_id] = _common.infer(self.agents[agent_id]._actor_t, next_obs[agent_id], _lib.INFER_TANH, self.device, ad)
        return obs, action, reward, next_obs, done, next_action

    def learn(self, batch_size, gamma, tau, *, indices=None):
        """``indices``: optional list (one per agent, agent order) of index arrays overriding the per-agent fresh samples."""
        self._learn(batch_size, gamma, tau, indices, None)

    def _learn(self, batch_size, gamma, tau, indices, td3):
        """One learn() over all agents.  ``td3`` (MATD3 only): dict(policy_step, policy_freq, total_it, smoothing, policy_noise,
        noise_clip, max_action, policy_noise_scale, noise) with ``noise[i][j]`` = device randn [B, act_j] of agent j inside agent i's sample."""
        ids = list(self.agents.keys())
        total = len(self.buffers[self.agent_x])
        outs = []
        for i, agent_id in enumerate(ids):
            ag = self.agents[agent_id]
            if indices is None:
                idx = _common.make_indices(self.mode, total, batch_size, 1, self.device, self._seed, self._n_learn * len(ids) + i)
            else:
                idx = self.buffers[agent_id]._indices_to_device(indices[i]).reshape(1, -1)
            B = idx.shape[1]
            a = _lib.AcArgs()
            a.actor, a.actor_target = ag._actor.c_struct(), ag._actor_t.c_struct()
            a.critic, a.critic_target = ag._critic.c_struct(), ag._critic_t.c_struct()
            a.n_heads, a.actor_kind = self.n_heads, _lib.ACTOR_TANH
            a.replay = self.buffers[agent_id].c_struct()
            a.indices, a.B, a.n_updates = idx.data_ptr(), B, 1
            a.seed, a.gamma, a.tau = self._seed, gamma, tau
            a.lr_actor, a.lr_critic = ag.actor_lr, ag.critic_lr
            a.beta1, a.beta2, a.eps, a.wd_critic, a.max_norm = 0.9, 0.999, 1e-8, ag.weight_decay, 0.5
            a.step_actor0, a.step_critic0, a.total_it0 = ag.actor_step, ag.critic_step, self._n_learn
            a.policy_freq, a.target_smoothing, a.max_action, a.policy_noise_scale = 1, 0, 1.0, 1.0
            policy_step = True
            if td3 is not None:
                # the kernel runs the actor stages when (total_it0 + 1) % policy_freq == 0 (total_it counts learn() calls)
                policy_step = bool(td3["policy_step"])
                a.policy_freq, a.total_it0 = td3["policy_freq"], td3["total_it"] - 1
                a.target_smoothing = int(td3["smoothing"])
                a.policy_noise, a.noise_clip = td3["policy_noise"], td3["noise_clip"]
                a.max_action, a.policy_noise_scale = td3["max_action"], td3["policy_noise_scale"]
                if td3["noise"] is not None:
                    for j in range(len(ids)):
                        a.ma_noise_next[j] = td3["noise"][i][j].data_ptr()
            out = torch.zeros((1, 8), dtype=torch.float32, device=self.device)
            a.gpart, a.sumsq = self._scratch.gpart.data_ptr(), self._scratch.sumsq.data_ptr()
            a.stats, a.out = self._scratch.stats.data_ptr(), out.data_ptr()
            a.n_agents, a.agent_index, a.defer_polyak = len(ids), i, 1
            a.xchg = self._scratch.xchg(B, self.device).data_ptr()
            for j, other in enumerate(ids):
                a.ma_replay[j] = self.buffers[other].c_struct()
                a.ma_actor_target[j] = self.agents[other]._actor_t.c_struct()
                if self._bon:
                    a.obs_norm[j] = self.batch_size_obs_norm[other].data_ptr()
            if self._bon:
                a.obs_norm_n0 = self.batch_size_obs_norm[self.agent_x].running_ms.n
            _lib.check(_lib.lib().frl_ac_learn(ctypes.byref(a), _lib.stream_ptr(self.device)), "frl_ac_learn")
            ag.actor_step += 1 if policy_step else 0
            ag.critic_step += 1
            if self._bon:                     # every agent's sample() updates every agent's statistics (MADDPG.py:192-196)
                for nm in self.batch_size_obs_norm.values():
                    nm.running_ms.n += 1
            outs.append((out, idx))
        if td3 is None or td3["policy_step"]:
            self.update_target(tau)
        self._n_learn += 1
        self.last_metrics = torch.cat([o for o, _ in outs])

    def update_target(self, tau):
        for ag in self.agents.values():
            for src, tgt in ((ag._actor, ag._actor_t), (ag._critic, ag._critic_t)):
                _lib.check(_lib.lib().frl_polyak(ctypes.byref(src.c_struct()), ctypes.byref(tgt.c_struct()), float(tau),
                                                 _lib.stream_ptr(self.device)), "frl_polyak")

    def save(self, model_path):
        torch.save({name: {k: v.detach().clone().cpu() for k, v in agent.actor.state_dict().items()} for name, agent in self.agents.items()},
                   os.path.join(model_path, 'MADDPG.pth'))

    @staticmethod
    def load(dim_info, is_continue, model_dir, trick=None, supplement=None, device=None):
        device = device if device is not None else torch.device("cuda")
        policy = MADDPG(dim_info, is_continue=is_continue, actor_lr=0, critic_lr=0, buffer_size=0, device=device, trick=trick, supplement=supplement)
        data = torch.load(os.path.join(model_dir, 'MADDPG.pth'), map_location=device)
        for agent_id, agent in policy.agents.items():
            agent.actor.load_state_dict(data[agent_id])
        return policy
