"""MATD3 with the reference's class API (``MADDPG_file/MATD3_simple.py:151-275``) on the fused multi-agent actor-critic kernel.

``MATD3(dim_info: dict, is_continue, actor_lr, critic_lr, buffer_size, device, trick=None, realize=None)``;
``learn(batch_size, gamma, tau, policy_noise_scale, policy_noise, noise_clip, max_action, policy_freq)`` (the reference's
positional order, ``:217``).  ``realize = {'clip_double', 'policy_noise', 'twin_delay'}``: twin centralised critics with
``min(Q1', Q2')`` targets and ``Q1`` in the actor loss; target policy smoothing on EVERY agent's next action (one
``randn_like`` per agent per sample, ``:203-205``); actors and all targets move only when ``total_it % policy_freq == 0``
where ``total_it`` counts ``learn()`` calls.  As in MADDPG every agent draws a FRESH sample inside the agent loop (``:226``).
"""
import os

import torch

from . import _common
from .MADDPG import MADDPG as _MADDPG
from .MADDPG_simple import _OFF


class MATD3(_MADDPG):
    def __init__(self, dim_info, is_continue, actor_lr, critic_lr, buffer_size, device, trick=None, realize=None, mode=None):
        self.realize = realize if realize is not None else {'clip_double': True, 'policy_noise': True, 'twin_delay': True}
        self.n_heads = 2 if self.realize['clip_double'] else 1
        super().__init__(dim_info, is_continue, actor_lr, critic_lr, buffer_size, device, trick, dict(_OFF), mode=mode)
        self.total_it = 0

    def learn(self, batch_size, gamma, tau, policy_noise_scale, policy_noise, noise_clip, max_action, policy_freq, *,
              indices=None, noise=None):
        """``noise``: optional ``noise[i][j]`` = randn [B, act_j] for agent j's target action inside agent i's sample."""
        self.total_it += 1
        pf = int(policy_freq) if self.realize['twin_delay'] else 1
        ids = list(self.agents.keys())
        smoothing = bool(self.realize['policy_noise'])
        dev_noise = None
        if smoothing and (noise is not None or self.mode == "parity"):
            B = min(batch_size, len(self.buffers[self.agent_x])) if indices is None else len(indices[0])
            dev_noise = []
            for i in range(len(ids)):
                row = []
                for j, k in enumerate(ids):
                    ad = self.dim_info[k][1]
                    if noise is not None:
                        row.append(torch.as_tensor(noise[i][j], dtype=torch.float32).to(self.device).reshape(B, ad).contiguous())
                    else:
                        row.append(_common.reference_randn((B, ad), self.device))        # torch.randn_like(action_j)
                dev_noise.append(row)
        self._learn(batch_size, gamma, tau, indices, dict(
            policy_step=(self.total_it % pf == 0), policy_freq=pf, total_it=self.total_it, smoothing=smoothing,
            policy_noise=float(policy_noise), noise_clip=float(noise_clip), max_action=float(max_action),
            policy_noise_scale=float(policy_noise_scale), noise=dev_noise))
        self._keep_noise = dev_noise

    # save(): inherited — the reference writes 'MADDPG.pth' here too (MATD3_simple.py:268-272)

    @staticmethod
    def load(dim_info, is_continue, model_dir, trick=None, realize=None, device=None):
        device = device if device is not None else torch.device("cuda")
        policy = MATD3(dim_info, is_continue, 0, 0, 0, device, trick=trick, realize=realize)
        data = torch.load(os.path.join(model_dir, 'MADDPG.pth'), map_location=device)
        for agent_id, agent in policy.agents.items():
            agent.actor.load_state_dict(data[agent_id])
        return policy
