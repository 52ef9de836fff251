"""TD3 with the reference's class API (``TD3_file/TD3.py:52-256``) on the fused B200 kernel.

``realize`` switches: ``clip_double`` (twin critic, min target, actor on Q1), ``policy_noise`` (target smoothing),
``twin_delay`` (actor + Polyak every ``policy_freq`` learns).  ``learn(batch_size, gamma, tau, policy_noise,
noise_clip, max_action, policy_freq, policy_noise_scale)`` keeps the reference's positional signature.
"""
import os

import torch

from . import _common, _lib
from ._actor_critic import ACBase


class TD3(ACBase):
    sac = False

    def __init__(self, dim_info, is_continue, actor_lr, critic_lr, buffer_size, device, trick=None, realize=None, mode=None):
        self.trick = trick
        self.realize = realize if realize is not None else {"clip_double": True, "policy_noise": True, "twin_delay": True}
        self.n_heads = 2 if self.realize["clip_double"] else 1
        self._setup(dim_info, is_continue, actor_lr, critic_lr, buffer_size, device, mode)
        self.total_it = 0

    def select_action(self, obs):
        x, single = _common.as_obs_batch(obs, self.obs_dim)
        a = _common.infer(self.agent._actor, x, _lib.INFER_TANH, self.device, self.action_dim).cpu().numpy()
        return a[0] if single else a

    def evaluate_action(self, obs):
        return self.select_action(obs)

    def learn(self, batch_size, gamma, tau, policy_noise, noise_clip, max_action, policy_freq, policy_noise_scale, *,
              n_updates=1, indices=None, noise=None):
        a, idx, B, out = self._base_args(batch_size, gamma, tau, n_updates, indices)
        a.total_it0 = self.total_it
        smoothing = bool(self.realize["policy_noise"])
        if smoothing and self.mode == "parity" and noise is None:
            noise = torch.stack([_common.reference_randn((B, self.action_dim), self.device) for _ in range(n_updates)])  # randn_like (TD3.py:197)
        nz = self._noise(noise, n_updates, B) if smoothing else None
        a.noise_next = nz.data_ptr() if nz is not None else None
        a.target_smoothing = int(smoothing)
        a.policy_noise, a.noise_clip, a.max_action = policy_noise, noise_clip, max_action
        a.policy_noise_scale = policy_noise_scale
        pf = policy_freq if self.realize["twin_delay"] else 1
        a.policy_freq = pf
        self._launch(a, (idx, nz), n_updates, out)
        n_policy = (self.total_it + n_updates) // pf - self.total_it // pf
        self.total_it += n_updates
        self.agent.critic_step += n_updates
        self.agent.actor_step += n_policy
        self._n_learn += n_updates

    def save(self, model_dir):
        self._save_actor(os.path.join(model_dir, "TD3.pt"))

    @staticmethod
    def load(dim_info, is_continue, model_dir, trick=None, realize=None, device=None):
        device = device if device is not None else torch.device("cuda")
        policy = TD3(dim_info, is_continue, 0, 0, 0, device=device, trick=trick, realize=realize)
        policy.agent.actor.load_state_dict(torch.load(os.path.join(model_dir, "TD3.pt"), map_location=device))
        return policy
