"""DDPG with the reference's class API (``DDPG_file/DDPG.py:58-245``) on the fused B200 kernel.

Supplements kept: ``weight_decay`` (critic Adam with L2 1e-3, DDPG.py:131-134) and ``net_init`` (uniform
re-initialisation, note ``fan_in = weight.size(0)`` = out-features, DDPG.py:58-68).  ``Batch_ObsNorm`` (the default supplement, DDPG.py:446) runs inside the kernel (``freerl_b200/normalization.py``).
"""
import os

import torch
import torch.nn as nn

from . import _common, _lib
from ._actor_critic import ACBase


def _reference_net_init(module, names):
    """``other_net_init`` on the two hidden layers, ``final_net_init`` (+-3e-3) on the last (DDPG.py:58-68,78-84)."""
    for name in names:
        layer = getattr(module, name)
        if name in (names[-1],) or name in ("l3", "l6") and len(names) <= 3:
            nn.init.uniform_(layer.weight, -3e-3, 3e-3)
            nn.init.uniform_(layer.bias, -3e-3, 3e-3)
        else:
            limit = 1.0 / (layer.weight.data.size(0) ** 0.5)
            nn.init.uniform_(layer.weight, -limit, limit)
            nn.init.uniform_(layer.bias, -limit, limit)


class DDPG(ACBase):
    n_heads = 1
    sac = False

    def __init__(self, dim_info, is_continue, actor_lr, critic_lr, buffer_size, device, trick=None, supplement=None, mode=None):
        self.trick = trick
        self.supplement = supplement if supplement is not None else {
            "weight_decay": False, "OUNoise": False, "ObsNorm": False, "net_init": False, "Batch_ObsNorm": False}
        post = _reference_net_init if self.supplement.get("net_init") else None
        self._setup(dim_info, is_continue, actor_lr, critic_lr, buffer_size, device, mode, post_init=post,
                    batch_obs_norm=self.supplement.get("Batch_ObsNorm", False))

    def select_action(self, obs):
        x, single = _common.as_obs_batch(obs, self.obs_dim)
        a = _common.infer(self.agent._actor, x, _lib.INFER_TANH, self.device, self.action_dim, obs_norm=self._obs_norm()).cpu().numpy()
        return a[0] if single else a

    def evaluate_action(self, obs):
        """The reference's evaluate_action does NOT apply Batch_ObsNorm (DDPG.py:175-183); kept."""
        x, single = _common.as_obs_batch(obs, self.obs_dim)
        a = _common.infer(self.agent._actor, x, _lib.INFER_TANH, self.device, self.action_dim).cpu().numpy()
        return a[0] if single else a

    def learn(self, batch_size, gamma, tau, *, n_updates=1, indices=None):
        a, idx, B, out = self._base_args(batch_size, gamma, tau, n_updates, indices)
        a.wd_critic = 1e-3 if self.supplement.get("weight_decay") else 0.0
        self._launch(a, (idx,), n_updates, out)
        self.agent.critic_step += n_updates
        self.agent.actor_step += n_updates
        self._n_learn += n_updates

    def save(self, model_dir):
        self._save_actor(os.path.join(model_dir, "DDPG.pt"))

    @staticmethod
    def load(dim_info, is_continue, model_dir, trick=None, supplement=None, device=None):
        device = device if device is not None else torch.device("cuda")
        policy = DDPG(dim_info, is_continue, 0, 0, 0, device=device, trick=trick, supplement=supplement)
        policy.agent.actor.load_state_dict(torch.load(os.path.join(model_dir, "DDPG.pt"), map_location=device))
        return policy
