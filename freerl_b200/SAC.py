"""SAC with the reference's class API (``SAC_file/SAC.py:60-282``) on the fused B200 kernel.

``SAC(dim_info, is_continue, actor_lr, critic_lr, buffer_size, device, trick=None)``; ``select_action`` (stochastic
tanh-Gaussian), ``evaluate_action`` (tanh(mean)), ``add``, ``sample``, ``learn(batch_size, gamma, tau)``,
``update_target``, ``save``/``load``.  Reference quirks kept: the target action comes from an ACTOR TARGET
(SAC.py:227), ``log_std`` is a state-independent parameter clamped to [-20, 2], alpha starts at 0.01 with its own
Adam (lr 1e-4) on ``log_alpha``, both optimisers clip the global grad norm at 0.5, Polyak on critic and actor.
"""
import os

import numpy as np
import torch

from . import _common, _lib
from ._actor_critic import ACBase


class Alpha:
    """``SAC.py:154-169``.  State lives on the device: ``state = [log_alpha, exp_avg, exp_avg_sq, 0]``."""

    def __init__(self, action_dim, device, alpha_lr=0.0001, alpha=0.2, requires_grad=False, is_continue=True):
        self.state = torch.zeros(4, dtype=torch.float32, device=device)
        self.state[0] = float(np.float32(np.log(alpha)))
        self.alpha_lr = alpha_lr
        self.requires_grad = requires_grad
        if is_continue:
            self.target_entropy = -action_dim
        else:
            self.target_entropy = float(0.6 * (-np.log(1.0 / action_dim)))
        self.step = 0

    @property
    def log_alpha(self):
        return self.state[0]

    @property
    def alpha(self):
        return self.state[0].exp()


class SAC(ACBase):
    n_heads = 2
    sac = True

    def __init__(self, dim_info, is_continue, actor_lr, critic_lr, buffer_size, device, trick=None, mode=None):
        self.trick = trick if trick is not None else {}
        self._setup(dim_info, is_continue, actor_lr, critic_lr, buffer_size, device, mode,
                    batch_obs_norm=self.trick.get("Batch_ObsNorm", False))
        self.adaptive_alpha = True
        print('adaptive_alpha:', self.adaptive_alpha)
        self.alphas = Alpha(self.action_dim, self.device, alpha=0.01, requires_grad=self.adaptive_alpha, is_continue=is_continue)

    # ---- acting ------------------------------------------------------------------------------------
    def select_action(self, obs, *, noise=None):
        x, single = _common.as_obs_batch(obs, self.obs_dim)
        n = x.shape[0]
        if noise is None and self.mode == "parity":
            noise = _common.reference_randn((n, self.action_dim), self.device)       # dist.rsample() (SAC.py:82)
        elif noise is not None:
            noise = torch.as_tensor(noise, dtype=torch.float32).to(self.device).reshape(n, self.action_dim).contiguous()
        self._n_act += 1
        a = _common.infer(self.agent._actor, x, _lib.INFER_SAC_SAMPLE, self.device, self.action_dim, noise=noise,
                          seed=self._seed, counter=self._n_act, obs_norm=self._obs_norm()).cpu().numpy()
        return a[0] if single else a

    def evaluate_action(self, obs):
        x, single = _common.as_obs_batch(obs, self.obs_dim)
        a = _common.infer(self.agent._actor, x, _lib.INFER_SAC_MEAN, self.device, self.action_dim).cpu().numpy()
        return a[0] if single else a

    # ---- learning ----------------------------------------------------------------------------------
    def learn(self, batch_size, gamma, tau, *, n_updates=1, indices=None, noise_next=None, noise_new=None):
        a, idx, B, out = self._base_args(batch_size, gamma, tau, n_updates, indices)
        if self.mode == "parity" and noise_next is None and noise_new is None:
            # reference order per learn(): eps for a' (actor_target rsample) then eps for the new action
            nn_, nw_ = [], []
            for _ in range(n_updates):
                nn_.append(_common.reference_randn((B, self.action_dim), self.device))
                nw_.append(_common.reference_randn((B, self.action_dim), self.device))
            noise_next, noise_new = torch.stack(nn_), torch.stack(nw_)
        nx, nw = self._noise(noise_next, n_updates, B), self._noise(noise_new, n_updates, B)
        a.noise_next = nx.data_ptr() if nx is not None else None
        a.noise_new = nw.data_ptr() if nw is not None else None
        al = self.alphas
        a.alpha_state, a.adaptive_alpha = al.state.data_ptr(), int(self.adaptive_alpha)
        a.alpha_lr, a.target_entropy, a.step_alpha0 = al.alpha_lr, float(al.target_entropy), al.step
        self._launch(a, (idx, nx, nw), n_updates, out)
        self.agent.actor_step += n_updates
        self.agent.critic_step += n_updates
        al.step += n_updates
        self._n_learn += n_updates

    # ---- checkpoint ---------------------------------------------------------------------------------
    def save(self, model_dir):
        self._save_actor(os.path.join(model_dir, "SAC.pt"))

    @staticmethod
    def load(dim_info, is_continue, model_dir, trick=None, device=None):
        device = device if device is not None else torch.device("cuda")
        policy = SAC(dim_info, is_continue, 0, 0, 0, device=device, trick=trick)
        policy.agent.actor.load_state_dict(torch.load(os.path.join(model_dir, "SAC.pt"), map_location=device))
        return policy
