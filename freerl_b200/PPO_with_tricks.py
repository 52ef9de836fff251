"""PPO with the reference's trick switches and class API (``PPO_file/PPO_with_tricks.py:190-374``) on the fused PPO kernels.

Upstream this file's ``learn()`` cannot run (``np.zeros(self.horizon, dtype=torch.float32)`` raises, ``:302``); what it is meant to
compute — the plain-Adam PPO of ``PPO_advance/PPO.py`` plus switches — is what this class runs, pinned against the reference
executed with that one call patched (``oracle/make_golden_ppo_tricks.py``):

* ``adv_norm``        ``(adv - adv.mean()) / (adv.std() + 1e-8)`` over the horizon after GAE (``:314-315``) -> ``frl_adv_norm``
* ``adam_eps``        both Adams with ``eps = 1e-5`` (``:198-200``)
* ``lr_decay``        ``lr_decay(episode_num, max_episodes)``: linear decay of both learning rates (``:356-362``)
* ``orthogonal_init`` orthogonal weights (gain 1, 0.01 on the action head), zero biases, applied inside each module constructor
                      (``:71-77,90-93,166-169``)
* ``ObsNorm`` / ``reward_norm`` / ``reward_scaling`` live in the train loop, not in the class (``:515-540``) -> ``freerl_b200.vecloop``

* ``tanh``            tanh instead of ReLU hidden activations in actor and critic (``:95,172``; continuous actor and critic — the
                      discrete actor takes no trick argument upstream and keeps ReLU and the default init, ``:106-118``) -> ``frl_ppo_args_t.hidden_tanh`` / ``frl_infer_args_t.hidden_tanh``

* ``Batch_ObsNorm``   running statistics over the MEAN of each rollout (``Normalization_batch_size``, ``:227-228``): ``learn`` folds the rollout's
                      observations in and trains on the normalised ``obs`` / ``next_obs`` (``:296-298``), ``select_action`` applies them with
                      ``update=False`` (``:235-236``); like upstream ``evaluate_action`` does not normalise

* ``beta=True``       the Beta policy head (``Actor_Beta``, ``:120-150``): ``alpha, beta = softplus(.) + 1`` from two heads stored as one device layer,
                      ``Beta.log_prob`` / ``.entropy`` and their gradients (digamma / trigamma) inside the fused update (``frl_ppo_args_t.continuous = 2``);
                      sampling in ``select_action`` is torch's own gamma sampler on the kernel's network output; ``evaluate_action`` = ``2 (mean - 0.5)``
"""
import os

import torch
from torch import nn

from . import _lib
from .PPO_advance import PPO as _PPOAdvance

_TRICKS = ('adv_norm', 'ObsNorm', 'Batch_ObsNorm', 'reward_norm', 'reward_scaling', 'lr_decay', 'orthogonal_init', 'adam_eps', 'tanh')


def orthogonal_init(layer, gain=1.0):
    """``PPO_with_tricks.py:71-77``"""
    nn.init.orthogonal_(layer.weight, gain=gain)
    nn.init.constant_(layer.bias, 0)


class PPO(_PPOAdvance):
    def __init__(self, dim_info, is_continue, actor_lr, critic_lr, horizon, device, trick=None, beta=False, mode=None):
        t = {k: False for k in _TRICKS}
        t.update(trick or {})
        self._beta = bool(beta and is_continue)          # Actor_Beta exists for continuous actions only (:187-192)
        # Actor_discrete takes no trick argument upstream (:106-118, :187): ReLU body and default init whatever the switches say
        self.hidden_tanh = (3 if is_continue else 2) if t['tanh'] else 0
        self.adam_eps = 1e-5 if t['adam_eps'] else 1e-8
        self.actor_dist = {'Beta': self._beta}
        print('actor_dist:Beta' if self._beta else 'actor_dist:Gaussian')
        if t['orthogonal_init']:
            def hook(module, names, which):
                if which == "actor" and not is_continue:
                    return
                for n in names:
                    orthogonal_init(getattr(module, n), gain=0.01 if (which == "actor" and n in ("mean_layer", "alpha_layer", "beta_layer")) else 1.0)
            self._init_hook = hook
        super().__init__(dim_info, is_continue, actor_lr, critic_lr, horizon, device, trick=t, mode=mode)
        self.actor_lr, self.critic_lr = actor_lr, critic_lr
        if t['Batch_ObsNorm']:
            from .normalization import Normalization_batch_size
            self.batch_size_obs_norm = Normalization_batch_size(shape=self.obs_dim, device=self.device)       # read by select_action

    def learn(self, minibatch_size, gamma, lmbda, clip_param, K_epochs, entropy_coefficient, *, permutations=None):
        b = self.buffer
        raw = None
        if self.trick['Batch_ObsNorm']:          # :296-298 — statistics updated with this rollout's mean, next_obs normalised without update
            raw = (b.obs, b.next_obs)
            b.obs = self.batch_size_obs_norm(raw[0]).contiguous()
            b.next_obs = self.batch_size_obs_norm(raw[1], update=False).contiguous()
        try:
            self._learn_normalised(minibatch_size, gamma, lmbda, clip_param, K_epochs, entropy_coefficient, permutations)
        finally:
            if raw is not None:
                b.obs, b.next_obs = raw

    def _learn_normalised(self, minibatch_size, gamma, lmbda, clip_param, K_epochs, entropy_coefficient, permutations):
        adv, v_target = self.compute_gae(gamma, lmbda)
        if self.trick['adv_norm']:
            _lib.check(_lib.lib().frl_adv_norm(_lib.ptr(adv), adv.numel(), 1e-8, _lib.ptr(adv), _lib.stream_ptr(self.device)), "frl_adv_norm")
        self.last_adv, self.last_v_target = adv, v_target
        self._update(adv, v_target, minibatch_size, K_epochs, clip_param, entropy_coefficient, permutations)
        self.buffer.clear()

    def lr_decay(self, episode_num, max_episodes):
        self.agent.lr = self.actor_lr * (1 - episode_num / max_episodes)
        self.agent.lr_critic = self.critic_lr * (1 - episode_num / max_episodes)

    @staticmethod
    def load(dim_info, is_continue, model_dir, trick=None, beta=False, device=None):
        device = device if device is not None else torch.device("cuda")
        policy = PPO(dim_info, is_continue, 0, 0, 0, device=device, trick=trick, beta=beta)
        policy.agent.actor.load_state_dict(torch.load(os.path.join(model_dir, "PPO.pt"), map_location=device))
        return policy
