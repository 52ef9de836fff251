"""Plain-Adam PPO with the reference's class API (``PPO_advance/PPO.py:133-271``) on the fused PPO kernels.

Same rollout store, GAE scan and clipped-surrogate minibatch kernel as :mod:`freerl_b200.PPO`; what differs from
``PPO_file/PPO.py`` is the optimiser — two ``torch.optim.Adam`` (eps 1e-8), ``actor_lr`` for the actor and ``critic_lr`` for
the critic, ``clip_grad_norm_(0.5)`` on each, actor step then critic step per minibatch (``PPO_advance/PPO.py:118-133``) —
and the discrete head, which returns softmax probabilities for ``Categorical(probs=...)`` (``:89,155,236``; the same
distribution as the logits head, so the fused log-softmax path serves it).  In the kernel the two Adams are ONE sweep over
the merged parameter block with a per-network learning rate (``frl_ppo_args_t.lr_critic``): Adam is element-wise, the clip
norms are per network already, and the critic's forward does not depend on the actor, so the result is identical to the
reference's two sequential steps.
"""
import os

import torch

from . import _lib
from .PPO import PPO as _PPO


class PPO(_PPO):
    optimizer = _lib.OPT_ADAM
    adam_eps = 1e-8

    def __init__(self, dim_info, is_continue, actor_lr, critic_lr, horizon, device, trick=None, mode=None):
        super().__init__(dim_info, is_continue, actor_lr, critic_lr, horizon, device, trick=trick, mode=mode)
        self.agent.lr_critic = critic_lr

    @staticmethod
    def load(dim_info, is_continue, model_dir, trick=None, device=None):
        device = device if device is not None else torch.device("cuda")
        policy = PPO(dim_info, is_continue, 0, 0, 0, device=device, trick=trick)
        policy.agent.actor.load_state_dict(torch.load(os.path.join(model_dir, "PPO.pt"), map_location=device))
        return policy
