"""``DDPG_file/DDPG_simple.py:43-222`` — DDPG without the supplements (no critic weight decay, torch-default init, no
Batch_ObsNorm) and without the ``supplement`` constructor argument — on the fused actor-critic kernel."""
import os

import torch

from .DDPG import DDPG as _DDPG


class DDPG(_DDPG):
    def __init__(self, dim_info, is_continue, actor_lr, critic_lr, buffer_size, device, trick=None, mode=None):
        super().__init__(dim_info, is_continue, actor_lr, critic_lr, buffer_size, device, trick=trick, supplement=None, mode=mode)

    @staticmethod
    def load(dim_info, is_continue, model_dir, trick=None, device=None):
        device = device if device is not None else torch.device("cuda")
        policy = DDPG(dim_info, is_continue, 0, 0, 0, device=device, trick=trick)
        policy.agent.actor.load_state_dict(torch.load(os.path.join(model_dir, "DDPG.pt"), map_location=device))
        return policy
