"""``MADDPG_file/MADDPG_simple.py:107-200`` — MADDPG without the supplements (no critic weight decay, torch-default init, no
Batch_ObsNorm) and without the ``supplement`` constructor argument — on the fused multi-agent actor-critic kernel."""
import os

import torch

from .MADDPG import MADDPG as _MADDPG

_OFF = {"weight_decay": False, "OUNoise": False, "ObsNorm": False, "net_init": False, "Batch_ObsNorm": False}


class MADDPG(_MADDPG):
    def __init__(self, dim_info, is_continue, actor_lr, critic_lr, buffer_size, device, trick=None, mode=None):
        super().__init__(dim_info, is_continue, actor_lr, critic_lr, buffer_size, device, trick, dict(_OFF), mode=mode)

    @staticmethod
    def load(dim_info, is_continue, model_dir, trick=None, device=None):
        device = device if device is not None else torch.device("cuda")
        policy = MADDPG(dim_info, is_continue, 0, 0, 0, device, trick=trick)
        data = torch.load(os.path.join(model_dir, 'MADDPG.pth'), map_location=device)
        for agent_id, agent in policy.agents.items():
            agent.actor.load_state_dict(data[agent_id])
        return policy
