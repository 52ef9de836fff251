"""HAPPO with the reference's class API (``MAPPO_file/HAPPO.py:226-470``) on the fused PPO kernels.

MAPPO's centralised (joint-observation) critics, joint GAE and joint ``adv_norm``; the agents are updated SEQUENTIALLY in
``torch.randperm`` order and agent m's clipped surrogate is weighted row-wise by ``factor = prod exp(logp_new - logp_old)`` of
the agents visited before it, evaluated on the full horizon (``HAPPO.py:365-377, 437-445``).  Since ``factor > 0``,
``factor * min(r A, clip(r) A) = min(r (factor A), clip(r) (factor A))``: the factor is folded into the advantage columns the
kernel reads, no kernel change.  Separate Adams for actor / critic (eps 1e-5 with ``adam_eps``; one per-network-lr sweep) and
``clip_grad_norm_(0.5)`` each (``:236-253``).  Continuous actions; the reference's discrete factor update indexes the LAST
minibatch instead of the horizon (``:443-444``) and is not reproduced.
"""
import os

import numpy as np
import torch

from . import _common, _lib
from .MAPPO import MAPPO as _MAPPO


class HAPPO(_MAPPO):
    max_norm = 0.5

    def __init__(self, dim_info, is_continue, actor_lr, critic_lr, horizon, device, trick=None, mode=None):
        super().__init__(dim_info, is_continue, actor_lr, critic_lr, horizon, device, trick=trick, mode=mode)
        for ag in self.agents.values():
            ag.lr_critic = critic_lr
        self.adam_eps = 1e-5 if trick['adam_eps'] else 1e-8
        self.agent_ids = list(self.agents.keys())
        self.actor_lr, self.critic_lr = actor_lr, critic_lr

    def lr_decay(self, episode_num, max_episodes):
        self._lr_decay_two_optimisers(episode_num, max_episodes)

    def _full_logp(self, agent_id):
        """log-probability of the STORED actions over the whole horizon with the agent's current actor: sum_j log N(action_j | mean_j, std_j),
        or log p(action) of the categorical head"""
        ag, b = self.agents[agent_id], self.buffers[agent_id]
        ad = self.dim_info[agent_id][1]
        raw = _common.infer(ag._net, b.obs, _lib.INFER_RAW, self.device, ad, l0=0, nl=3, layer_norm=self.layer_norm)
        if not self.is_continue:
            return torch.log_softmax(raw, dim=1).gather(1, b.actions.long())
        mean = torch.tanh(raw)
        std = torch.exp(torch.clamp(ag._net.extra().view(1, -1).expand_as(mean), -20, 2))
        return torch.distributions.Normal(mean, std).log_prob(b.actions).sum(dim=1, keepdim=True)

    def learn(self, minibatch_size, gamma, lmbda, clip_param, K_epochs, entropy_coefficient, huber_delta=None, *, permutations=None,
              order=None):
        adv, v_target, joint = self.compute_advantages(gamma, lmbda)
        self.last_adv, self.last_v_target = adv, v_target
        N = self.num_agents
        if order is None:
            order = torch.randperm(N).numpy()                                  # HAPPO.py:364 (CPU generator)
        if not self.is_continue and minibatch_size != self.horizon:
            # upstream the discrete factor update evaluates the new log-probs on the rows of the LAST minibatch only (HAPPO.py:449-450) and
            # then subtracts the full-horizon old ones: the shapes only broadcast when one minibatch covers the horizon
            raise RuntimeError("HAPPO (discrete): the size of tensor a (%d) must match the size of tensor b (%d) at non-singleton dimension 0 "
                               "— the reference's discrete factor update needs minibatch_size == horizon" % (minibatch_size, self.horizon))
        factor = torch.ones((self.horizon, 1), dtype=torch.float32, device=self.device)
        outs = []
        for pos, ai in enumerate(order):
            agent_id = self.agent_ids[int(ai)]
            ag = self.agents[agent_id]
            last = pos == N - 1
            old = None if last else self._full_logp(agent_id)
            perms = self._permutations(ag, K_epochs, None if permutations is None else permutations[agent_id])
            adv_f = (factor * adv).contiguous()
            outs.append(self._agent_update(agent_id, adv_f, v_target, joint, minibatch_size, K_epochs, clip_param, entropy_coefficient,
                                           huber_delta, perms))
            if not last:
                new = self._full_logp(agent_id)
                if not self.is_continue:
                    # reference quirk kept: the new log-probs are in the order of the last minibatch's permutation, the old ones in time order
                    idx_last = self._keep[0][-1]
                    new = new[idx_last]
                factor = factor * torch.exp(new - old)
        self.last_factor = factor
        self.last_metrics = torch.cat(outs)               # rows in VISITING order
        for buffer in self.buffers.values():
            buffer.clear()

    def save(self, model_dir):
        torch.save({name: {k: v.detach().clone().cpu() for k, v in agent.actor.state_dict().items()} for name, agent in self.agents.items()},
                   os.path.join(model_dir, 'HAPPO.pth'))

    @staticmethod
    def load(dim_info, is_continue, model_dir, trick=None, device=None):
        device = device if device is not None else torch.device("cuda")
        policy = HAPPO(dim_info, is_continue=is_continue, actor_lr=0, critic_lr=0, horizon=0, device=device, trick=trick)
        data = torch.load(os.path.join(model_dir, 'HAPPO.pth'), map_location=device)
        for agent_id, agent in policy.agents.items():
            agent.actor.load_state_dict(data[agent_id])
        return policy
