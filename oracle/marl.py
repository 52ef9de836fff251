"""CPU restatement (PyTorch fp32 + autograd) of the multi-agent learn steps.  TEST INFRASTRUCTURE ONLY (see oracle/algos.py).

MADDPG: ``MADDPG_file/MADDPG.py:186-237`` — per agent i: a FRESH ``sample`` (new indices for every agent inside the
loop, :211), ``next_action_j = actor_target_j(next_obs_j)`` for all j, ``y_i = r_i + gamma Q'_i(all s', all a') (1-d_i)``,
critic MSE (Adam with L2 weight_decay 1e-3, clip 0.5), actor loss ``-Q_i(all s, a with a_i <- pi_i(s_i))`` (clip 0.5);
Polyak of every agent's actor and critic only AFTER the agent loop.
Pinned by ``tests/golden/maddpg.npz`` / ``mappo.npz`` (generated from the reference by ``oracle/make_golden_marl.py``).
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

from .algos import (AdamState, adam_step, clip_grad_norm, clone_net, mlp2, polyak, tanh_actor, _leaf)


class MADDPGOracle:
    def __init__(self, actors, critics, actor_lr, critic_lr, weight_decay=True):
        """actors / critics: OrderedDict agent_id -> net dict"""
        self.ids = list(actors.keys())
        self.actor = OrderedDict((k, _leaf(v)) for k, v in actors.items())
        self.critic = OrderedDict((k, _leaf(v)) for k, v in critics.items())
        self.actor_target = OrderedDict((k, clone_net(v)) for k, v in actors.items())
        self.critic_target = OrderedDict((k, clone_net(v)) for k, v in critics.items())
        self.opt_a = {k: AdamState(list(self.actor[k].values()), actor_lr) for k in self.ids}
        self.opt_c = {k: AdamState(list(self.critic[k].values()), critic_lr, weight_decay=1e-3 if weight_decay else 0.0) for k in self.ids}

    def learn(self, batches, gamma, tau):
        """batches: list (one per agent, in agent order) of dict agent_id -> (obs, act, rew, nobs, done)"""
        out = []
        for aid, batch in zip(self.ids, batches):
            obs = [batch[k][0] for k in self.ids]
            act = [batch[k][1] for k in self.ids]
            nobs = [batch[k][3] for k in self.ids]
            with torch.no_grad():
                nact = [tanh_actor(self.actor_target[k], batch[k][3]) for k in self.ids]
                nq = mlp2(self.critic_target[aid], torch.cat(nobs + nact, dim=1))
                target = batch[aid][2] + gamma * nq * (1 - batch[aid][4])
            q = mlp2(self.critic[aid], torch.cat(obs + act, dim=1))
            critic_loss = F.mse_loss(q, target)
            cp = list(self.critic[aid].values())
            g, _ = clip_grad_norm(torch.autograd.grad(critic_loss, cp), 0.5)
            adam_step(cp, g, self.opt_c[aid])
            new_a = tanh_actor(self.actor[aid], batch[aid][0])
            act2 = [new_a if k == aid else batch[k][1] for k in self.ids]
            actor_loss = -mlp2(self.critic[aid], torch.cat(obs + act2, dim=1)).mean()
            ap = list(self.actor[aid].values())
            ga, _ = clip_grad_norm(torch.autograd.grad(actor_loss, ap), 0.5)
            adam_step(ap, ga, self.opt_a[aid])
            out.append((critic_loss.item(), actor_loss.item()))
        for k in self.ids:
            polyak(self.actor_target[k], self.actor[k], tau)
            polyak(self.critic_target[k], self.critic[k], tau)
        return out
