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
    def __init__(self, actors, critics, actor_lr, critic_lr, weight_decay=True, obs_norms=None):
        """actors / critics: OrderedDict agent_id -> net dict; obs_norms: agent_id -> algos.BatchObsNorm (supplement
        Batch_ObsNorm: EVERY agent's statistics are updated by EVERY agent's sample(), MADDPG.py:192-196)"""
        self.ids = list(actors.keys())
        self.obs_norms = obs_norms
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
            if self.obs_norms is not None:
                batch = {k: (self.obs_norms[k](b[0], update=True), b[1], b[2], self.obs_norms[k](b[3], update=False), b[4])
                         for k, b in batch.items()}
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


class MATD3Oracle(MADDPGOracle):
    """``MADDPG_file/MATD3_simple.py:151-262``: MADDPG with twin centralised critics (``Critic_TD3``: l1-l3 / l4-l6),
    target policy smoothing on EVERY agent's next action (``:203-205``, one ``randn_like`` per agent per sample), clipped
    double-Q targets, ``Q1`` only in the actor loss, and actor + Polyak updates only when ``total_it % policy_freq == 0``
    (``total_it`` counts learn() calls, not agent updates).  No weight decay, default torch init."""

    def __init__(self, actors, critics, actor_lr, critic_lr):
        super().__init__(actors, critics, actor_lr, critic_lr, weight_decay=False)
        self.total_it = 0

    @staticmethod
    def _twin(net, x):
        return mlp2(net, x, ("l1", "l2", "l3")), mlp2(net, x, ("l4", "l5", "l6"))

    def learn(self, batches, noises, gamma, tau, policy_noise_scale, policy_noise, noise_clip, max_action, policy_freq):
        """batches[i]: dict agent_id -> (obs, act, rew, nobs, done) of agent i's fresh sample; noises[i][j]: randn [B, act_j]"""
        self.total_it += 1
        out = []
        for (aid, batch), nz in zip(zip(self.ids, batches), noises):
            obs = [batch[k][0] for k in self.ids]
            act = [batch[k][1] for k in self.ids]
            nobs = [batch[k][3] for k in self.ids]
            with torch.no_grad():
                nact = []
                for j, k in enumerate(self.ids):
                    noise = (policy_noise_scale * (nz[j] * policy_noise)).clamp(-noise_clip, noise_clip)
                    nact.append((tanh_actor(self.actor_target[k], batch[k][3]) * max_action + noise).clamp(-max_action, max_action) / max_action)
                q1, q2 = self._twin(self.critic_target[aid], torch.cat(nobs + nact, dim=1))
                target = batch[aid][2] + gamma * torch.min(q1, q2) * (1 - batch[aid][4])
            c1, c2 = self._twin(self.critic[aid], torch.cat(obs + act, dim=1))
            critic_loss = F.mse_loss(c1, target) + F.mse_loss(c2, target)
            cp = list(self.critic[aid].values())
            g, _ = clip_grad_norm(torch.autograd.grad(critic_loss, cp), 0.5)
            adam_step(cp, g, self.opt_c[aid])
            rec = [critic_loss.item(), None]
            if self.total_it % policy_freq == 0:
                new_a = tanh_actor(self.actor[aid], batch[aid][0])
                act2 = [new_a if k == aid else batch[k][1] for k in self.ids]
                actor_loss = -mlp2(self.critic[aid], torch.cat(obs + act2, dim=1), ("l1", "l2", "l3")).mean()
                ap = list(self.actor[aid].values())
                ga, _ = clip_grad_norm(torch.autograd.grad(actor_loss, ap), 0.5)
                adam_step(ap, ga, self.opt_a[aid])
                rec[1] = actor_loss.item()
            out.append(tuple(rec))
        if self.total_it % policy_freq == 0:
            for k in self.ids:
                polyak(self.actor_target[k], self.actor[k], tau)
                polyak(self.critic_target[k], self.critic[k], tau)
        return out


# --------------------------------------------------------------------------------------------------
# MAPPO  (MAPPO_file/MAPPO.py:106-482)
# --------------------------------------------------------------------------------------------------
def _ln(x):
    return F.layer_norm(x, x.size()[1:])


def mappo_body(net, x, names, trick):
    """``Actor.forward`` / ``Critic.forward`` body (MAPPO.py:143-160, 204-218): optional feature_norm on the input and
    LayerNorm (no affine) after each hidden ReLU."""
    a, b, c = names
    if trick['feature_norm']:
        x = _ln(x)
    x = F.relu(F.linear(x, net[a + ".weight"], net[a + ".bias"]))
    if trick['LayerNorm']:
        x = _ln(x)
    x = F.relu(F.linear(x, net[b + ".weight"], net[b + ".bias"]))
    if trick['LayerNorm']:
        x = _ln(x)
    return F.linear(x, net[c + ".weight"], net[c + ".bias"])


def huber_loss(e, d):
    a = (abs(e) <= d).float()
    b = (abs(e) > d).float()
    return a * e ** 2 / 2 + b * d * (abs(e) - d / 2)


class MAPPOOracle:
    """Continuous-action MAPPO with separated nets.  Quirks kept: ``surr = ratio[mb,1] * adv[index][mb,N]`` (every agent's
    ratio multiplies ALL agents' advantages), ``v_s.repeat(1,N)`` against ``v_target[mb,N]``, value clip around ``v_s``
    + huber with ``max(original, clipped)``, merged Adam(eps 1e-5, lr = actor_lr) over actor+critic, NO grad clip,
    joint advantage normalisation with the unbiased std."""

    def __init__(self, actors, critics, lr, trick, is_continue=True):
        self.ids = list(actors.keys())
        self.actor = OrderedDict((k, _leaf(v)) for k, v in actors.items())
        self.critic = OrderedDict((k, _leaf(v)) for k, v in critics.items())
        self.trick = trick
        self.is_continue = is_continue         # False: Actor_discrete -> Categorical(probs=softmax(l3))  (MAPPO.py:163-186, 405-407)
        self.opt = {k: AdamState(list(self.actor[k].values()) + list(self.critic[k].values()), lr, eps=1e-5) for k in self.ids}

    def actor_dist(self, k, obs):
        mean = torch.tanh(mappo_body(self.actor[k], obs, ("l1", "l2", "mean_layer"), self.trick))
        std = torch.exp(torch.clamp(self.actor[k]["log_std"].expand_as(mean), -20, 2))
        return mean, std

    def values(self, k, joint):
        return mappo_body(self.critic[k], joint, ("l1", "l2", "l3"), self.trick)

    def advantages(self, data, gamma, lmbda):
        """data: dict agent -> (obs, action, reward, next_obs, done, logp, adv_done) tensors over the horizon."""
        ids = self.ids
        with torch.no_grad():
            joint = torch.cat([data[k][0] for k in ids], dim=1)
            joint_n = torch.cat([data[k][3] for k in ids], dim=1)
            vs = torch.cat([self.values(k, joint) for k in ids], dim=1)
            vs_ = torch.cat([self.values(k, joint_n) for k in ids], dim=1)
            reward = torch.cat([data[k][2] for k in ids], dim=1)
            done = torch.cat([data[k][4] for k in ids], dim=1)
            adv_dones = torch.cat([data[k][6] for k in ids], dim=1)
            td = reward + gamma * (1.0 - done) * vs_ - vs
            H = td.shape[0]
            adv = torch.zeros(H, len(ids))
            gae = 0
            for i in reversed(range(H)):
                gae = td[i] + gamma * lmbda * gae * (1.0 - adv_dones[i])
                adv[i] = gae
            v_target = adv + vs
            if self.trick['adv_norm']:
                adv = (adv - adv.mean()) / (adv.std() + 1e-8)
        return adv, v_target

    def learn(self, data, permutations, minibatch_size, gamma, lmbda, clip_param, entropy_coefficient, huber_delta):
        """permutations: dict agent -> list (K epochs) of index permutations"""
        ids = self.ids
        adv, v_target = self.advantages(data, gamma, lmbda)
        H = adv.shape[0]
        losses = []
        for k in ids:
            obs, action, logp_old = data[k][0], data[k][1], data[k][5]
            for perm in permutations[k]:
                for s in range(0, H, minibatch_size):
                    index = perm[s:s + minibatch_size]
                    if self.is_continue:
                        mean, std = self.actor_dist(k, obs[index])
                        dist = torch.distributions.Normal(mean, std)
                        ent = dist.entropy().sum(dim=1, keepdim=True)
                        logp = dist.log_prob(action[index])
                    else:
                        probs = torch.softmax(mappo_body(self.actor[k], obs[index], ("l1", "l2", "l3"), self.trick), dim=1)
                        dist = torch.distributions.Categorical(probs=probs)
                        ent = dist.entropy().reshape(-1, 1)
                        logp = dist.log_prob(action[index].reshape(-1)).reshape(-1, 1)
                    ratios = torch.exp(logp.sum(dim=1, keepdim=True) - logp_old[index].sum(dim=1, keepdim=True))
                    surr1 = ratios * adv[index]
                    surr2 = torch.clamp(ratios, 1 - clip_param, 1 + clip_param) * adv[index]
                    actor_loss = -torch.min(surr1, surr2).mean() - entropy_coefficient * ent.mean()
                    joint = torch.cat([data[j][0][index] for j in ids], dim=1)
                    v_s = self.values(k, joint).repeat(1, len(ids))
                    vt = v_target[index]
                    if self.trick['ValueClip']:
                        vt_clip = torch.clamp(vt, v_s - clip_param, v_s + clip_param)
                        if self.trick['huber_loss']:
                            l_clip = huber_loss(vt_clip - v_s, huber_delta).mean()
                            l_orig = huber_loss(vt - v_s, huber_delta).mean()
                        else:
                            l_clip = F.mse_loss(vt_clip, v_s)
                            l_orig = F.mse_loss(vt, v_s)
                        critic_loss = torch.max(l_orig, l_clip)
                    elif self.trick['huber_loss']:
                        critic_loss = huber_loss(vt - v_s, huber_delta).mean()
                    else:
                        critic_loss = F.mse_loss(vt, v_s)
                    ap, cp = list(self.actor[k].values()), list(self.critic[k].values())
                    ga = torch.autograd.grad(actor_loss, ap)
                    gc = torch.autograd.grad(critic_loss, cp)
                    adam_step(ap + cp, list(ga) + list(gc), self.opt[k])
                    losses.append((actor_loss.item(), critic_loss.item()))
        return {"adv": adv, "v_target": v_target, "losses": losses}


# --------------------------------------------------------------------------------------------------
# IPPO  (MAPPO_file/IPPO.py:100-330)
# --------------------------------------------------------------------------------------------------
class IPPOOracle:
    """``IPPO.py:249-316``: every agent is an independent PPO on ITS OWN observation — decentralised critic (``Critic`` takes the
    agent's obs_dim, ``:146-156``), per-agent flat GAE in float64 numpy and per-agent ``adv_norm`` (``:255-270``), the MAPPO
    network body (feature_norm / LayerNorm), separate ``Adam(eps=1e-5)`` for actor (``actor_lr``) and critic (``critic_lr``) with
    ``clip_grad_norm_(0.5)`` each (``:165-184``), ValueClip + huber value loss; the discrete actor returns softmax
    probabilities for ``Categorical(probs=...)`` (``:118-129``)."""

    def __init__(self, actors, critics, actor_lr, critic_lr, trick, is_continue):
        self.ids = list(actors.keys())
        self.actor = OrderedDict((k, _leaf(v)) for k, v in actors.items())
        self.critic = OrderedDict((k, _leaf(v)) for k, v in critics.items())
        self.trick, self.is_continue = trick, is_continue
        eps = 1e-5 if trick['adam_eps'] else 1e-8
        self.opt_a = {k: AdamState(list(self.actor[k].values()), actor_lr, eps=eps) for k in self.ids}
        self.opt_c = {k: AdamState(list(self.critic[k].values()), critic_lr, eps=eps) for k in self.ids}

    def dist(self, k, obs):
        if self.is_continue:
            mean = torch.tanh(mappo_body(self.actor[k], obs, ("l1", "l2", "mean_layer"), self.trick))
            std = torch.exp(torch.clamp(self.actor[k]["log_std"].expand_as(mean), -20, 2))
            return torch.distributions.Normal(mean, std)
        return torch.distributions.Categorical(probs=torch.softmax(mappo_body(self.actor[k], obs, ("l1", "l2", "l3"), self.trick), dim=-1))

    def advantages(self, k, data, gamma, lmbda):
        from .algos import gae_reference
        obs, action, reward, next_obs, done, logp_old, adv_dones = data
        with torch.no_grad():
            vs = mappo_body(self.critic[k], obs, ("l1", "l2", "l3"), self.trick)
            vs_ = mappo_body(self.critic[k], next_obs, ("l1", "l2", "l3"), self.trick)
            td = reward + gamma * (1.0 - done) * vs_ - vs
            adv = gae_reference(td.reshape(-1).numpy(), adv_dones.reshape(-1).numpy(), gamma, lmbda)
            adv = torch.as_tensor(adv, dtype=torch.float32).reshape(-1, 1)
            v_target = adv + vs
            if self.trick['adv_norm']:
                adv = (adv - adv.mean()) / (adv.std() + 1e-8)
        return adv, v_target

    def learn(self, data, permutations, minibatch_size, gamma, lmbda, clip_param, entropy_coefficient, huber_delta):
        """data: dict agent -> 7 tensors; permutations: dict agent -> list (K epochs) of index permutations"""
        out = {"losses": [], "adv": {}, "v_target": {}}
        for k in self.ids:
            obs, action, logp_old = data[k][0], data[k][1], data[k][5]
            adv, v_target = self.advantages(k, data[k], gamma, lmbda)
            out["adv"][k], out["v_target"][k] = adv, v_target
            H = adv.shape[0]
            for perm in permutations[k]:
                for s in range(0, H, minibatch_size):
                    index = perm[s:s + minibatch_size]
                    dist = self.dist(k, obs[index])
                    if self.is_continue:
                        ent = dist.entropy().sum(dim=1, keepdim=True)
                        logp = dist.log_prob(action[index])
                    else:
                        ent = dist.entropy().reshape(-1, 1)
                        logp = dist.log_prob(action[index].reshape(-1)).reshape(-1, 1)
                    ratios = torch.exp(logp.sum(dim=1, keepdim=True) - logp_old[index].sum(dim=1, keepdim=True))
                    surr1 = ratios * adv[index]
                    surr2 = torch.clamp(ratios, 1 - clip_param, 1 + clip_param) * adv[index]
                    actor_loss = -torch.min(surr1, surr2).mean() - entropy_coefficient * ent.mean()
                    ap = list(self.actor[k].values())
                    ga, _ = clip_grad_norm(torch.autograd.grad(actor_loss, ap), 0.5)
                    adam_step(ap, list(ga), self.opt_a[k])
                    v_s = mappo_body(self.critic[k], obs[index], ("l1", "l2", "l3"), self.trick)
                    vt = v_target[index]
                    if self.trick['ValueClip']:
                        vt_clip = torch.clamp(vt, v_s - clip_param, v_s + clip_param)
                        if self.trick['huber_loss']:
                            critic_loss = torch.max(huber_loss(vt - v_s, huber_delta).mean(), huber_loss(vt_clip - v_s, huber_delta).mean())
                        else:
                            critic_loss = torch.max(F.mse_loss(vt, v_s), F.mse_loss(vt_clip, v_s))
                    elif self.trick['huber_loss']:
                        critic_loss = huber_loss(vt - v_s, huber_delta).mean()
                    else:
                        critic_loss = F.mse_loss(vt, v_s)
                    cp = list(self.critic[k].values())
                    gc, _ = clip_grad_norm(torch.autograd.grad(critic_loss, cp), 0.5)
                    adam_step(cp, list(gc), self.opt_c[k])
                    out["losses"].append((actor_loss.item(), critic_loss.item()))
        return out


# --------------------------------------------------------------------------------------------------
# HAPPO  (MAPPO_file/HAPPO.py:263-458)
# --------------------------------------------------------------------------------------------------
class HAPPOOracle(MAPPOOracle):
    """``HAPPO.py:338-447`` (continuous actions): MAPPO's joint-observation critics, joint GAE and joint ``adv_norm``, but the
    agents are visited sequentially in ``torch.randperm`` order and agent m's surrogate is weighted row-wise by
    ``factor = prod_{agents before m} exp(logp_new - logp_old)`` evaluated on the FULL horizon with that agent's actor before /
    after its own epochs (``:367-377, 437-445``; float32 numpy).  Separate ``Adam`` for actor / critic (eps 1e-5 with
    ``adam_eps``), ``clip_grad_norm_(0.5)`` each, actor step then critic step (``:236-253``)."""

    def __init__(self, actors, critics, actor_lr, critic_lr, trick):
        super().__init__(actors, critics, actor_lr, trick, is_continue=True)
        eps = 1e-5 if trick['adam_eps'] else 1e-8
        self.opt_a = {k: AdamState(list(self.actor[k].values()), actor_lr, eps=eps) for k in self.ids}
        self.opt_c = {k: AdamState(list(self.critic[k].values()), critic_lr, eps=eps) for k in self.ids}

    def full_logp(self, k, obs, action):
        with torch.no_grad():
            mean, std = self.actor_dist(k, obs)
            return torch.distributions.Normal(mean, std).log_prob(action).sum(dim=1, keepdim=True)

    def learn(self, data, order, permutations, minibatch_size, gamma, lmbda, clip_param, entropy_coefficient, huber_delta):
        import numpy as np
        ids = self.ids
        adv, v_target = self.advantages(data, gamma, lmbda)
        H = adv.shape[0]
        factor = np.ones((H, 1), dtype=np.float32)
        losses = []
        for pos, ai in enumerate(order):
            k = ids[int(ai)]
            obs, action, logp_old = data[k][0], data[k][1], data[k][5]
            if pos != len(ids) - 1:
                old_full = self.full_logp(k, obs, action)
            factor_t = torch.as_tensor(factor, dtype=torch.float32).reshape(-1, 1)
            for perm in permutations[k]:
                for s in range(0, H, minibatch_size):
                    index = perm[s:s + minibatch_size]
                    mean, std = self.actor_dist(k, obs[index])
                    dist = torch.distributions.Normal(mean, std)
                    ent = dist.entropy().sum(dim=1, keepdim=True)
                    logp = dist.log_prob(action[index])
                    ratios = torch.exp(logp.sum(dim=1, keepdim=True) - logp_old[index].sum(dim=1, keepdim=True))
                    surr1 = ratios * adv[index]
                    surr2 = torch.clamp(ratios, 1 - clip_param, 1 + clip_param) * adv[index]
                    actor_loss = -(factor_t[index] * torch.min(surr1, surr2)).mean() - entropy_coefficient * ent.mean()
                    ap = list(self.actor[k].values())
                    ga, _ = clip_grad_norm(torch.autograd.grad(actor_loss, ap), 0.5)
                    adam_step(ap, list(ga), self.opt_a[k])
                    joint = torch.cat([data[j][0][index] for j in ids], dim=1)
                    v_s = self.values(k, joint).repeat(1, len(ids))
                    vt = v_target[index]
                    if self.trick['huber_loss']:        # ValueClip: max(original, clipped) == original (see freerl_b200/IPPO.py)
                        critic_loss = huber_loss(vt - v_s, huber_delta).mean()
                        if self.trick['ValueClip']:
                            vt_clip = torch.clamp(vt, v_s - clip_param, v_s + clip_param)
                            critic_loss = torch.max(critic_loss, huber_loss(vt_clip - v_s, huber_delta).mean())
                    else:
                        critic_loss = F.mse_loss(vt, v_s)
                        if self.trick['ValueClip']:
                            critic_loss = torch.max(critic_loss, F.mse_loss(torch.clamp(vt, v_s - clip_param, v_s + clip_param), v_s))
                    cp = list(self.critic[k].values())
                    gc, _ = clip_grad_norm(torch.autograd.grad(critic_loss, cp), 0.5)
                    adam_step(cp, list(gc), self.opt_c[k])
                    losses.append((actor_loss.item(), critic_loss.item()))
            if pos != len(ids) - 1:
                new_full = self.full_logp(k, obs, action)
                factor = factor * torch.exp(new_full - old_full).reshape(-1, 1).numpy()
        return {"adv": adv, "v_target": v_target, "losses": losses, "factor": factor}
