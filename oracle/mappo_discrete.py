"""CPU restatement of ``MAPPO_file/MAPPO_discrete.py`` (shared actor / critic over N homogeneous agents, episode-batched
``ReplayBuffer``) — TEST INFRASTRUCTURE ONLY: imported by ``tests/`` as the checker, never by ``freerl_b200/``.
Pinned against fixtures generated from the unmodified reference (``oracle/make_golden_mappo_discrete.py`` ->
``tests/golden/mappo_discrete_{simple,clip}.npz``; ``tests/test_parity_mappo_discrete.py::test_oracle_vs_reference_fixture``).

What the reference does (and this file restates, plain torch + autograd):
  * nets ``MAPPO_discrete.py:95-153``: actor obs->128->128->A with softmax head, critic (N*obs)->128->128->1; ReLU bodies.
    ``LayerNorm`` / ``feature_norm`` call ``F.layer_norm(x, x.size()[1:])``: inside ``learn`` the tensors are 4-D, so the statistics run
    jointly over (episode step, agent, feature) of each episode, while acting normalises per row.  The discrete actor computes the
    normalised input and then overwrites it (``:113-115``): ``feature_norm`` only reaches the critic.
  * ONE Adam over actor + critic, lr = actor_lr, eps 1e-5 with ``adam_eps`` (``:160-166``); ``update_ac`` clips the JOINT gradient norm
    to 10 and steps (``:188-192``), and ``learn`` then calls ``ac_optimizer.step()`` AGAIN on the same gradients (``:371``).
  * GAE ``:302-315``: float32 tensors, ``delta = r + gamma v[t+1] (1 - done) - v[t]``, ``gae = delta + gamma lmbda gae`` with NO done mask on
    the recursion, zero tail per episode; ``v_target = adv + v[:-1]``; ``adv_norm`` over the whole [B, T, N] block (unbiased std, + 1e-8).
  * minibatches ``:326``: ``BatchSampler(SequentialSampler(range(batch_size)), minibatch_size, False)`` over EPISODES — no shuffling.
  * losses ``:333-361``: ``-min(ratio adv, clamp(ratio) adv) - c entropy`` and ``(V - v_target)^2`` (or the ``ValueClip`` maximum; with ``huber_loss`` the
    squared maximum of the two batch-mean huber SCALARS), each averaged over (episodes, steps, agents), summed.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _ln(x):
    return F.layer_norm(x, x.size()[1:])          # every axis but the first: (T, N, features) inside learn, (features,) when acting


class _Actor(nn.Module):
    def __init__(self, od, ad, trick):
        super().__init__()
        self.l1, self.l2, self.l3 = nn.Linear(od, 128), nn.Linear(128, 128), nn.Linear(128, ad)
        self.ln = bool(trick["LayerNorm"])

    def forward(self, obs):
        x = F.relu(self.l1(obs))                   # :113-115: the feature_norm result is overwritten — the actor reads the raw observation
        if self.ln:
            x = _ln(x)
        x = F.relu(self.l2(x))
        if self.ln:
            x = _ln(x)
        return torch.softmax(self.l3(x), dim=-1)


class _Critic(nn.Module):
    def __init__(self, sd, trick):
        super().__init__()
        self.l1, self.l2, self.l3 = nn.Linear(sd, 128), nn.Linear(128, 128), nn.Linear(128, 1)
        self.ln, self.fn = bool(trick["LayerNorm"]), bool(trick["feature_norm"])

    def forward(self, s):
        if self.fn:
            s = _ln(s)
        q = F.relu(self.l1(s))
        if self.ln:
            q = _ln(q)
        q = F.relu(self.l2(q))
        if self.ln:
            q = _ln(q)
        return self.l3(q)


def huber_loss(e, d):                              # :197-200
    a = (abs(e) <= d).float()
    b = (abs(e) > d).float()
    return a * e ** 2 / 2 + b * d * (abs(e) - d / 2)


class MAPPODiscreteOracle:
    def __init__(self, actor_sd, critic_sd, n_agents, obs_dim, act_dim, actor_lr, trick):
        self.N, self.trick = n_agents, trick
        self.actor, self.critic = _Actor(obs_dim, act_dim, trick), _Critic(n_agents * obs_dim, trick)
        self.actor.load_state_dict(actor_sd)
        self.critic.load_state_dict(critic_sd)
        self.params = list(self.actor.parameters()) + list(self.critic.parameters())
        self.opt = torch.optim.Adam(self.params, lr=actor_lr, eps=1e-5 if trick["adam_eps"] else 1e-8)          # :160-166

    def advantages(self, batch, gamma, lmbda):
        T = batch["r_n"].shape[1]
        v = batch["v_n"]
        deltas = batch["r_n"] + gamma * v[:, 1:] * (1 - batch["done_n"]) - v[:, :-1]                            # :309
        adv, gae = [], 0
        for t in reversed(range(T)):
            gae = deltas[:, t] + gamma * lmbda * gae                                                            # :311 (no done mask)
            adv.insert(0, gae)
        adv = torch.stack(adv, dim=1)
        v_target = adv + v[:, :-1]
        if self.trick["adv_norm"]:
            adv = (adv - adv.mean()) / (adv.std() + 1e-8)                                                       # :316
        return adv, v_target

    def learn(self, batch, minibatch_size, gamma, lmbda, clip_param, K_epochs, entropy_coefficient, huber_delta=None):
        """batch: dict of tensors in the reference's layout ([B, T, N, ...], ``a_n`` long).  Returns adv, v_target and, per
        minibatch update, (actor loss incl. entropy term, critic loss)."""
        with torch.no_grad():
            adv, v_target = self.advantages(batch, gamma, lmbda)
        Bn = batch["r_n"].shape[0]
        actor_in = batch["obs_n"]
        critic_in = batch["s"].unsqueeze(2).repeat(1, 1, self.N, 1)                                            # :282
        losses = []
        for _ in range(K_epochs):
            for lo in range(0, Bn, minibatch_size):                                                             # sequential sampler
                idx = slice(lo, min(lo + minibatch_size, Bn))
                probs = self.actor(actor_in[idx])
                values = self.critic(critic_in[idx]).squeeze(-1)
                dist = torch.distributions.Categorical(probs)
                logp = dist.log_prob(batch["a_n"][idx])
                ratios = torch.exp(logp - batch["a_logprob_n"][idx])
                surr1 = ratios * adv[idx]
                surr2 = torch.clamp(ratios, 1 - clip_param, 1 + clip_param) * adv[idx]
                actor_loss = -torch.min(surr1, surr2) - entropy_coefficient * dist.entropy()
                if self.trick["ValueClip"]:                                                                     # :350-357
                    v_old = batch["v_n"][idx, :-1]
                    e_clip = torch.clamp(values - v_old, -clip_param, clip_param) + v_old - v_target[idx]
                    if self.trick["huber_loss"]:                                                                # :353-355: two batch-mean SCALARS
                        e_clip = huber_loss(e_clip, huber_delta).mean()
                        e_orig = huber_loss(values - v_target[idx], huber_delta).mean()
                    else:
                        e_orig = values - v_target[idx]
                    critic_loss = torch.max(e_clip ** 2, e_orig ** 2)
                else:
                    critic_loss = (values - v_target[idx]) ** 2
                la, lc = actor_loss.mean(), critic_loss.mean()
                self.opt.zero_grad()
                (la + lc).backward()
                torch.nn.utils.clip_grad_norm_(self.params, 10)                                                 # :191
                self.opt.step()
                self.opt.step()                                                                                 # :371, same gradients
                losses.append((float(la.detach()), float(lc.detach())))
        return {"adv": adv, "v_target": v_target, "losses": losses}
