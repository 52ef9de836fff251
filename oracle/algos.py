"""CPU restatement (PyTorch fp32 + autograd) of the reference ``learn()`` steps.

TEST INFRASTRUCTURE ONLY — this module is the *checker*: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu-baseline / ``--impl reference`` legs may import it.  Nothing under
``freerl_b200/`` does; the product path has no CPU fallback.

Pinned against the reference itself: ``oracle/make_golden.py`` runs the unmodified reference classes
from ``/root/reference`` on seeded synthetic batches and stores (state before, batch, noise, losses,
state after); ``tests/test_oracle_golden.py`` replays those fixtures through this restatement.  The
reference holds no tests / golden vectors of its own.

Conventions: a *net* is an ordered dict ``name -> tensor`` in torch ``state_dict`` layout
(``Linear.weight`` is ``[out, in]``); an *optimizer state* is :class:`AdamState`.  All functions take the
sampled batch and every random tensor as explicit arguments so both sides of a parity test consume
identical numbers (RNG call order is listed in SURVEY.md §8c, last row).
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

LOG2 = float(np.log(2))


# --------------------------------------------------------------------------------------------------
# optimisers
# --------------------------------------------------------------------------------------------------
class AdamState:
    """State of one ``torch.optim.Adam`` instance (single param group)."""

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step = 0
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]


@torch.no_grad()
def adam_step(params, grads, st: AdamState):
    """torch.optim.Adam, single-tensor path (what runs on CPU): L2 ``weight_decay`` folded into the
    gradient, ``m.lerp_(g, 1-b1)``, ``v = b2 v + (1-b2) g*g``, ``denom = sqrt(v)/sqrt(bc2) + eps``,
    ``p += (-lr/bc1) * m / denom``.  Call sites: ``DQN_file/DQN.py:54``, ``SAC_file/SAC.py:135-136,158``,
    ``TD3_file/TD3.py:131-132``, ``DDPG_file/DDPG.py:124-134``, ``MADDPG_file/MADDPG.py:117-121``,
    ``MAPPO_file/MAPPO.py:230`` (eps 1e-5)."""
    st.step += 1
    b1, b2 = st.betas
    bc1 = 1 - b1 ** st.step
    bc2_sqrt = (1 - b2 ** st.step) ** 0.5
    step_size = st.lr / bc1
    for p, g, m, v in zip(params, grads, st.m, st.v):
        if g is None:
            continue
        if st.weight_decay != 0:
            g = g.add(p, alpha=st.weight_decay)
        m.lerp_(g, 1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / bc2_sqrt).add_(st.eps)
        p.addcdiv_(m, denom, value=-step_size)


class CautiousAdamWState(AdamState):
    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-6):
        super().__init__(params, lr, betas, eps, 0.0)


@torch.no_grad()
def cautious_adamw_step(params, grads, st: CautiousAdamWState):
    """``PPO_file/c_adamw.py:65-122``: eps (1e-6) added after ``sqrt(v)`` with no bias-correction inside
    the denominator, ``step = lr*sqrt(bc2)/bc1``, cautious mask ``(m*g>0)`` renormalised by its
    PER-TENSOR mean clamped at 1e-3.  (weight_decay is 0 in PPO.py.)"""
    st.step += 1
    b1, b2 = st.betas
    step_size = st.lr * math.sqrt(1.0 - b2 ** st.step) / (1.0 - b1 ** st.step)
    for p, g, m, v in zip(params, grads, st.m, st.v):
        if g is None:
            continue
        m.mul_(b1).add_(g, alpha=1.0 - b1)
        v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
        denom = v.sqrt().add_(st.eps)
        mask = (m * g > 0).to(g.dtype)
        mask.div_(mask.mean().clamp_(min=1e-3))
        p.add_((m * mask) / denom, alpha=-step_size)


def clip_grad_norm(grads, max_norm):
    """``torch.nn.utils.clip_grad_norm_`` (L2): total = ||(||g_i||)_i||, coef = clamp(max/(total+1e-6), max=1).
    Call sites ``SAC_file/SAC.py:144,150`` etc.  Returns (clipped grads, total_norm)."""
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in grads]))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return [g * coef for g in grads], total


@torch.no_grad()
def polyak(target, source, tau):
    """``theta' <- theta'(1-tau) + theta*tau`` per tensor (``DQN_file/DQN.py:120-128`` and siblings)."""
    for k in target:
        target[k].copy_(target[k] * (1.0 - tau) + source[k] * tau)


def clone_net(net):
    return OrderedDict((k, v.detach().clone()) for k, v in net.items())


def _leaf(net):
    return OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in net.items())


# --------------------------------------------------------------------------------------------------
# network forward passes (functional)
# --------------------------------------------------------------------------------------------------
def mlp2(net, x, names=("l1", "l2", "l3"), act=F.relu):
    """relu(l1) -> relu(l2) -> l3 (linear), the 128-128 body every actor / critic uses (``act=torch.tanh``: the ``tanh`` switch of
    ``PPO_file/PPO_with_tricks.py:95,172``)."""
    a, b, c = names
    h = act(F.linear(x, net[a + ".weight"], net[a + ".bias"]))
    h = act(F.linear(h, net[b + ".weight"], net[b + ".bias"]))
    return F.linear(h, net[c + ".weight"], net[c + ".bias"])


def qnet_dqn(net, obs):
    """``DQN_file/DQN.py:32-45``: obs -> relu(l1) -> l2."""
    return F.linear(F.relu(F.linear(obs, net["l1.weight"], net["l1.bias"])), net["l2.weight"], net["l2.bias"])


def sac_actor(net, obs, eps=None, deterministic=False, with_logprob=True):
    """``SAC_file/SAC.py:60-97``: state-independent ``log_std`` clamped to [-20, 2]; ``u = mean + std*eps``;
    ``log_pi = sum logN(u) - sum 2(log2 - u - softplus(-2u))``; action = tanh(u)."""
    mean = mlp2(net, obs, ("l1", "l2", "mean_layer"))
    log_std = torch.clamp(net["log_std"].expand_as(mean), -20, 2)
    std = torch.exp(log_std)
    u = mean if deterministic else mean + std * eps
    log_pi = None
    if with_logprob:
        var = std ** 2
        logn = -((u - mean) ** 2) / (2 * var) - log_std - math.log(math.sqrt(2 * math.pi))
        log_pi = logn.sum(dim=1, keepdim=True)
        log_pi = log_pi - (2 * (LOG2 - u - F.softplus(-2 * u))).sum(dim=1, keepdim=True)
    return torch.tanh(u), log_pi


def twin_q(net, obs, act):
    """``SAC_file/SAC.py:103-127`` / ``TD3_file/TD3.py:87-121``: Q1 = l1-l3, Q2 = l4-l6 on cat(obs, act)."""
    oa = torch.cat([obs, act], dim=1)
    return mlp2(net, oa, ("l1", "l2", "l3")), mlp2(net, oa, ("l4", "l5", "l6"))


def single_q(net, obs, act):
    """``DDPG_file/DDPG.py:89-116`` / ``TD3_file/TD3.py:67-85``."""
    return mlp2(net, torch.cat([obs, act], dim=1))


def tanh_actor(net, obs):
    """``TD3_file/TD3.py:52-65`` / ``DDPG_file/DDPG.py:70-93`` / ``MADDPG_file/MADDPG.py:63-83``."""
    return torch.tanh(mlp2(net, obs))


# --------------------------------------------------------------------------------------------------
# Batch_ObsNorm  (SAC.py:390-421, DDPG.py:372-403, MADDPG.py:366-397)
# --------------------------------------------------------------------------------------------------
class BatchObsNorm:
    """Welford statistics over *batch means*: first call sets mean = std = x_bar."""

    def __init__(self, obs_dim):
        self.n = 0
        self.mean = torch.zeros(obs_dim)
        self.S = torch.zeros(obs_dim)
        self.std = torch.sqrt(self.S)

    def __call__(self, x, update=True):
        if update:
            xb = x.mean(dim=0, keepdim=True)
            self.n += 1
            if self.n == 1:
                self.mean = xb
                self.std = xb
            else:
                old = self.mean
                self.mean = old + (xb - old) / self.n
                self.S = self.S + (xb - old) * (xb - self.mean)
                self.std = torch.sqrt(self.S / self.n)
        return (x - self.mean) / (self.std + 1e-8)


# --------------------------------------------------------------------------------------------------
# DQN  (DQN_file/DQN.py:104-128)
# --------------------------------------------------------------------------------------------------
class DQNOracle:
    def __init__(self, qnet, lr):
        self.q = _leaf(qnet)
        self.q_target = clone_net(qnet)
        self.opt = AdamState(list(self.q.values()), lr)

    def learn(self, batch, gamma, tau):
        obs, act, rew, nobs, done = batch
        with torch.no_grad():
            next_q = qnet_dqn(self.q_target, nobs).max(dim=1)[0].reshape(-1, 1)
            target = rew + gamma * next_q * (1 - done)
        cur = qnet_dqn(self.q, obs).gather(1, act.long())
        loss = F.mse_loss(cur, target)
        grads = torch.autograd.grad(loss, list(self.q.values()))
        adam_step(list(self.q.values()), grads, self.opt)       # no grad clip in DQN (DQN.py:56-59)
        polyak(self.q_target, self.q, tau)
        return {"loss": loss.item(), "grads": [g.clone() for g in grads]}


# --------------------------------------------------------------------------------------------------
# SAC  (SAC_file/SAC.py:222-271)
# --------------------------------------------------------------------------------------------------
class SACOracle:
    def __init__(self, actor, critic, actor_lr, critic_lr, act_dim, alpha0=0.01, alpha_lr=1e-4,
                 adaptive_alpha=True, obs_norm=None):
        self.obs_norm = obs_norm          # BatchObsNorm or None (trick Batch_ObsNorm, applied in sample(): SAC.py:215-217)
        self.actor, self.critic = _leaf(actor), _leaf(critic)
        self.actor_target, self.critic_target = clone_net(actor), clone_net(critic)
        self.opt_a = AdamState(list(self.actor.values()), actor_lr)
        self.opt_c = AdamState(list(self.critic.values()), critic_lr)
        # Alpha: SAC.py:154-169 — log_alpha is a 0-dim fp32 tensor with its own Adam(lr 1e-4).
        self.log_alpha = torch.tensor(np.log(alpha0), dtype=torch.float32, requires_grad=adaptive_alpha)
        self.opt_alpha = AdamState([self.log_alpha], alpha_lr)
        self.alpha = self.log_alpha.exp()
        self.target_entropy = -act_dim
        self.adaptive_alpha = adaptive_alpha

    def learn(self, batch, eps_next, eps_new, gamma, tau):
        obs, act, rew, nobs, done = batch
        if self.obs_norm is not None:
            obs = self.obs_norm(obs)
            nobs = self.obs_norm(nobs, update=False)
        alpha = self.alpha.detach()
        with torch.no_grad():
            na, nlogp = sac_actor(self.actor_target, nobs, eps_next)         # uses the ACTOR TARGET (SAC.py:227)
            q1t, q2t = twin_q(self.critic_target, nobs, na)
            target = rew + gamma * (1 - done) * (torch.min(q1t, q2t) + alpha * (-nlogp))
        q1, q2 = twin_q(self.critic, obs, act)
        critic_loss = F.mse_loss(q1, target) + F.mse_loss(q2, target)
        cp = list(self.critic.values())
        g = torch.autograd.grad(critic_loss, cp)
        g, cnorm = clip_grad_norm(g, 0.5)
        adam_step(cp, g, self.opt_c)

        new_a, logp = sac_actor(self.actor, obs, eps_new)
        entropy = -logp
        q1p, q2p = twin_q(self.critic, obs, new_a)                           # UPDATED critic (SAC.py:246)
        q_pi = torch.mean(torch.stack((q1p, q2p)), dim=0)
        actor_loss = (-q_pi - alpha * entropy).mean()
        ap = list(self.actor.values())
        ga = torch.autograd.grad(actor_loss, ap)
        ga, anorm = clip_grad_norm(ga, 0.5)
        adam_step(ap, ga, self.opt_a)

        polyak(self.critic_target, self.critic, tau)
        polyak(self.actor_target, self.actor, tau)

        out = {"critic_loss": critic_loss.item(), "actor_loss": actor_loss.item(),
               "critic_gnorm": cnorm.item(), "actor_gnorm": anorm.item()}
        if self.adaptive_alpha:
            alpha_loss = (self.log_alpha.exp() * (entropy - self.target_entropy).detach()).mean()
            (ga_,) = torch.autograd.grad(alpha_loss, [self.log_alpha])
            adam_step([self.log_alpha], [ga_], self.opt_alpha)
            self.alpha = self.log_alpha.exp()
            out["alpha_loss"] = alpha_loss.item()
        out["alpha"] = self.alpha.item()
        return out


class SACDiscreteOracle:
    """The ``hands_on`` discrete branch of ``SAC_file/SAC_add_discrete.py`` (:137-177 nets, :299-348 learn, :350-360 targets,
    :206-223 Alpha with target entropy ``0.6 * -log(1 / n_actions)``): softmax actor, twin critic heads ``obs -> Q(s, .)``,
    next-state probabilities from the ONLINE actor, ``log(p + 1e-8)``."""

    def __init__(self, actor, critic, actor_lr, critic_lr, n_actions, alpha0=0.01, alpha_lr=1e-4):
        self.actor, self.critic = _leaf(actor), _leaf(critic)
        self.actor_target, self.critic_target = clone_net(actor), clone_net(critic)
        self.opt_a = AdamState(list(self.actor.values()), actor_lr)
        self.opt_c = AdamState(list(self.critic.values()), critic_lr)
        self.log_alpha = torch.tensor(np.log(alpha0), dtype=torch.float32, requires_grad=True)
        self.opt_alpha = AdamState([self.log_alpha], alpha_lr)
        self.alpha = self.log_alpha.exp()
        self.target_entropy = 0.6 * (-torch.log(torch.tensor(1.0 / n_actions)))

    @staticmethod
    def probs(net, obs):
        return torch.softmax(mlp2(net, obs), dim=1)

    @staticmethod
    def heads(net, obs):
        return mlp2(net, obs, ("l1", "l2", "l3")), mlp2(net, obs, ("l4", "l5", "l6"))

    def learn(self, batch, gamma, tau):
        obs, act, rew, nobs, done = batch
        alpha = self.alpha.detach()
        with torch.no_grad():
            npb = self.probs(self.actor, nobs)
            nlog = torch.log(npb + 1e-8)
            v1t, v2t = self.heads(self.critic_target, nobs)
            next_q = torch.sum(npb * torch.min(v1t, v2t), dim=1, keepdim=True)
            ent_next = -torch.sum(npb * nlog, dim=1, keepdim=True)
            target = rew + gamma * (1 - done) * (next_q + alpha * ent_next)
        v1, v2 = self.heads(self.critic, obs)
        q1, q2 = v1.gather(1, act.long()), v2.gather(1, act.long())
        critic_loss = F.mse_loss(q1, target) + F.mse_loss(q2, target)
        cp = list(self.critic.values())
        g = torch.autograd.grad(critic_loss, cp)
        g, cnorm = clip_grad_norm(g, 0.5)
        adam_step(cp, g, self.opt_c)

        pb = self.probs(self.actor, obs)
        logp = torch.log(pb + 1e-8)
        entropy = -torch.sum(pb * logp, dim=1, keepdim=True)
        with torch.no_grad():
            v1p, v2p = self.heads(self.critic, obs)                          # UPDATED critic
        q_pi = torch.sum(pb * torch.min(v1p, v2p), dim=1, keepdim=True)
        actor_loss = (-q_pi - alpha * entropy).mean()
        ap = list(self.actor.values())
        ga = torch.autograd.grad(actor_loss, ap)
        ga, anorm = clip_grad_norm(ga, 0.5)
        adam_step(ap, ga, self.opt_a)

        polyak(self.critic_target, self.critic, tau)
        polyak(self.actor_target, self.actor, tau)
        alpha_loss = (self.log_alpha.exp() * (entropy - self.target_entropy).detach()).mean()
        (ga_,) = torch.autograd.grad(alpha_loss, [self.log_alpha])
        adam_step([self.log_alpha], [ga_], self.opt_alpha)
        self.alpha = self.log_alpha.exp()
        return {"critic_loss": critic_loss.item(), "actor_loss": actor_loss.item(), "critic_gnorm": cnorm.item(), "actor_gnorm": anorm.item(),
                "alpha_loss": alpha_loss.item(), "alpha": self.alpha.item()}


# --------------------------------------------------------------------------------------------------
# TD3  (TD3_file/TD3.py:189-244)  and DDPG  (DDPG_file/DDPG.py:203-233)
# --------------------------------------------------------------------------------------------------
class TD3Oracle:
    def __init__(self, actor, critic, actor_lr, critic_lr, clip_double=True, policy_noise=True, twin_delay=True,
                 critic_weight_decay=0.0, obs_norm=None):
        self.obs_norm = obs_norm          # BatchObsNorm or None (DDPG.py:196-198)
        self.actor, self.critic = _leaf(actor), _leaf(critic)
        self.actor_target, self.critic_target = clone_net(actor), clone_net(critic)
        self.opt_a = AdamState(list(self.actor.values()), actor_lr)
        self.opt_c = AdamState(list(self.critic.values()), critic_lr, weight_decay=critic_weight_decay)
        self.clip_double, self.policy_noise, self.twin_delay = clip_double, policy_noise, twin_delay
        self.total_it = 0

    def learn(self, batch, randn, gamma, tau, policy_noise=0.1, noise_clip=0.5, max_action=1.0, policy_freq=2,
              policy_noise_scale=1.0):
        self.total_it += 1
        obs, act, rew, nobs, done = batch
        if self.obs_norm is not None:
            obs = self.obs_norm(obs)
            nobs = self.obs_norm(nobs, update=False)
        with torch.no_grad():
            if self.policy_noise:
                noise = (policy_noise_scale * (randn * policy_noise)).clamp(-noise_clip, noise_clip)
                na = (tanh_actor(self.actor_target, nobs) * max_action + noise).clamp(-max_action, max_action) / max_action
            else:
                na = tanh_actor(self.actor_target, nobs)
            if self.clip_double:
                q1t, q2t = twin_q(self.critic_target, nobs, na)
                nq = torch.min(q1t, q2t)
            else:
                nq = single_q(self.critic_target, nobs, na)
            target = rew + gamma * nq * (1 - done)
        if self.clip_double:
            q1, q2 = twin_q(self.critic, obs, act)
            critic_loss = F.mse_loss(q1, target) + F.mse_loss(q2, target)
        else:
            critic_loss = F.mse_loss(single_q(self.critic, obs, act), target)
        cp = list(self.critic.values())
        g = torch.autograd.grad(critic_loss, cp)
        g, _ = clip_grad_norm(g, 0.5)
        adam_step(cp, g, self.opt_c)
        out = {"critic_loss": critic_loss.item()}
        if not self.twin_delay:
            policy_freq = 1
        if self.total_it % policy_freq == 0:
            new_a = tanh_actor(self.actor, obs)
            if self.clip_double:
                oa = torch.cat([obs, new_a], dim=1)
                actor_loss = -mlp2(self.critic, oa, ("l1", "l2", "l3")).mean()          # Q1 only (TD3.py:221)
            else:
                actor_loss = -single_q(self.critic, obs, new_a).mean()
            ap = list(self.actor.values())
            ga = torch.autograd.grad(actor_loss, ap)
            ga, _ = clip_grad_norm(ga, 0.5)
            adam_step(ap, ga, self.opt_a)
            polyak(self.critic_target, self.critic, tau)
            polyak(self.actor_target, self.actor, tau)
            out["actor_loss"] = actor_loss.item()
        return out


class DDPGOracle(TD3Oracle):
    """DDPG = single critic, no target smoothing, actor every step; critic Adam has L2 weight_decay 1e-3
    when ``supplement['weight_decay']`` (``DDPG_file/DDPG.py:131-134``)."""

    def __init__(self, actor, critic, actor_lr, critic_lr, weight_decay=True, obs_norm=None):
        super().__init__(actor, critic, actor_lr, critic_lr, clip_double=False, policy_noise=False,
                         twin_delay=False, critic_weight_decay=1e-3 if weight_decay else 0.0, obs_norm=obs_norm)

    def learn(self, batch, gamma, tau):
        return super().learn(batch, None, gamma, tau)


# --------------------------------------------------------------------------------------------------
# PPO  (PPO_file/PPO.py:213-286)
# --------------------------------------------------------------------------------------------------
def gae_reference(td_delta, adv_dones, gamma, lmbda):
    """``PPO_file/PPO.py:222-231``: float64 numpy reverse scan over the flat horizon, zero-initialised."""
    td = np.asarray(td_delta).reshape(-1)
    ad = np.asarray(adv_dones).reshape(-1)
    adv = np.zeros(td.shape[0])
    gae = 0
    for i in reversed(range(td.shape[0])):
        gae = td[i] + gamma * lmbda * gae * (1.0 - ad[i])
        adv[i] = gae
    return adv


def ppo_actor_cont(net, obs, act=F.relu):
    """``PPO_file/PPO.py:58-76``: mean = tanh(mean_layer(.)), std = exp(clamp(log_std))."""
    mean = torch.tanh(mlp2(net, obs, ("l1", "l2", "mean_layer"), act=act))
    std = torch.exp(torch.clamp(net["log_std"].expand_as(mean), -20, 2))
    return mean, std


class PPOOracle:
    act = staticmethod(F.relu)          # critic hidden activation (PPOTricksOracle(tanh=True) switches to torch.tanh)
    act_actor = staticmethod(F.relu)    # actor hidden activation (tanh only for the continuous actor: Actor_discrete has no switch)

    def __init__(self, actor, critic, lr, is_continue):
        self.actor, self.critic = _leaf(actor), _leaf(critic)
        self.is_continue = is_continue
        self.params = list(self.actor.values()) + list(self.critic.values())   # merged list, lr = actor_lr (PPO.py:121)
        self.opt = CautiousAdamWState(self.params, lr)

    def advantages(self, data, gamma, lmbda):
        obs, action, reward, next_obs, done, logp_old, adv_dones = data
        with torch.no_grad():
            vs = mlp2(self.critic, obs, act=self.act)
            vs_ = mlp2(self.critic, next_obs, act=self.act)
            td = reward + gamma * (1.0 - done) * vs_ - vs
            adv = gae_reference(td.reshape(-1).numpy(), adv_dones.reshape(-1).numpy(), gamma, lmbda)
            adv = torch.as_tensor(adv, dtype=torch.float32).reshape(-1, 1)
            v_target = adv + vs
        return adv, v_target

    def minibatch(self, data, adv, v_target, index, clip_param, entropy_coefficient):
        obs, action, reward, next_obs, done, logp_old, adv_dones = data
        if self.is_continue:
            mean, std = ppo_actor_cont(self.actor, obs[index])
            dist = torch.distributions.Normal(mean, std)
            ent = dist.entropy().sum(dim=1, keepdim=True)
            logp = dist.log_prob(action[index])
        else:
            dist = torch.distributions.Categorical(logits=mlp2(self.actor, obs[index]))
            ent = dist.entropy().reshape(-1, 1)
            logp = dist.log_prob(action[index].reshape(-1)).reshape(-1, 1)
        ratios = torch.exp(logp.sum(dim=1, keepdim=True) - logp_old[index].sum(dim=1, keepdim=True))
        surr1 = ratios * adv[index]
        surr2 = torch.clamp(ratios, 1 - clip_param, 1 + clip_param) * adv[index]
        actor_loss = -torch.min(surr1, surr2).mean() - entropy_coefficient * ent.mean()
        v_s = mlp2(self.critic, obs[index])
        critic_loss = F.mse_loss(v_target[index], v_s)
        ap, cp = list(self.actor.values()), list(self.critic.values())
        ga = torch.autograd.grad(actor_loss, ap)
        gc = torch.autograd.grad(critic_loss, cp)
        ga, _ = clip_grad_norm(ga, 0.5)
        gc, _ = clip_grad_norm(gc, 0.5)
        cautious_adamw_step(self.params, list(ga) + list(gc), self.opt)
        return actor_loss.item(), critic_loss.item()

    def learn(self, data, permutations, minibatch_size, gamma, lmbda, clip_param, entropy_coefficient):
        """``permutations``: list (len K_epochs) of index permutations of the horizon (``np.random.permutation``)."""
        adv, v_target = self.advantages(data, gamma, lmbda)
        horizon = data[0].shape[0]
        losses = []
        for perm in permutations:
            for s in range(0, horizon, minibatch_size):
                losses.append(self.minibatch(data, adv, v_target, perm[s:s + minibatch_size], clip_param,
                                             entropy_coefficient))
        return {"adv": adv, "v_target": v_target, "losses": losses}


class PPOAdvanceOracle(PPOOracle):
    """``PPO_advance/PPO.py:118-119,122-133,198-260``: the standard PPO — separate ``torch.optim.Adam`` (eps 1e-8) for actor
    (``actor_lr``) and critic (``critic_lr``), ``clip_grad_norm_(0.5)`` each, actor step then critic step per minibatch;
    the discrete actor returns ``softmax`` probabilities fed to ``Categorical(probs=...)`` (``:89,236``)."""

    def __init__(self, actor, critic, actor_lr, critic_lr, is_continue):
        self.actor, self.critic = _leaf(actor), _leaf(critic)
        self.is_continue = is_continue
        self.opt_a = AdamState(list(self.actor.values()), actor_lr)
        self.opt_c = AdamState(list(self.critic.values()), critic_lr)

    def minibatch(self, data, adv, v_target, index, clip_param, entropy_coefficient):
        obs, action, reward, next_obs, done, logp_old, adv_dones = data
        if self.is_continue:
            mean, std = ppo_actor_cont(self.actor, obs[index], act=self.act_actor)
            dist = torch.distributions.Normal(mean, std)
            ent = dist.entropy().sum(dim=1, keepdim=True)
            logp = dist.log_prob(action[index])
        else:
            dist = torch.distributions.Categorical(probs=torch.softmax(mlp2(self.actor, obs[index], act=self.act_actor), dim=1))
            ent = dist.entropy().reshape(-1, 1)
            logp = dist.log_prob(action[index].reshape(-1)).reshape(-1, 1)
        ratios = torch.exp(logp.sum(dim=1, keepdim=True) - logp_old[index].sum(dim=1, keepdim=True))
        surr1 = ratios * adv[index]
        surr2 = torch.clamp(ratios, 1 - clip_param, 1 + clip_param) * adv[index]
        actor_loss = -torch.min(surr1, surr2).mean() - entropy_coefficient * ent.mean()
        ap = list(self.actor.values())
        ga, _ = clip_grad_norm(torch.autograd.grad(actor_loss, ap), 0.5)
        adam_step(ap, list(ga), self.opt_a)
        v_s = mlp2(self.critic, obs[index], act=self.act)
        critic_loss = F.mse_loss(v_target[index], v_s)
        cp = list(self.critic.values())
        gc, _ = clip_grad_norm(torch.autograd.grad(critic_loss, cp), 0.5)
        adam_step(cp, list(gc), self.opt_c)
        return actor_loss.item(), critic_loss.item()


class PPOTricksOracle(PPOAdvanceOracle):
    """``PPO_file/PPO_with_tricks.py:190-362`` with the ``np.zeros(dtype=torch.float32)`` call (``:302``) read as float32 zeros:
    ``PPOAdvanceOracle`` plus ``adam_eps`` (both Adams eps 1e-5, ``:198-200``), ``adv_norm`` (``:314-315``) and ``lr_decay``
    (``:356-362``)."""

    def __init__(self, actor, critic, actor_lr, critic_lr, is_continue, adam_eps=False, adv_norm=False, tanh=False, beta=False):
        super().__init__(actor, critic, actor_lr, critic_lr, is_continue)
        self.beta = bool(beta and is_continue)          # Actor_Beta (``:120-150``): alpha, beta = softplus(head) + 1
        if tanh:
            self.act = torch.tanh
            if is_continue:
                self.act_actor = torch.tanh
        self.actor_lr, self.critic_lr, self.adv_norm = actor_lr, critic_lr, adv_norm
        if adam_eps:
            self.opt_a.eps = self.opt_c.eps = 1e-5

    def advantages(self, data, gamma, lmbda):
        adv, v_target = super().advantages(data, gamma, lmbda)
        if self.adv_norm:
            adv = (adv - adv.mean()) / (adv.std() + 1e-8)
        return adv, v_target

    def lr_decay(self, episode_num, max_episodes):
        self.opt_a.lr = self.actor_lr * (1 - episode_num / max_episodes)
        self.opt_c.lr = self.critic_lr * (1 - episode_num / max_episodes)

    def beta_params(self, obs):
        h = self.act_actor(F.linear(obs, self.actor["l1.weight"], self.actor["l1.bias"]))
        h = self.act_actor(F.linear(h, self.actor["l2.weight"], self.actor["l2.bias"]))
        alpha = F.softplus(F.linear(h, self.actor["alpha_layer.weight"], self.actor["alpha_layer.bias"])) + 1.0
        beta = F.softplus(F.linear(h, self.actor["beta_layer.weight"], self.actor["beta_layer.bias"])) + 1.0
        return alpha, beta

    def minibatch(self, data, adv, v_target, index, clip_param, entropy_coefficient):
        if not self.beta:
            return super().minibatch(data, adv, v_target, index, clip_param, entropy_coefficient)
        obs, action, reward, next_obs, done, logp_old, adv_dones = data
        alpha, beta = self.beta_params(obs[index])                      # ``:326-333``
        dist = torch.distributions.Beta(alpha, beta)
        ent = dist.entropy().sum(dim=1, keepdim=True)
        logp = dist.log_prob(action[index])
        ratios = torch.exp(logp.sum(dim=1, keepdim=True) - logp_old[index].sum(dim=1, keepdim=True))
        surr1 = ratios * adv[index]
        surr2 = torch.clamp(ratios, 1 - clip_param, 1 + clip_param) * adv[index]
        actor_loss = -torch.min(surr1, surr2).mean() - entropy_coefficient * ent.mean()
        ap = list(self.actor.values())
        ga, _ = clip_grad_norm(torch.autograd.grad(actor_loss, ap), 0.5)
        adam_step(ap, list(ga), self.opt_a)
        v_s = mlp2(self.critic, obs[index], act=self.act)
        critic_loss = F.mse_loss(v_target[index], v_s)
        cp = list(self.critic.values())
        gc, _ = clip_grad_norm(torch.autograd.grad(critic_loss, cp), 0.5)
        adam_step(cp, list(gc), self.opt_c)
        return actor_loss.item(), critic_loss.item()

