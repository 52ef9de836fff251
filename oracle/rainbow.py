"""CPU restatement (PyTorch fp32 + autograd) of the Rainbow DQN learn step.  TEST INFRASTRUCTURE ONLY (see oracle/algos.py).

Follows ``DQN_file/DQN_with_tricks.py:81-160`` (Categorical: dueling logits, softmax over 51 atoms, expected value,
``projection_dist`` with ``index_add_``), ``:242-284`` (learn with PER weights, Double selection, n-step gamma) and
``DQN_file/Noisy_net.py:17-76`` (factorised noise ``f(x) = sign(x) sqrt|x|``, resampled on every forward: V then A,
``randn(in)`` then ``randn(out)``).  Noise is an explicit argument: ``eps[f] = (V_in, V_out, A_in, A_out)`` raw
``torch.randn`` draws of forward f (0: online on s' [Double], 1: target on s', 2: online on s).
Pinned by ``tests/golden/rainbow.npz`` (generated from the reference by ``oracle/make_golden_rainbow.py``).
"""
from collections import OrderedDict

import torch

from .algos import AdamState, adam_step, clone_net, polyak, _leaf


def f_noise(x):
    return x.sign() * torch.sqrt(abs(x))


def noisy_linear(net, name, x, eps_in_raw, eps_out_raw):
    ei, ej = f_noise(eps_in_raw), f_noise(eps_out_raw)
    w = net[name + ".weight_mu"] + net[name + ".weight_sigma"] * torch.ger(ej, ei)
    b = net[name + ".bias_mu"] + net[name + ".bias_sigma"] * ej
    return torch.nn.functional.linear(x, w, b)


def predict(net, obs, eps, n_actions, n_atoms, z):
    """``Categorical._predict`` with Noisy + Dueling: returns (dist [B,A,Z], q [B,A])."""
    x = torch.relu(torch.nn.functional.linear(obs, net["l1.weight"], net["l1.bias"]))
    V = noisy_linear(net, "V", x, eps[0], eps[1]).reshape(-1, 1, n_atoms)
    A = noisy_linear(net, "A", x, eps[2], eps[3]).reshape(-1, n_actions, n_atoms)
    logits = V + A - A.mean(dim=1, keepdim=True)
    dist = torch.softmax(logits.reshape(-1, n_actions, n_atoms), dim=2)
    return dist, (dist * z).sum(dim=2)


class RainbowOracle:
    def __init__(self, qnet, lr, n_actions, n_atoms=51, v_min=-100.0, v_max=100.0):
        trainable = OrderedDict((k, v) for k, v in qnet.items() if "epsilon" not in k)
        self.q = _leaf(trainable)
        self.q_target = clone_net(trainable)
        self.opt = AdamState(list(self.q.values()), lr)
        self.nA, self.nZ, self.v_min, self.v_max = n_actions, n_atoms, v_min, v_max
        self.z = torch.linspace(v_min, v_max, steps=n_atoms)
        self.delta_z = (v_max - v_min) / (n_atoms - 1)

    def learn(self, batch, eps, gamma, tau, is_weight=None, double_q=True):
        obs, act, rew, nobs, done = batch
        B = obs.shape[0]
        offset = torch.linspace(0, (B - 1) * self.nZ, B).reshape(-1, 1).long()
        with torch.no_grad():
            if double_q:
                _, q1 = predict(self.q, nobs, eps[0], self.nA, self.nZ, self.z)
                na = torch.argmax(q1, dim=1)
            dist_t, q_t = predict(self.q_target, nobs, eps[1], self.nA, self.nZ, self.z)
            if not double_q:
                na = torch.argmax(q_t, dim=1)
            next_dist = dist_t[torch.arange(B), na]
            t_z = (rew + gamma * self.z * (1 - done)).clamp(self.v_min, self.v_max)
            b = (t_z - self.v_min) / self.delta_z
            l, u = b.floor().long(), b.ceil().long()
            m = torch.zeros(B, self.nZ)
            m.reshape(-1).index_add_(0, (l + offset).reshape(-1), ((u + (l == u) - b) * next_dist).reshape(-1))
            m.reshape(-1).index_add_(0, (u + offset).reshape(-1), ((b - l) * next_dist).reshape(-1))
        dist_all, _ = predict(self.q, obs, eps[2], self.nA, self.nZ, self.z)
        dist = dist_all[torch.arange(B), act.reshape(-1).long()]
        logp = dist.clamp(1e-5, 1 - 1e-5).log()
        error = (m * logp).sum(1)
        if is_weight is not None:
            loss = -(m * logp * is_weight.reshape(-1, 1)).sum(1).mean()
        else:
            loss = -(m * logp).sum(1).mean()
        params = list(self.q.values())
        grads = torch.autograd.grad(loss, params)
        adam_step(params, grads, self.opt)
        polyak(self.q_target, self.q, tau)
        return {"loss": loss.item(), "error": error.detach().clone(), "m": m}
