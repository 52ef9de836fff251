"""CPU restatement (PyTorch fp32 + autograd) of the NON-distributional branch of ``DQN_with_tricks.DQN.learn``.
TEST INFRASTRUCTURE ONLY (see oracle/algos.py).

Follows ``DQN_file/DQN_with_tricks.py:40-79`` (``MLP`` / ``Dueling``: ``Q = V + A - A.mean(dim=1, keepdim=True)``) and
``:261-283``: Double selection (``Qnet(s').argmax`` evaluated by ``Qnet_target``), n-step gamma, PER.  Reference quirk kept:
with PER ``loss = (is_weight * td_error ** 2).mean()`` multiplies ``is_weight [B]`` with ``td_error [B, 1]`` and therefore
averages the ``[B, B]`` outer product (= mean(w) * mean(td^2)); priorities come from ``td_error = Q(s,a) - y``.
Pinned by ``tests/golden/dqn_tricks_*.npz`` (generated from the reference by ``oracle/make_golden_dqn_tricks.py``).
"""
import torch
import torch.nn.functional as F

from .algos import AdamState, adam_step, clone_net, polyak, _leaf


def q_values(net, obs, dueling):
    x = torch.relu(F.linear(obs, net["l1.weight"], net["l1.bias"]))
    if not dueling:
        return F.linear(x, net["l2.weight"], net["l2.bias"])
    V = F.linear(x, net["V.weight"], net["V.bias"])
    A = F.linear(x, net["A.weight"], net["A.bias"])
    return V + A - A.mean(dim=1, keepdim=True)


class DQNTricksOracle:
    def __init__(self, qnet, lr, dueling=False, double_q=False):
        self.q = _leaf(qnet)
        self.q_target = clone_net(qnet)
        self.opt = AdamState(list(self.q.values()), lr)
        self.dueling, self.double_q = dueling, double_q

    def learn(self, batch, gamma, tau, is_weight=None):
        """``gamma`` is ``buffer.n_step_gamma`` when N_Step is on (``:269-270``)."""
        obs, act, rew, nobs, done = batch
        with torch.no_grad():
            if self.double_q:
                na = q_values(self.q, nobs, self.dueling).argmax(dim=1).reshape(-1, 1)
                next_q = q_values(self.q_target, nobs, self.dueling).gather(1, na.long())
            else:
                next_q = q_values(self.q_target, nobs, self.dueling).max(dim=1)[0].reshape(-1, 1)
            target = rew + gamma * next_q * (1 - done)
        cur = q_values(self.q, obs, self.dueling).gather(1, act.long())
        if is_weight is not None:
            td = cur - target
            loss = (is_weight * (td ** 2)).mean()            # [B] * [B,1] -> [B,B] (reference broadcast, :277)
        else:
            td = cur - target
            loss = F.mse_loss(cur, target)
        params = list(self.q.values())
        grads = torch.autograd.grad(loss, params)
        adam_step(params, grads, self.opt)
        polyak(self.q_target, self.q, tau)
        return {"loss": loss.item(), "td_error": td.detach().clone()}
