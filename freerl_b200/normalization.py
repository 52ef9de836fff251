"""Batch_ObsNorm state with the reference's attribute surface (``SAC_file/SAC.py:390-421``, ``DDPG_file/DDPG.py:372-403``,
``MADDPG_file/MADDPG.py:366-397``).

``Normalization_batch_size(shape, device)`` keeps ``running_ms.{n, mean, S, std}``; the numbers live in ONE device
tensor ``state[3, obs_dim]`` = {mean, S, std} that the fused learn kernel (``frl_ac_learn``: ``obs_norm``) updates in
place once per learn and that ``frl_policy_infer`` applies with ``update=False``.  ``running_ms.mean`` / ``.std`` are
``[1, obs_dim]`` views of that tensor, so the train scripts' end-of-run ``np.save`` / ``pickle.dump`` of the statistics
(``SAC.py:586``, ``DDPG.py:578``, ``MADDPG.py:566``) see what the kernel wrote.
"""
import torch


class RunningMeanStd_batch_size:
    def __init__(self, shape, device):
        self.n = 0
        self.state = torch.zeros((3, int(shape)), dtype=torch.float32, device=device)

    @property
    def mean(self):
        return self.state[0].reshape(1, -1)

    @property
    def S(self):
        return self.state[1].reshape(1, -1)

    @property
    def std(self):
        return self.state[2].reshape(1, -1)

    def update(self, x):
        """Welford over BATCH MEANS; the first call sets mean = std = x_bar (reference quirk, SAC.py:402-404).
        Only the public ``sample()`` API comes through here — ``learn()`` does the same update inside the kernel."""
        x = x.mean(dim=0, keepdim=True).reshape(-1)
        self.n += 1
        if self.n == 1:
            self.state[0] = x
            self.state[2] = x
        else:
            old = self.state[0].clone()
            self.state[0] = old + (x - old) / self.n
            self.state[1] = self.state[1] + (x - old) * (x - self.state[0])
            self.state[2] = torch.sqrt(self.state[1] / self.n)


class Normalization_batch_size:
    def __init__(self, shape, device):
        self.running_ms = RunningMeanStd_batch_size(shape, device)

    def __call__(self, x, update=True):
        if update:
            self.running_ms.update(x)
        return (x - self.running_ms.mean) / (self.running_ms.std + 1e-8)

    # ---- kernel plumbing -------------------------------------------------------------------------
    def data_ptr(self):
        return self.running_ms.state.data_ptr()
