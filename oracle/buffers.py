"""CPU restatement (numpy) of the reference replay / rollout buffers.

TEST INFRASTRUCTURE ONLY — this module is the *checker*.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu-baseline / ``--impl reference`` legs may import it; nothing under
``freerl_b200/`` does.

Parity status: the reference ships no tests or golden vectors for this path, so
the restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the build container by
``oracle/make_golden.py`` and committed under ``tests/golden/`` (``tests/test_oracle_golden.py``).

Every class cites the reference ``file:line`` it follows (paths relative to the reference root).
The restatement is written independently (flat record arrays, explicit heap helpers) — it is not a
copy of the reference source.
"""
from collections import deque

import numpy as np


class RingReplay:
    """Ring replay of (obs, action, reward, next_obs, done).

    Follows ``SAC_file/Buffer.py:11-61`` (identical copies in DQN/TD3/DDPG/MADDPG): float64 storage,
    bool dones, write cursor ``_index`` wrapping ``% capacity``, ``_size`` saturating at capacity;
    ``sample(indices)`` gathers rows, casts to float32 and shapes reward / done as ``[B, 1]``.
    Returned values are numpy arrays (the reference wraps the same numbers in torch tensors).
    """

    def __init__(self, capacity, obs_dim, act_dim):
        self.capacity = int(capacity)                       # Buffer.py:15 (float 1e6 allowed)
        self.obs_dim, self.act_dim = int(obs_dim), int(act_dim)
        self.obs = np.zeros((self.capacity, self.obs_dim), np.float64)
        self.actions = np.zeros((self.capacity, self.act_dim), np.float64)
        self.rewards = np.zeros(self.capacity, np.float64)
        self.next_obs = np.zeros((self.capacity, self.obs_dim), np.float64)
        self.dones = np.zeros(self.capacity, bool)
        self._index = 0
        self._size = 0

    def add(self, obs, action, reward, next_obs, done):    # Buffer.py:29-38
        i = self._index
        self.obs[i] = obs
        self.actions[i] = action
        self.rewards[i] = reward
        self.next_obs[i] = next_obs
        self.dones[i] = done
        self._index = (i + 1) % self.capacity
        self._size = min(self._size + 1, self.capacity)

    def sample(self, indices):                              # Buffer.py:40-57
        idx = np.asarray(indices)
        f32 = np.float32
        return (self.obs[idx].astype(f32), self.actions[idx].astype(f32),
                self.rewards[idx].astype(f32).reshape(-1, 1), self.next_obs[idx].astype(f32),
                self.dones[idx].astype(f32).reshape(-1, 1))

    def __len__(self):
        return self._size


def uniform_indices(total_size, batch_size):
    """``np.random.choice(total, B, replace=False)`` on the legacy global stream.

    ``DQN_file/DQN.py:97``, ``SAC_file/SAC.py:213``, ``TD3_file/TD3.py:183``, ``DDPG_file/DDPG.py:194``,
    ``MADDPG_file/MADDPG.py:188``.  (Equivalent to ``np.random.permutation(total)[:B]``; SURVEY App. B.)
    """
    return np.random.choice(total_size, batch_size, replace=False)


class SumTreeOracle:
    """Array-heap sum tree, ``DQN_file/Buffer.py:134-194``.

    ``tree`` has ``2*cap-1`` float64 slots, leaf ``i`` lives at ``i + cap - 1``; there is NO power-of-two
    padding so for general capacities leaves sit on two depths.  ``update`` propagates the *difference*
    (``+= change``) up to the root — ancestors are not re-summed, which fixes the last-ulp values.
    """

    def __init__(self, capacity):
        self.capacity = int(capacity)
        self.tree = np.zeros(2 * self.capacity - 1, np.float64)

    def set_leaf(self, buffer_index, priority):             # Buffer.py:150-166 (add + update)
        node = int(buffer_index) + self.capacity - 1
        priority = float(np.asarray(priority).reshape(-1)[0])   # fp32 (1,) arrays widen exactly to float64
        change = priority - self.tree[node]
        self.tree[node] = priority
        while node != 0:
            node = (node - 1) // 2
            self.tree[node] += change

    def find(self, s):                                      # Buffer.py:168-188
        node, n = 0, self.tree.shape[0]
        while 2 * node + 1 < n:
            left = 2 * node + 1
            if s <= self.tree[left]:
                node = left
            else:
                s = s - self.tree[left]
                node = left + 1
        return self.tree[node], node - self.capacity + 1

    def total(self):                                        # Buffer.py:190-191
        return self.tree[0]

    def max_leaf(self):                                     # Buffer.py:193-194
        return np.max(self.tree[-self.capacity:])


class PrioritizedReplay:
    """Proportional PER, ``DQN_file/Buffer.py:66-132`` (defaults alpha .5, beta .4, +.001, eps .01,
    probability floor 1e-7).  ``prob_floor`` is 1e-10 in the SAC/TD3/DDPG/MADDPG copies
    (``SAC_file/Buffer.py:115``)."""

    def __init__(self, capacity, obs_dim, act_dim, alpha=0.5, beta=0.4, beta_increment=0.001,
                 epsilon=0.01, prob_floor=1e-7):
        self.capacity = int(capacity)
        self.alpha, self.beta, self.beta_increment, self.epsilon = alpha, beta, beta_increment, epsilon
        self.prob_floor = prob_floor
        self.sumtree = SumTreeOracle(self.capacity)
        self.buffer = RingReplay(self.capacity, obs_dim, act_dim)

    def add(self, obs, action, reward, next_obs, done):    # Buffer.py:91-97
        p = 1.0 if len(self.buffer) == 0 else self.sumtree.max_leaf()
        self.sumtree.set_leaf(self.buffer._index, p)
        self.buffer.add(obs, action, reward, next_obs, done)

    def sample(self, batch_size):                           # Buffer.py:99-124
        idx = np.zeros(batch_size, np.int64)
        pri = np.zeros(batch_size, np.float32)              # NOTE float32 container (Buffer.py:104)
        seg = self.sumtree.total() / batch_size
        self.beta = np.min([1.0, self.beta + self.beta_increment])
        for i in range(batch_size):
            s = np.random.uniform(seg * i, seg * (i + 1))
            pri[i], idx[i] = self.sumtree.find(s)
        prob = pri / self.sumtree.total()                   # f32 / f64 scalar -> f64 under NumPy 2 (NEP 50)
        prob = np.clip(prob, self.prob_floor, None)
        w = (len(self.buffer) * prob) ** (-self.beta)
        w = w / w.max()
        return idx, w.astype(np.float32)

    def update_priorities(self, indices, td_error):         # Buffer.py:126-129
        pri = (np.abs(td_error) + self.epsilon) ** self.alpha
        for i, p in zip(indices, pri):
            self.sumtree.set_leaf(i, p)

    def __len__(self):
        return len(self.buffer)


def fold_n_step(window, gamma):
    """Fold a full n-step window into one transition, ``DQN_file/Buffer.py:261-269`` (=350-358).

    ``window`` is a sequence of (obs, action, reward, next_obs, done) oldest→newest.  Start from the
    newest reward/next_obs/done; walking back ``R = r_i + gamma*R*(1-d_i)`` and a done at step i
    replaces (next_obs, done) by step i's.  Python float64 arithmetic, as in the reference.
    """
    obs, action = window[0][0], window[0][1]
    reward, next_obs, done = window[-1][2], window[-1][3], window[-1][4]
    for i in range(len(window) - 2, -1, -1):
        _, _, r, n_o, d = window[i]
        reward = r + gamma * reward * (1 - d)
        if d:
            next_obs, done = n_o, d
    return obs, action, reward, next_obs, done


class NStepReplay:
    """``N_Step_Buffer``: ``DQN_file/Buffer.py:199-293`` (n_step default 2).  The deque is never reset at
    episode end and once full every add emits one folded transition."""

    def __init__(self, capacity, obs_dim, act_dim, gamma, n_step=2):
        self.ring = RingReplay(capacity, obs_dim, act_dim)
        self.gamma, self.n_step = gamma, n_step
        self.n_step_gamma = gamma ** n_step
        self.window = deque(maxlen=n_step)

    def add(self, obs, action, reward, next_obs, done):
        self.window.append((obs, action, reward, next_obs, done))
        if len(self.window) == self.n_step:
            self.ring.add(*fold_n_step(self.window, self.gamma))

    def sample(self, indices):
        return self.ring.sample(indices)

    def __len__(self):
        return len(self.ring)


class NStepPrioritizedReplay(PrioritizedReplay):
    """``N_Step_PER_Buffer``: ``DQN_file/Buffer.py:333-399`` (n_step default 3)."""

    def __init__(self, capacity, obs_dim, act_dim, alpha=0.5, beta=0.4, beta_increment=0.001,
                 epsilon=0.01, gamma=None, n_step=3):
        super().__init__(capacity, obs_dim, act_dim, alpha, beta, beta_increment, epsilon)
        self.gamma, self.n_step = gamma, n_step
        self.n_step_gamma = gamma ** n_step
        self.window = deque(maxlen=n_step)

    def add(self, obs, action, reward, next_obs, done):
        self.window.append((obs, action, reward, next_obs, done))
        if len(self.window) == self.n_step:
            super().add(*fold_n_step(self.window, self.gamma))


class RolloutStore:
    """``Buffer_for_PPO``: ``PPO_file/Buffer.py:266-323`` (= ``MAPPO_file/Buffer.py:266-323``).

    ``all()`` returns the FULL capacity arrays (not just ``_size`` rows) cast to float32.
    """

    def __init__(self, capacity, obs_dim, act_dim, logp_dim=None):
        self.capacity = int(capacity)
        c = self.capacity
        self.obs = np.zeros((c, obs_dim))
        self.actions = np.zeros((c, act_dim))
        self.rewards = np.zeros(c)
        self.next_obs = np.zeros((c, obs_dim))
        self.dones = np.zeros(c, bool)
        self.action_log_probs = np.zeros((c, act_dim if logp_dim is None else logp_dim))
        self.adv_dones = np.zeros(c, bool)
        self._index = 0
        self._size = 0

    def add(self, obs, action, reward, next_obs, done, action_log_probs, adv_done):
        i = self._index
        self.obs[i] = obs
        self.actions[i] = action
        self.rewards[i] = reward
        self.next_obs[i] = next_obs
        self.dones[i] = done
        self.action_log_probs[i] = action_log_probs
        self.adv_dones[i] = adv_done
        self._index = (i + 1) % self.capacity
        self._size = min(self._size + 1, self.capacity)

    def clear(self):
        self._index = 0
        self._size = 0

    def all(self):
        f32 = np.float32
        return (self.obs.astype(f32), self.actions.astype(f32), self.rewards.astype(f32).reshape(-1, 1),
                self.next_obs.astype(f32), self.dones.astype(f32).reshape(-1, 1),
                self.action_log_probs.astype(f32), self.adv_dones.astype(f32).reshape(-1, 1))

    def __len__(self):
        return self._size
