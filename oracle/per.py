"""CPU restatement of segment_tree.py and replay_buffer.py -- ORACLE, test infrastructure only.

Pinned against the reference's own segment_tree.py / replay_buffer.py run in the build
container (tests/golden/make_golden.py -> tests/golden/per_*.npz).

Quirks reproduced: Q4 (``sum(0, max_idx-1)`` excludes the last slot),
Q12 irrelevant here; Q1-Q3 are import/attribute bugs that only need the obvious fix
(qualified class names, element-wise leaf gather, 'PER' branch).
"""
import random

import numpy as np


class SegmentTree:
    """segment_tree.py:4-90 -- array-backed tree, root at 1, leaves at [cap, 2cap)."""

    def __init__(self, capacity, op, neutral):
        assert capacity > 0 and capacity & (capacity - 1) == 0
        self.cap = capacity
        self.val = [neutral] * (2 * capacity)
        self.op = op

    def _red(self, start, end, node, lo, hi):               # :36-49
        if start == lo and end == hi:
            return self.val[node]
        mid = (lo + hi) // 2
        if end <= mid:
            return self._red(start, end, 2 * node, lo, mid)
        if mid + 1 <= start:
            return self._red(start, end, 2 * node + 1, mid + 1, hi)
        return self.op(self._red(start, mid, 2 * node, lo, mid), self._red(mid + 1, end, 2 * node + 1, mid + 1, hi))

    def reduce(self, start=0, end=None):                    # :51-74 (note end -= 1)
        if end is None:
            end = self.cap
        if end < 0:
            end += self.cap
        end -= 1
        return self._red(start, end, 1, 0, self.cap - 1)

    def __setitem__(self, idx, v):                          # :76-86
        i = idx + self.cap
        self.val[i] = v
        i //= 2
        while i >= 1:
            self.val[i] = self.op(self.val[2 * i], self.val[2 * i + 1])
            i //= 2

    def __getitem__(self, idx):                             # :88-90
        assert 0 <= idx < self.cap
        return self.val[self.cap + idx]


class SumSegmentTree(SegmentTree):
    def __init__(self, capacity):
        super().__init__(capacity, lambda a, b: a + b, 0.0)

    def sum(self, start=0, end=None):
        return self.reduce(start, end)

    def find_prefixsum_idx(self, prefixsum):                # :105-131
        assert 0 <= prefixsum <= self.sum() + 1e-5
        i = 1
        while i < self.cap:
            if self.val[2 * i] > prefixsum:
                i = 2 * i
            else:
                prefixsum -= self.val[2 * i]
                i = 2 * i + 1
        return i - self.cap


class MinSegmentTree(SegmentTree):
    def __init__(self, capacity):
        super().__init__(capacity, min, float('inf'))

    def min(self, start=0, end=None):
        return self.reduce(start, end)


def _rows(obses_t, rewards, obses_t1, dVdxs, dones, terms):
    # replay_buffer.py:63-72
    cat = [np.concatenate(x, axis=0) for x in (obses_t, rewards, obses_t1, dVdxs, dones, terms)]
    return np.concatenate((cat[0], cat[1].reshape(-1, 1), cat[2], cat[3], cat[4].reshape(-1, 1), cat[5].reshape(-1, 1)), axis=1)


class ReplayBuffer:
    """replay_buffer.py:9-83 (uniform ring buffer)."""

    def __init__(self, conf):
        self.conf = conf
        self.storage_mat = np.zeros((conf.REPLAY_SIZE, 3 * conf.nb_state + 3))
        self.next_idx = 0
        self.full = 0
        self.exp_counter = np.zeros(conf.REPLAY_SIZE)

    def add(self, obses_t, rewards, obses_t1, dVdxs, dones, terms):
        data = _rows(obses_t, rewards, obses_t1, dVdxs, dones, terms)
        R = self.conf.REPLAY_SIZE
        if len(data) + self.next_idx > R:                   # :29-32
            self.storage_mat[self.next_idx:, :] = data[:R - self.next_idx, :]
            self.storage_mat[:self.next_idx + len(data) - R, :] = data[R - self.next_idx:, :]
            self.full = 1
        else:
            self.storage_mat[self.next_idx:self.next_idx + len(data), :] = data
        self._on_add(len(data))
        self.next_idx = (self.next_idx + len(data)) % R

    def _on_add(self, n):
        pass

    def _max_idx(self):
        return self.conf.REPLAY_SIZE if self.full else self.next_idx

    def _gather(self, idx):
        ns = self.conf.nb_state
        m = self.storage_mat
        f32 = np.float32
        return (m[idx, :ns].astype(f32), m[idx, ns:ns + 1].astype(f32), m[idx, ns + 1:2 * ns + 1].astype(f32),
                m[idx, 2 * ns + 1:3 * ns + 1].astype(f32), m[idx, 3 * ns + 1:3 * ns + 2].astype(f32),
                m[idx, 3 * ns + 2:3 * ns + 3])

    def sample(self, idxes=None):
        """:38-61.  ``idxes`` may be injected (quirk Q5: the reference draws them from the
        unseeded global np.random)."""
        if idxes is None:
            idxes = np.random.randint(0, self._max_idx(), size=self.conf.BATCH_SIZE)
        s, r, s1, dv, d, term = self._gather(idxes)
        w = np.ones((self.conf.BATCH_SIZE, 1), dtype=np.float32)
        return s, r, s1, dv, d, term, w, None


class PrioritizedReplayBuffer(ReplayBuffer):
    """replay_buffer.py:87-240."""

    def __init__(self, conf):
        super().__init__(conf)
        self.priorities = np.empty(conf.REPLAY_SIZE)
        assert conf.prioritized_replay_alpha >= 0 and conf.prioritized_replay_beta > 0
        cap = 1
        while cap < conf.REPLAY_SIZE:
            cap *= 2
        self._it_sum = SumSegmentTree(cap)
        self._it_min = MinSegmentTree(cap)
        self._max_priority = 1.0

    def _on_add(self, n):                                   # :133-135
        a = self.conf.prioritized_replay_alpha
        for i in range(n):
            j = (self.next_idx + i) % self.conf.REPLAY_SIZE
            self._it_sum[j] = self._max_priority ** a
            self._it_min[j] = self._max_priority ** a

    def _sample_proportional(self, uniforms=None):          # :139-157
        B = self.conf.BATCH_SIZE
        idx = np.zeros(B)
        p_total = self._it_sum.sum(0, self._max_idx() - 1)
        segment = p_total / B
        for i in range(B):
            u = random.random() if uniforms is None else uniforms[i]
            p = u * segment + i * segment
            idx[i] = self._it_sum.find_prefixsum_idx(p)
        return idx

    def sample(self, uniforms=None):                        # :159-188
        max_idx = self._max_idx()
        beta = self.conf.prioritized_replay_beta
        bi = self._sample_proportional(uniforms).astype(int)
        p_min = self._it_min.min() / self._it_sum.sum()
        max_weight = (p_min * max_idx) ** (-beta)
        self.exp_counter[bi] += 1
        tot = self._it_sum.sum()
        self.priorities[bi] = np.array([self._it_sum[int(i)] for i in bi]) / tot
        w = (self.priorities[bi] * max_idx) ** (-beta) / max_weight
        s, r, s1, dv, d, term = self._gather(bi)
        return s, r, s1, dv, d, term, w.astype(np.float32), bi

    def update_priorities(self, idxes, reward_to_go_batch, critic_value, target_critic_value=None):   # :190-218
        c = self.conf
        if getattr(self, 'RB_type', 'PER') == 'ReLO':                # :193-196 (keras MSE, reduction NONE: mean over the last axis)
            r = np.asarray(reward_to_go_batch, dtype=np.float32)
            td = (((r - np.asarray(critic_value, dtype=np.float32)) ** 2).mean(axis=-1)
                  - ((r - np.asarray(target_critic_value, dtype=np.float32)) ** 2).mean(axis=-1))
            td = np.clip(td, 0, np.max(td))
        else:
            td = np.abs(np.asarray(reward_to_go_batch, dtype=np.float32) - np.asarray(critic_value, dtype=np.float32))[:, 0]
        fresh = c.fresh_factor ** self.exp_counter[idxes]
        new_p = fresh * td + c.prioritized_replay_eps
        assert len(idxes) == len(new_p)
        for i, p in zip(idxes, new_p):
            assert p > 0
            i = int(i)
            self._it_sum[i] = p ** c.prioritized_replay_alpha
            self._it_min[i] = p ** c.prioritized_replay_alpha
            self._max_priority = max(self._max_priority, p)
