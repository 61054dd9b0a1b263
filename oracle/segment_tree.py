"""CPU oracle (test infrastructure): array-backed sum / min segment trees and the
prioritized-replay sampling arithmetic.

Restates the semantics of ``baselines/common/segment_tree.py`` (reference):
* node array of 2*capacity float64, leaves at [capacity, 2*capacity)  (:33)
* ``set(i, v)`` writes the leaf then recomputes every ancestor as op(left, right) (:76-86)
* ``reduce(start, end)``: ``end`` exclusive, ``None`` -> capacity, negative wraps by
  ``+capacity`` (:69-74); the reference folds sub-ranges top-down, left operand first
* ``find_prefixsum_idx(p)``: descend from the root, go LEFT iff ``value[2i] > p`` else
  subtract and go right (:124-131)

Implementation here is iterative over a numpy float64 array (the reference recurses
over a python list); additions are performed in the same left-to-right association as
the reference's top-down decomposition so sums are bit-identical.
"""
import numpy as np


class _Tree:
    def __init__(self, capacity, op, neutral):
        assert capacity > 0 and capacity & (capacity - 1) == 0, "capacity must be a power of two"
        self.capacity = capacity
        self.op = op
        self.neutral = float(neutral)
        self.value = np.full(2 * capacity, neutral, dtype=np.float64)

    def set(self, idx, val):
        i = int(idx) + self.capacity
        v = self.value
        v[i] = val
        i >>= 1
        while i >= 1:
            v[i] = self.op(v[2 * i], v[2 * i + 1])
            i >>= 1

    def get(self, idx):
        assert 0 <= idx < self.capacity
        return float(self.value[self.capacity + idx])

    def _fold(self, lo, hi, node, nlo, nhi):
        # same decomposition (and association) as segment_tree.py:36-49
        if lo == nlo and hi == nhi:
            return float(self.value[node])
        mid = (nlo + nhi) // 2
        if hi <= mid:
            return self._fold(lo, hi, 2 * node, nlo, mid)
        if mid + 1 <= lo:
            return self._fold(lo, hi, 2 * node + 1, mid + 1, nhi)
        return self.op(self._fold(lo, mid, 2 * node, nlo, mid),
                       self._fold(mid + 1, hi, 2 * node + 1, mid + 1, nhi))

    def reduce(self, start=0, end=None):
        if end is None:
            end = self.capacity
        if end < 0:
            end += self.capacity
        end -= 1
        return self._fold(start, end, 1, 0, self.capacity - 1)


class SumTree(_Tree):
    def __init__(self, capacity):
        super().__init__(capacity, lambda a, b: a + b, 0.0)

    def sum(self, start=0, end=None):
        return self.reduce(start, end)

    def find_prefixsum_idx(self, prefixsum):
        assert 0 <= prefixsum <= self.sum() + 1e-5
        i = 1
        v = self.value
        p = float(prefixsum)
        while i < self.capacity:
            left = v[2 * i]
            if left > p:
                i = 2 * i
            else:
                p -= left
                i = 2 * i + 1
        return i - self.capacity


class MinTree(_Tree):
    def __init__(self, capacity):
        super().__init__(capacity, min, float("inf"))

    def min(self, start=0, end=None):
        return self.reduce(start, end)


class PrioritizedSampler:
    """The index / weight arithmetic of ``PrioritizedReplayBuffer``
    (``baselines/deepq/replay_buffer.py:71-191``) without the python-object storage.

    * capacity rounded up to a power of two (:92-94)
    * ``add`` writes ``max_priority ** alpha`` to both trees at the ring slot (:100-105)
    * ``sample_idx(uniforms)``: ``p_total = sum(0, len-1)`` -- NOTE the reference's
      exclusive-end call drops the last stored element (:109); stratified mass
      ``u*range + i*range`` with ``range = p_total / batch`` (:110-112)
    * ``weights``: ``p_min = min()/sum()``; ``max_w = (p_min*n)**-beta``;
      ``w_i = (p_i/sum() * n)**-beta / max_w`` (:157-165), float64
    * ``update_priorities``: ``p**alpha`` into both trees, running max (:169-191)
    """

    def __init__(self, size, alpha):
        assert alpha >= 0
        self.maxsize = int(size)
        self.alpha = float(alpha)
        cap = 1
        while cap < size:
            cap *= 2
        self.sum_tree = SumTree(cap)
        self.min_tree = MinTree(cap)
        self.max_priority = 1.0
        self.next_idx = 0
        self.n = 0

    def add(self):
        idx = self.next_idx
        self.n = min(self.n + 1, self.maxsize)
        self.next_idx = (self.next_idx + 1) % self.maxsize
        p = self.max_priority ** self.alpha
        self.sum_tree.set(idx, p)
        self.min_tree.set(idx, p)
        return idx

    def sample_idx(self, uniforms):
        batch = len(uniforms)
        p_total = self.sum_tree.sum(0, self.n - 1)
        every = p_total / batch
        return [self.sum_tree.find_prefixsum_idx(float(uniforms[i]) * every + i * every)
                for i in range(batch)]

    def weights(self, idxes, beta):
        assert beta > 0
        total = self.sum_tree.sum()
        p_min = self.min_tree.min() / total
        max_w = (p_min * self.n) ** (-beta)
        out = []
        for i in idxes:
            p = self.sum_tree.get(i) / total
            out.append((p * self.n) ** (-beta) / max_w)
        return np.array(out, dtype=np.float64)

    def update_priorities(self, idxes, priorities):
        assert len(idxes) == len(priorities)
        for i, p in zip(idxes, priorities):
            assert p > 0
            assert 0 <= i < self.n
            v = float(p) ** self.alpha
            self.sum_tree.set(i, v)
            self.min_tree.set(i, v)
            self.max_priority = max(self.max_priority, float(p))
