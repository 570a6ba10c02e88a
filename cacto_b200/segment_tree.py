"""GPU segment trees with the reference's ``segment_tree.py`` interface.

``SegmentTree`` keeps its 2*capacity fp64 nodes in HBM (root at 1, leaves at [capacity, 2*capacity),
exactly the reference's array, segment_tree.py:31-34).  Writes go through the level-synchronous kernel
``cacto_segtree_update``; reductions and prefix-sum searches through ``cacto_segtree_reduce`` /
``cacto_segtree_find``.  Every internal node is ``left + right`` (or ``min``) in fp64 round-to-nearest,
the reference's reduction order, so node values are bit-identical to the reference's.

Scalar ``tree[idx] = v`` / ``tree[idx]`` / ``sum()`` / ``min()`` / ``find_prefixsum_idx`` mirror the
reference; ``set_batch`` / ``find_prefixsum_idx_batch`` are the batched entry points the replay buffer
uses (one launch per batch instead of 2B Python tree walks, replay_buffer.py:210-216).
"""
import numpy as np
import torch

from ._lib import check, lib, ptr, stream_ptr
from .ops import ops


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError('cacto_b200 needs a CUDA device (no CPU fallback on the hot path)')
    return torch.device('cuda', torch.cuda.current_device())


class SegmentTree(object):
    """segment_tree.py:4-90.  ``kind`` is 'sum' or 'min'."""

    def __init__(self, capacity, kind, neutral_element):
        assert capacity > 0 and capacity & (capacity - 1) == 0, "capacity must be positive and a power of 2."
        self._capacity = capacity
        self._kind = kind
        self._value = torch.full((2 * capacity,), float(neutral_element), dtype=torch.float64, device=_dev())
        self._stamp = torch.full((capacity,), -1, dtype=torch.int32, device=_dev())
        self._out = torch.zeros(2, dtype=torch.float64, device=_dev())

    def _ptrs(self):
        return (ptr(self._value), ptr(None)) if self._kind == 'sum' else (ptr(None), ptr(self._value))

    def set_batch(self, idx, val):
        """tree[idx[i]] = val[i] for all i, in order (the last duplicate wins)."""
        idx = torch.as_tensor(np.asarray(idx, dtype=np.int64) if not isinstance(idx, torch.Tensor) else idx).to(_dev(), torch.int64).contiguous()
        val = torch.as_tensor(np.asarray(val, dtype=np.float64) if not isinstance(val, torch.Tensor) else val).to(_dev(), torch.float64).contiguous()
        ops.segtree_update(self._value if self._kind == 'sum' else None, self._value if self._kind == 'min' else None, self._capacity, idx, val, self._stamp)

    def reduce(self, start=0, end=None):
        """segment_tree.py:51-74."""
        if end is None:
            end = self._capacity
        s, m = self._ptrs()
        check(lib.cacto_segtree_reduce(s, m, self._capacity, int(start), int(end), ptr(self._out), stream_ptr()), 'segtree_reduce')
        return float(self._out[0 if self._kind == 'sum' else 1])

    def __setitem__(self, idx, val):
        self.set_batch([int(idx)], [float(val)])

    def __getitem__(self, idx):
        assert 0 <= idx < self._capacity
        return float(self._value[self._capacity + idx])

    def leaves(self, idx):
        """Element-wise leaf read for an index array (the fix of quirk Q2, replay_buffer.py:175)."""
        idx = torch.as_tensor(np.asarray(idx, dtype=np.int64)).to(_dev())
        return self._value[self._capacity + idx]


class SumSegmentTree(SegmentTree):
    """segment_tree.py:93-131."""

    def __init__(self, capacity):
        super(SumSegmentTree, self).__init__(capacity=capacity, kind='sum', neutral_element=0.0)

    def sum(self, start=0, end=None):
        return super(SumSegmentTree, self).reduce(start, end)

    def find_prefixsum_idx_batch(self, prefixsums):
        p = torch.as_tensor(np.asarray(prefixsums, dtype=np.float64)).to(_dev()).contiguous()
        out = torch.empty(p.numel(), dtype=torch.int64, device=p.device)
        check(lib.cacto_segtree_find(ptr(self._value), self._capacity, ptr(p), p.numel(), ptr(out), stream_ptr()), 'segtree_find')
        return out

    def find_prefixsum_idx(self, prefixsum):
        assert 0 <= prefixsum <= self.sum() + 1e-5
        return int(self.find_prefixsum_idx_batch([prefixsum])[0])


class MinSegmentTree(SegmentTree):
    """segment_tree.py:134-145."""

    def __init__(self, capacity):
        super(MinSegmentTree, self).__init__(capacity=capacity, kind='min', neutral_element=float('inf'))

    def min(self, start=0, end=None):
        return super(MinSegmentTree, self).reduce(start, end)
