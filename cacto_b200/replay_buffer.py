"""GPU replay buffers with the reference's ``replay_buffer.py`` interface.

Storage (``storage_mat[REPLAY_SIZE, 3*ns+3]`` fp64, rows = [s, partial_rtg, s_next, dVdx, done, term],
replay_buffer.py:20,72) lives in HBM; ``sample`` gathers the drawn rows with one kernel
(``cacto_buffer_gather``) and returns float32 CUDA tensors, the analogue of
``convert_sample_to_tensor`` (replay_buffer.py:74-83).

Bit-exactness of indices and sampled transitions (BASELINE.md parity gate) is kept by leaving on the
host exactly the parts whose bits depend on the host's libraries:
  * the RNG streams -- ``np.random.randint`` for the uniform buffer (replay_buffer.py:45) and
    ``random.random()`` for the PER strata (:153) -- are drawn on the host and shipped to the kernel;
  * ``priority ** alpha`` and the importance weights ``(p * N) ** -beta`` use the host libm ``pow``
    (CUDA's is not correctly rounded); they touch only B scalars per call.
The tree walks (2B O(log N) Python loops + B descents per update in the reference) run in
``cacto_segtree_update`` / ``cacto_segtree_sample``.

The three reference bugs that make PER unusable as shipped (unqualified tree class
names, ndarray passed to ``__getitem__``, ``RB_type`` never set) are fixed the obvious way; quirk Q4
(the newest slot is excluded from the sampled mass) is reproduced because it changes indices.
"""
import random

import numpy as np
import torch

from ._lib import check, lib, ptr, stream_ptr
from .ops import ops
from .segment_tree import MinSegmentTree, SumSegmentTree, _dev


class ReplayBuffer(object):
    """replay_buffer.py:9-83."""

    def __init__(self, conf):
        self.conf = conf
        self.ns = int(conf.nb_state)
        self.width = 3 * self.ns + 3
        self.storage_mat = torch.zeros((conf.REPLAY_SIZE, self.width), dtype=torch.float64, device=_dev())
        self.next_idx = 0
        self.full = 0
        self.exp_counter = np.zeros(conf.REPLAY_SIZE)

    # replay_buffer.py:63-72
    def concatenate_sample(self, obses_t, rewards, obses_t1, dVdxs, dones, terms):
        cat = [np.concatenate([np.asarray(a, dtype=np.float64) for a in x], axis=0) for x in (obses_t, rewards, obses_t1, dVdxs, dones, terms)]
        return np.concatenate((cat[0], cat[1].reshape(-1, 1), cat[2], cat[3], cat[4].reshape(-1, 1), cat[5].reshape(-1, 1)), axis=1)

    def add(self, obses_t, rewards, obses_t1, dVdxs, dones, terms):
        """replay_buffer.py:25-36 (ring write with wrap-around)."""
        data = self.concatenate_sample(obses_t, rewards, obses_t1, dVdxs, dones, terms)
        self.add_rows(torch.as_tensor(data).to(self.storage_mat.device))

    def add_rows(self, data):
        """Same as ``add`` for rows already concatenated ([n, 3ns+3] fp64, host or device)."""
        data = data.to(self.storage_mat.device, torch.float64)
        R, n = self.conf.REPLAY_SIZE, data.shape[0]
        if n + self.next_idx > R:
            self.storage_mat[self.next_idx:, :] = data[:R - self.next_idx, :]
            self.storage_mat[:self.next_idx + n - R, :] = data[R - self.next_idx:, :]
            self.full = 1
        else:
            self.storage_mat[self.next_idx:self.next_idx + n, :] = data
        self._on_add(n)
        self.next_idx = (self.next_idx + n) % R

    def _on_add(self, n):
        pass

    def _max_idx(self):
        return self.conf.REPLAY_SIZE if self.full else self.next_idx

    def _gather(self, idx_dev, out=None):
        n, ns, dev = idx_dev.numel(), self.ns, self.storage_mat.device
        if out is not None:          # pre-allocated tensors of a captured update graph (RL.UpdateGraph.io)
            s, r, s1, dv, d, term = out['state'], out['partial_rtg'], out['state_next'], out['dVdx'], out['done'], out['term']
            assert s.shape == (n, ns)
        else:
            f32 = dict(dtype=torch.float32, device=dev)
            s, r, s1, dv, d = (torch.empty((n, ns), **f32), torch.empty((n, 1), **f32), torch.empty((n, ns), **f32),
                               torch.empty((n, ns), **f32), torch.empty((n, 1), **f32))
            term = torch.empty((n, 1), dtype=torch.float64, device=dev)
        ops.buffer_gather(self.storage_mat, ns, idx_dev, s, r, s1, dv, d, term, None, None)
        return s, r, s1, dv, d, term

    def sample(self, idxes=None, out=None):
        """replay_buffer.py:38-61.  ``idxes`` may be injected (the reference draws them from the global,
        unseeded ``np.random``: quirk Q5); ``out`` = pre-allocated tensors to fill (RL.UpdateGraph.io)."""
        if idxes is None:
            idxes = np.random.randint(0, self._max_idx(), size=self.conf.BATCH_SIZE)
        if isinstance(idxes, torch.Tensor) and idxes.is_cuda:        # a row of draw_indices(): already on the device
            idx_dev = idxes
        else:
            idx_dev = torch.as_tensor(np.asarray(idxes, dtype=np.int64)).to(self.storage_mat.device, non_blocking=True)
        s, r, s1, dv, d, term = self._gather(idx_dev, out)
        if out is not None:
            w_ = out['weights']                                      # importance weights of the uniform buffer: ones, written once --
            if out.get('_ones_version') != w_._version:              # again only if something wrote into the tensor since (version counter)
                w_.fill_(1.0)
                out['_ones_version'] = w_._version
            weights = out['weights']
        else:
            weights = torch.ones((idx_dev.numel(), 1), dtype=torch.float32, device=s.device)
        return s, r, s1, dv, d, term, weights, None


    def draw_indices(self, n_batches):
        """The index draws of ``n_batches`` consecutive ``sample()`` calls (replay_buffer.py:45) as ONE ``np.random.randint`` call and
        one host-to-device copy: int64 [n_batches, BATCH_SIZE] on the device, row k = what the k-th ``sample()`` would have drawn
        (same values, same generator state afterwards: the legacy generator fills a request element by element;
        tests/test_host_random.py).  Pass row k as ``sample(idxes=rows[k], out=...)``.  The buffer must not change in between."""
        idx = np.random.randint(0, self._max_idx(), size=(int(n_batches), int(self.conf.BATCH_SIZE)))
        return torch.as_tensor(np.asarray(idx, dtype=np.int64)).to(self.storage_mat.device, non_blocking=True)

    def index_stream(self, n_batches, chunk_indices=32768):
        """Generator over the rows of ``draw_indices`` for ``n_batches`` consecutive updates, drawn in chunks of about ``chunk_indices``
        indices: a chunk's draw (≈ 10 ns per index) then overlaps the updates of the previous chunk instead of idling the device
        up front (30 x 16 384 indices are 5 ms).  Yields ``None`` rows for a buffer that cannot draw ahead (PER)."""
        per = max(1, int(chunk_indices) // max(1, int(self.conf.BATCH_SIZE)))
        done = 0
        while done < n_batches:
            rows = self.draw_indices(min(per, n_batches - done))
            if rows is None:
                for _ in range(n_batches - done):
                    yield None
                return
            for k in range(rows.shape[0]):
                yield rows[k]
            done += rows.shape[0]


def python_randoms(n):
    """``[random.random() for _ in range(n)]`` (replay_buffer.py:142-147 draws one per stratum) as an array: same generator, same
    bits, same state afterwards.  From 1024 draws on, the interpreter's MT19937 state is advanced by one C call
    (``cacto_host_mt19937_random``) instead of n Python-level calls (B = 4096: the loop was a third of the PER round)."""
    state = random.getstate()
    if n < 1024 or state[0] != 3 or len(state[1]) != 625:
        return np.array([random.random() for _ in range(n)])
    buf = np.fromiter(state[1], dtype=np.uint32, count=625)
    out = np.empty(n, dtype=np.float64)
    check(lib.cacto_host_mt19937_random(buf.ctypes.data, out.ctypes.data, n), 'host_mt19937_random')
    random.setstate((state[0], tuple(buf.tolist()), state[2]))
    return out


class PrioritizedReplayBuffer(ReplayBuffer):
    """replay_buffer.py:87-240."""

    def __init__(self, conf):
        super().__init__(conf)
        self.priorities = np.empty(conf.REPLAY_SIZE)
        assert conf.prioritized_replay_alpha >= 0
        assert conf.prioritized_replay_beta > 0
        it_capacity = 1
        while it_capacity < conf.REPLAY_SIZE:
            it_capacity *= 2
        self._capacity = it_capacity
        self._it_sum = SumSegmentTree(it_capacity)
        self._it_min = MinSegmentTree(it_capacity)
        self._max_priority = 1.0
        self.RB_type = 'PER'
        dev = self.storage_mat.device
        self._stamp = torch.full((it_capacity,), -1, dtype=torch.int32, device=dev)
        self._totals = torch.zeros(3, dtype=torch.float64, device=dev)

    def draw_indices(self, n_batches):
        """Prioritized draws depend on the priorities the previous update wrote: nothing to draw ahead."""
        return None

    def _tree_update(self, idx, val):
        dev = self.storage_mat.device
        idx = idx if isinstance(idx, torch.Tensor) else torch.as_tensor(np.asarray(idx, dtype=np.int64)).to(dev, non_blocking=True)
        val = torch.as_tensor(np.asarray(val, dtype=np.float64)).to(dev, non_blocking=True)
        ops.segtree_update(self._it_sum._value, self._it_min._value, self._capacity, idx, val, self._stamp)

    def _on_add(self, n):
        """replay_buffer.py:133-135: new rows enter with max_priority ** alpha in both trees."""
        R = self.conf.REPLAY_SIZE
        idx = (self.next_idx + np.arange(n)) % R
        v = self._max_priority ** self.conf.prioritized_replay_alpha
        self._tree_update(idx, np.full(n, v, dtype=np.float64))

    def _sample_proportional(self, uniforms=None):
        """replay_buffer.py:139-157.  Returns device idx, device leaf values; both and the totals (sum(0, max_idx-1), sum(), min())
        live in ONE packed device block so that sample() brings them to the host with a single copy."""
        B = self.conf.BATCH_SIZE
        if uniforms is None:
            uniforms = python_randoms(B)
        dev = self.storage_mat.device
        u = torch.as_tensor(np.asarray(uniforms, dtype=np.float64)).to(dev, non_blocking=True)
        pk = getattr(self, '_pack', None)
        if pk is None or pk[0].numel() != 2 * B + 4:
            pk = (torch.empty(2 * B + 4, dtype=torch.float64, device=dev), torch.empty(2 * B + 4, dtype=torch.float64).pin_memory())
            self._pack = pk
        pack = pk[0]
        idx, leaf, totals = pack[:B].view(torch.int64), pack[B:2 * B], pack[2 * B:2 * B + 3]
        ops.segtree_sample(self._it_sum._value, self._it_min._value, self._capacity, self._max_idx(), u, idx, leaf, totals)
        return idx, leaf

    def sample(self, uniforms=None, out=None):
        """replay_buffer.py:159-188."""
        max_idx = self._max_idx()
        beta = self.conf.prioritized_replay_beta
        B = self.conf.BATCH_SIZE
        idx_dev, leaf_dev = self._sample_proportional(uniforms)
        s, r, s1, dv, d, term = self._gather(idx_dev, out)
        # ONE small D2H (B indices + B leaves + 3 totals, packed); the pow() below must be the host's (bit-exactness)
        dev_pack, host_pack = self._pack
        host_pack.copy_(dev_pack, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        hp = host_pack.numpy()
        batch_idxes = hp[:B].view(np.int64).astype(int)
        leaf = hp[B:2 * B].copy()
        _, tot, mn = hp[2 * B:2 * B + 3]
        self._last_idx = (batch_idxes, idx_dev.clone())        # update_priorities of the same batch re-uses the device copy
        p_min = mn / tot
        max_weight = (p_min * max_idx) ** (-beta)
        self.exp_counter[batch_idxes] += 1
        self.priorities[batch_idxes] = leaf / tot
        weights = (self.priorities[batch_idxes] * max_idx) ** (-beta) / max_weight
        weights = torch.as_tensor(weights.astype(np.float32)).to(s.device, non_blocking=True)
        if out is not None:
            out['weights'].copy_(weights.reshape(-1, 1))
            weights = out['weights']
        return s, r, s1, dv, d, term, weights, batch_idxes

    def update_priorities(self, idxes, reward_to_go_batch, critic_value, target_critic_value=None):
        """replay_buffer.py:190-218.  ``RB_type`` 'PER' (default: |rtg - V|) or 'ReLO' (:193-196: MSE(rtg, V) - MSE(rtg, V_target)
        per sample, clipped to [0, max])."""
        c = self.conf
        idxes_in = idxes
        rtg = torch.as_tensor(reward_to_go_batch).reshape(-1, 1)
        V = torch.as_tensor(critic_value).reshape(-1, 1)
        if self.RB_type == 'ReLO':
            Vt = torch.as_tensor(target_critic_value).reshape(-1, 1)
            r32 = rtg.to(torch.float32)
            td = ((r32 - V.to(torch.float32)) ** 2).mean(dim=-1) - ((r32 - Vt.to(torch.float32).to(r32.device)) ** 2).mean(dim=-1)
            td = td.cpu().numpy()
            td = np.clip(td, 0, np.max(td))
        else:
            td = torch.abs(rtg.to(torch.float32) - V.to(torch.float32))[:, 0].cpu().numpy()
        idxes = np.asarray(idxes).astype(int)
        fresh = c.fresh_factor ** self.exp_counter[idxes]
        new_p = fresh * td + c.prioritized_replay_eps
        assert len(idxes) == len(new_p)
        assert (new_p > 0).all()
        # p ** alpha with the host libm's pow (what the reference's Python float ** evaluates), one C call for the batch
        new_p = np.ascontiguousarray(new_p, dtype=np.float64)
        vals = np.empty_like(new_p)
        check(lib.cacto_host_pow(new_p.ctypes.data, float(c.prioritized_replay_alpha), vals.ctypes.data, new_p.size), 'host_pow')
        last = getattr(self, '_last_idx', None)
        if last is not None and last[0] is idxes_in:              # the indices sample() returned: their device copy is still there
            self._tree_update(last[1], vals)
        else:
            self._tree_update(idxes, vals)
        self._max_priority = max(self._max_priority, float(new_p.max()))
