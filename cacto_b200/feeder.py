"""Host feeder: GPU warm-starts streamed into the TO worker pool.

The reference forks ``Pool(nb_cpus)`` workers that each build their own warm-start with ``RL_AC.create_TO_init``
(a B = 1 eager actor forward + one Pinocchio step per knot) and then run the CasADi/ipopt solve (main.py:174-195,
216-233).  CUDA contexts do not survive ``fork``, and the per-worker rollouts are exactly the loop kernel K1 fuses, so
here the PARENT produces the warm-starts of a whole chunk of initial conditions in one launch
(``RL_AC.rollout_to_host``: pinned host buffers, structure-of-arrays) and the workers only ever receive NumPy arrays.
Chunks are double-buffered: while the pool solves chunk k the GPU already rolls out chunk k + 1.  The trajectories cross PCIe in
the compact format by default (``compact=True``: no time row, fp32 controls) and are handed to the workers as the reference's fp64
arrays, rebuilt bit-identically (``RL.CompactRollouts``; tests/test_feeder.py).
"""
import multiprocessing as mp

import numpy as np


def _call(args):
    fn, ics, states, controls, T = args
    return fn(ics, states, controls, T)


class WarmStartFeeder:
    """``to_solve(ICS[ns], init_TO_states[T+1, ns], init_TO_controls[T, na], T)`` is the reference's per-episode TO call
    (``TO_Casadi.TO_Solve``, TO.py:102); it must be picklable and must not touch CUDA.  ``rollout_fn(ics, ep) ->
    (states[T_max+1, ns, B], controls[T_max, na, B], success[B], horizon[B])`` defaults to the fused GPU rollout."""

    def __init__(self, rl, to_solve, nb_cpus=2, chunk=4096, rollout_fn=None, mp_context='fork', compact=True):
        self.rl = rl
        self.compact = bool(compact)                   # compact PCIe format of rollout_to_host (bit-identical arrays, 25 % fewer bytes)
        self.to_solve = to_solve
        self.nb_cpus = int(nb_cpus)
        self.chunk = int(chunk)
        self.rollout_fn = rollout_fn or self._gpu_rollout
        self.ctx = mp.get_context(mp_context)
        self._bufs = {}

    def _gpu_rollout(self, ics, ep):
        import torch
        c = self.rl.conf
        B, T = len(ics), int(c.NSTEPS)
        slot = self._bufs.get('flip', 0)
        self._bufs['flip'] = 1 - slot
        key = (slot, B)
        if key not in self._bufs:                      # two sets of pinned buffers: the pool may still read the previous chunk
            ns_sent, u_dtype = (c.nb_state - 1, torch.float32) if self.compact else (c.nb_state, torch.float64)
            self._bufs[key] = (torch.empty((B, c.nb_state), dtype=torch.float64).pin_memory(),
                               torch.empty((T + 1, ns_sent, B), dtype=torch.float64).pin_memory(),
                               torch.empty((T, c.nb_action, B), dtype=u_dtype).pin_memory(),
                               torch.empty(B, dtype=torch.int32).pin_memory())
        ih, sh, ch, fh = self._bufs[key]
        ih.copy_(torch.as_tensor(np.asarray(ics, dtype=np.float64)))
        hz = self.rl.rollout_to_host(ih, ep, sh, ch, fh, compact=self.compact)
        if self.compact:                               # the reference-shaped per-rollout arrays are rebuilt on the host, bit-identically
            from .RL import CompactRollouts
            return CompactRollouts(c, ih.numpy().copy(), sh, ch, hz), None, fh.numpy(), hz
        return sh.numpy(), ch.numpy(), fh.numpy(), hz

    def _tasks(self, ics, rolled):
        states, controls, ok, hz = rolled
        for i in range(len(ics)):
            T = int(hz[i])
            if T == 0 or not ok[i]:                    # RL.py:202-203, :229-231: no warm-start -> the episode is skipped
                yield None
            elif controls is None:                     # compact transfer: (CompactRollouts, None, ok, horizon)
                yield (self.to_solve, np.array(ics[i]), states.states(i), states.controls(i), T)
            else:
                yield (self.to_solve, np.array(ics[i]), np.ascontiguousarray(states[:T + 1, :, i]),
                       np.ascontiguousarray(controls[:T, :, i]), T)

    def run(self, ICS, ep):
        """Solve every initial condition of ``ICS[E, ns]``; returns the list of ``to_solve`` results in input order
        (``None`` where the reference would have skipped the episode)."""
        ICS = np.asarray(ICS, dtype=np.float64)
        chunks = [ICS[i:i + self.chunk] for i in range(0, len(ICS), self.chunk)]
        results = []
        if not chunks:
            return results
        with self.ctx.Pool(self.nb_cpus) as pool:
            rolled = self.rollout_fn(chunks[0], ep)
            for k, ics in enumerate(chunks):
                tasks = list(self._tasks(ics, rolled))
                live = [t for t in tasks if t is not None]
                pending = pool.map_async(_call, live, chunksize=max(1, len(live) // (4 * self.nb_cpus) or 1))
                if k + 1 < len(chunks):
                    rolled = self.rollout_fn(chunks[k + 1], ep)        # overlaps the pool's work on chunk k
                out = iter(pending.get())
                results.extend(None if t is None else next(out) for t in tasks)
        return results
