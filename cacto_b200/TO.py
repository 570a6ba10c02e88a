"""Host mirror of the GPU-resident part of the reference's TO.py: ``TO_Casadi.backward_pass`` (TO.py:119-202).

The CasADi / ipopt solves (``TO_System_Solve``, ``TO_Solve``) stay on the host pool of the reference (north_star: out of
scope here); what moves to the GPU is the per-knot Python loop that turns a TO solution into ``dVdx`` -- kernel K6
(``cacto_backward_pass``, csrc/backward.cu).  ``backward_pass`` keeps the reference's signature and return value;
``backward_pass_batch`` processes all TO solutions of an episode batch in one launch pair.
"""
import numpy as np
import torch

from ._lib import check, lib, ptr, stream_ptr
from .segment_tree import _dev


class TO_Casadi:
    def __init__(self, env, conf, env_TO=None, w_S=0):
        self.env = env
        self.conf = conf
        self.nx = conf.nx
        self.nu = conf.na
        self.w_S = w_S
        self.CAMS = env_TO

    def TO_System_Solve(self, *a, **k):
        raise NotImplementedError('the CasADi/ipopt solve stays on the host pool of the reference (TO.py:35-100); '
                                  'cacto_b200 provides its warm-starts (RL_AC.rollout_batch) and consumes its solutions')

    TO_Solve = TO_System_Solve

    def backward_pass_batch(self, TO_states_list, TO_controls_list, mu=1e-9):
        """TO_states_list[e]: [T_e+1, >= nx] (a trailing time column is ignored); TO_controls_list[e]: [T_e, na].
        Returns (V_x, offsets): V_x CUDA fp64 [sum(T_e+1), nx+1] over the concatenated knots (time column 0, TO.py:166)."""
        c = self.conf
        nx, na = int(c.nb_state) - 1, int(c.nb_action)
        dev = _dev()
        lens = np.array([len(s) for s in TO_states_list], dtype=np.int64)
        offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        K, E = int(offsets[-1]), len(lens)
        X = np.zeros((K, nx))
        U = np.zeros((K, na))
        for e, (s, u) in enumerate(zip(TO_states_list, TO_controls_list)):
            s = np.asarray(s, dtype=np.float64)
            u = np.asarray(u, dtype=np.float64).reshape(-1, na)
            if len(u) < len(s) - 1:
                raise ValueError('trajectory %d: %d knots need %d controls, got %d' % (e, len(s), len(s) - 1, len(u)))
            X[offsets[e]:offsets[e + 1]] = s[:, :nx]
            U[offsets[e]:offsets[e] + len(s) - 1] = u[:len(s) - 1]
        Xd, Ud = torch.as_tensor(X).to(dev), torch.as_tensor(U).to(dev)
        off_dev = torch.as_tensor(offsets).to(dev)
        ws = torch.empty(int(lib.cacto_backward_pass_workspace_bytes(nx, na, K)) // 8 + 1, dtype=torch.float64, device=dev)
        Vx = torch.zeros((K, nx + 1), dtype=torch.float64, device=dev)
        check(lib.cacto_backward_pass(self.env._p, ptr(off_dev), E, ptr(Xd), ptr(Ud), K, float(mu), ptr(ws), ptr(Vx), stream_ptr()),
              'backward_pass')
        return Vx, offsets

    def backward_pass(self, T, TO_states, TO_controls, mu=1e-9):
        """TO.py:119: T knots of TO_states[T, nx] and TO_controls[T-1, na] -> V_x[T, nx+1] (NumPy, like the reference)."""
        Vx, _ = self.backward_pass_batch([np.asarray(TO_states)[:T]], [np.asarray(TO_controls)[:T - 1]], mu)
        return Vx.cpu().numpy()
