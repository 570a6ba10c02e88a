"""Host side of the compact transfer format (RL.CompactRollouts): the reference's per-rollout arrays rebuilt from the buffers
``rollout_to_host(compact=True)`` fills.  No GPU needed; the bit-identity with the full-format transfer is a GPU test
(tests/test_gpu_rollout.py::test_rollout_to_host_compact_is_bit_reconstructible)."""
from types import SimpleNamespace

import numpy as np
import torch


def test_compact_rollouts_rebuild_time_column_and_widen_controls():
    from cacto_b200.RL import CompactRollouts
    rng = np.random.default_rng(0)
    T, ns, na, B, dt = 12, 5, 2, 7, 0.005
    conf = SimpleNamespace(dt=dt)
    ics = rng.normal(size=(B, ns))
    ics[:, -1] = dt * rng.integers(0, T, B)
    hz = rng.integers(0, T + 1, B).astype(np.int32)
    st = torch.as_tensor(rng.normal(size=(T + 1, ns - 1, B)))
    ct = torch.as_tensor(rng.normal(size=(T, na, B)).astype(np.float32))
    view = CompactRollouts(conf, torch.as_tensor(ics), st, ct, hz)
    for b in range(B):
        Tb = int(hz[b])
        s, u = view.states(b), view.controls(b)
        assert s.shape == (Tb + 1, ns) and u.shape == (Tb, na) and s.dtype == np.float64 and u.dtype == np.float64
        assert s.flags['C_CONTIGUOUS'] and u.flags['C_CONTIGUOUS']
        assert np.array_equal(s[:, :-1], st.numpy()[:Tb + 1, :, b])
        t = ics[b, -1]
        for k in range(Tb + 1):                                 # t_{k+1} = t_k + dt, one addition after the other (what the kernels do)
            assert s[k, -1] == t
            t = t + dt
        assert np.array_equal(u, ct.numpy()[:Tb, :, b].astype(np.float64))
