"""The PyTorch custom-op layer (csrc/torch_ops.cpp, TORCH_LIBRARY(cacto, ...)) over the C ABI: north_star's boundary
("a thin C-ABI layer exposed as PyTorch custom ops").  CPU half: the shim library loads, registers every op the
Python mirror uses, and rejects bad arguments with a RuntimeError before anything is launched.  GPU half: an op and the ctypes
binding of the same symbol give the same bits, and errors of the C ABI surface as RuntimeError."""
import numpy as np
import pytest
import torch

from cacto_b200 import _lib, ops as O
from cacto_b200.conf import get_conf


def test_every_op_is_registered_with_a_schema():
    assert int(torch.ops.cacto.abi_version()) == 1
    for name in O.OP_NAMES:
        op = getattr(torch.ops.cacto, name)
        schema = str(op.default._schema)
        assert schema.startswith('cacto::' + name + '('), schema
    # outputs are declared as mutable arguments (the C ABI writes into caller-owned buffers)
    assert '(a!)' in str(torch.ops.cacto.dyn_step.default._schema)
    assert int(torch.ops.cacto.update_tc_workspace_bytes(256, 7, 3)) == int(_lib.lib.cacto_update_tc_workspace_bytes(256, 7, 3))


def test_ops_reject_bad_arguments_without_a_gpu():
    P = O.sys_tensor(_lib.make_sys_params(get_conf('manipulator')))
    assert P.dtype == torch.uint8 and P.numel() == __import__('ctypes').sizeof(_lib.SysParams)
    s, a = torch.zeros((4, 7)), torch.zeros((4, 3))
    # (without a device the CUDA-stream query may fail before the argument checks: either way a RuntimeError, never a result)
    with pytest.raises(RuntimeError):
        torch.ops.cacto.dyn_step(P, 0, s, a, torch.empty_like(s))               # host tensors: no CPU fallback
    with pytest.raises(RuntimeError):
        torch.ops.cacto.dyn_step(P[:-1].clone(), 0, s, a, torch.empty_like(s))  # truncated parameter block
    # element counts are checked by the shim (the C ABI takes bare pointers): an undersized output never reaches a kernel
    with pytest.raises(RuntimeError, match='out has 21 elements, expected 28'):
        torch.ops.cacto.dyn_step(P, 0, s, a, torch.empty((3, 7)))
    with pytest.raises(RuntimeError, match=r'state must be \[B\]\[7\]'):
        torch.ops.cacto.dyn_step(P, 0, torch.zeros((4, 6)), a, torch.empty((4, 6)))
    with pytest.raises(RuntimeError, match='states has'):
        torch.ops.cacto.rollout(P, None, 0, torch.zeros((4, 7), dtype=torch.float64), torch.zeros(4, dtype=torch.int32), 10,
                                torch.empty((10, 7, 4), dtype=torch.float64), torch.empty((10, 3, 4), dtype=torch.float64),
                                torch.empty(4, dtype=torch.int32), None)


@pytest.mark.gpu
def test_op_and_ctypes_binding_agree_bit_for_bit():
    from cacto_b200 import environment as genv
    from cacto_b200._lib import check, lib, ptr, stream_ptr
    conf = get_conf('manipulator')
    env = genv.make_env(conf)
    rng = np.random.default_rng(0)
    s = torch.tensor(rng.uniform(conf.x_init_min, conf.x_init_max, (1000, conf.nb_state)), device='cuda')
    a = torch.tensor(rng.uniform(conf.u_min, conf.u_max, (1000, conf.nb_action)), device='cuda')
    via_op = env.simulate_batch(s, a)
    via_c = torch.empty_like(s)
    check(lib.cacto_dyn_step(env._p, 1, 0, ptr(s), ptr(a), ptr(via_c), 1000, stream_ptr()), 'dyn_step')
    assert torch.equal(via_op, via_c)
    # dtype mismatch between state and action is caught by the shim, a bad size by the C ABI: both raise RuntimeError
    with pytest.raises(RuntimeError, match='dtype'):
        torch.ops.cacto.dyn_step(env._pt, 0, s, a.float(), torch.empty_like(s))
    with pytest.raises(RuntimeError, match='contiguous CUDA tensor'):
        torch.ops.cacto.dyn_step(env._pt, 0, s.cpu(), a, torch.empty_like(s))
    with pytest.raises(RuntimeError, match='sizeof'):
        torch.ops.cacto.dyn_step(env._pt[:-1].clone(), 0, s, a, torch.empty_like(s))
    ws = torch.empty(16, dtype=torch.uint8, device='cuda')
    na_, nc_ = int(lib.cacto_actor_param_count(7, 3)), int(lib.cacto_critic_param_count(7))
    act, cri, st = torch.zeros(na_, device='cuda'), torch.zeros(nc_, device='cuda'), torch.zeros((256, 7), device='cuda')
    with pytest.raises(RuntimeError, match='workspace has 16 elements'):
        torch.ops.cacto.actor_grad_tc(env._pt, act, cri, st, torch.zeros(256, dtype=torch.float64, device='cuda'), 1.0, act.clone(), None, ws)
    with pytest.raises(RuntimeError, match='grad has'):
        torch.ops.cacto.actor_grad_tc(env._pt, act, cri, st, torch.zeros(256, dtype=torch.float64, device='cuda'), 1.0, act[:-1].clone(), None, ws)


@pytest.mark.gpu
def test_ops_run_on_the_current_stream_and_inside_graph_capture():
    from cacto_b200 import environment as genv
    conf = get_conf('car')
    env = genv.make_env(conf)
    s = torch.rand((512, conf.nb_state), device='cuda')
    a = torch.rand((512, conf.nb_action), device='cuda')
    ref = env.simulate_batch(s, a)
    out = torch.empty_like(s)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        torch.ops.cacto.dyn_step(env._pt, 0, s, a, out)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    g = torch.cuda.CUDAGraph()
    out2 = torch.zeros_like(s)
    with torch.cuda.graph(g):
        torch.ops.cacto.dyn_step(env._pt, 0, s, a, out2)
    out2.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out2, ref)
