"""Batched n-step reward-to-go (kernel K5) -- the arithmetic of RL_AC.RL_Solve (RL.py:173-187).

``rtg_batch`` processes a ragged batch of TO trajectories in one launch of ``cacto_rtg_window``;
``RL.RL_AC.RL_Solve`` calls it with a single trajectory to keep the reference's method signature.
"""
import numpy as np
import torch

from .ops import ops
from .segment_tree import _dev


def rtg_batch(conf, states_list, step_cost_list, lengths=None):
    """states_list[e]: [T_e+1, ns] (TO_states with the time column, TO.py:114-115);
    step_cost_list[e]: [T_e+1] TO step costs (reward = -cost, RL.py:168).  Either argument may instead be a tensor over the
    concatenated knots (states [sum(T_e+1), ns] or [E, T+1, ns]; costs [E, T+1], or 1-D with ``lengths``).
    Returns a dict of CUDA fp64 tensors over the concatenated knots plus ``offsets`` (host int64):
    partial, total, state_next, done, term, rwrd, ep_return[E]."""
    ns = int(conf.nb_state)
    dev = _dev()
    if isinstance(step_cost_list, torch.Tensor):
        # pre-concatenated costs: a 1-D tensor over all knots needs explicit lengths; a 2-D [E, T+1] tensor is E equal-length trajectories
        if step_cost_list.dim() == 2:
            lens = np.full(step_cost_list.shape[0], step_cost_list.shape[1], dtype=np.int64)
        elif lengths is not None:
            lens = np.asarray(lengths, dtype=np.int64).reshape(-1)
        else:
            raise ValueError('rtg_batch: a 1-D cost tensor over concatenated knots needs lengths=[T_e + 1, ...]')
        rwrd = (-step_cost_list.to(dev, torch.float64)).reshape(-1).contiguous()
    else:
        lens = np.array([len(c) for c in step_cost_list], dtype=np.int64)
        rwrd = torch.as_tensor(-np.concatenate([np.asarray(c, dtype=np.float64).reshape(-1) for c in step_cost_list])).to(dev)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    if isinstance(states_list, torch.Tensor):
        states = states_list.to(dev, torch.float64).reshape(-1, ns).contiguous()
    else:
        states = torch.as_tensor(np.concatenate([np.asarray(s, dtype=np.float64).reshape(-1, ns) for s in states_list], axis=0)).to(dev)
    if int(offsets[-1]) != states.shape[0] or rwrd.numel() != states.shape[0]:
        raise ValueError('rtg_batch: %d states, %d costs, lengths sum to %d' % (states.shape[0], rwrd.numel(), int(offsets[-1])))
    total_knots = int(offsets[-1])
    E = len(lens)
    off_dev = torch.as_tensor(offsets).to(dev)
    f64 = dict(dtype=torch.float64, device=dev)
    out = dict(partial=torch.empty(total_knots, **f64), total=torch.empty(total_knots, **f64),
               state_next=torch.empty((total_knots, ns), **f64), done=torch.empty(total_knots, **f64),
               term=torch.empty(total_knots, **f64), ep_return=torch.empty(E, **f64), rwrd=rwrd, states=states, offsets=offsets)
    ops.rtg_window(off_dev, rwrd, states, ns, int(getattr(conf, 'nsteps_TD_N', 0)), int(bool(conf.MC)), out['partial'], out['total'], out['state_next'],
                   out['done'], out['term'], out['ep_return'])
    return out
