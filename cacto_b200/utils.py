"""utils.py of the reference (state normalisation) on torch tensors / NumPy.

The hot-path kernels fuse ``normalize_tensor`` into every network forward (cacto_b200/csrc/mlp.cuh
``normalize_component``); these host versions exist for API parity with utils.py:4-42.
"""
import numpy as np
import torch


def array2tensor(array):
    return torch.as_tensor(array).unsqueeze(0)


def normalize_tensor(state, state_norm_arr):
    """utils.py:17-24: x / norm for the non-time columns, 2 t / T - 1 for the time (last) column."""
    state = torch.as_tensor(state)
    norm = torch.as_tensor(np.asarray(state_norm_arr, dtype=np.float64), dtype=state.dtype, device=state.device)
    out = state / norm
    out[:, -1] = out[:, -1] * 2 - 1
    return out


def de_normalize_tensor(state, state_norm_arr):
    """utils.py:8-15."""
    state = torch.as_tensor(state)
    norm = torch.as_tensor(np.asarray(state_norm_arr, dtype=np.float64), dtype=state.dtype, device=state.device)
    out = state * norm
    out[:, -1] = (state[:, -1] + 1) * norm[-1] / 2
    return out


def normalize(state, state_norm_arr):
    """utils.py:34-40."""
    out = np.asarray(state, dtype=float) / np.asarray(state_norm_arr, dtype=float)
    out[-1] = out[-1] * 2 - 1
    return out


def de_normalize(state, state_norm_arr):
    """utils.py:26-32."""
    state = np.asarray(state, dtype=float)
    norm = np.asarray(state_norm_arr, dtype=float)
    out = np.empty_like(state)
    out[:-1] = state[:-1] * norm[:-1]
    out[-1] = (state[-1] + 1) * norm[-1] / 2
    return out
