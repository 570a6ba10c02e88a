"""Reader for the Keras-2.11 ``save_weights`` HDF5 files the reference writes (RL.py:191-195) and ships under
``Results */NNs/`` -- without h5py (not installed here).

Keras stores every variable as a contiguous little-endian float32 dataset; each dataset's object header carries a
version-3 contiguous data-layout message ``08 00 18 00 ?? 00 00 00 03 01 <addr:u64> <size:u64>``.
Datasets appear in object-header order as (kernel, bias) per layer group, groups sorted by name
(``dense_3`` before ``sinusodial_representation_dense*``), so layers are chained by shape starting from the network's
input width.  Kernels are (in, out) row-major, the layout of cacto_b200's parameter blocks.
"""
import re
import struct

import numpy as np

_LAYOUT_MSG = re.compile(rb'\x08\x00\x18\x00.\x00\x00\x00\x03\x01', re.S)


def read_datasets(path):
    """All contiguous float32 datasets of the file, in object-header order, paired as (kernel, bias)."""
    blob = open(path, 'rb').read()
    data = []
    for m in _LAYOUT_MSG.finditer(blob):
        addr, size = struct.unpack('<QQ', blob[m.end():m.end() + 16])
        if 0 < size and addr + size <= len(blob) and size % 4 == 0:
            data.append(np.frombuffer(blob[addr:addr + size], '<f4').copy())
    if len(data) % 2:
        raise ValueError(f'{path}: odd number of datasets ({len(data)}); not a Keras dense-network weight file?')
    return [(data[i], data[i + 1]) for i in range(0, len(data), 2)]


def chain_layers(pairs, fan_in):
    """Order (kernel, bias) pairs into network order starting from ``fan_in`` inputs (backtracking: a 128->1 head and a
    128->128 layer both fit after a 128-wide layer).  Returns [W1, b1, W2, b2, ...] with W reshaped to (in, out)."""
    def rec(rem, width):
        if not rem:
            return []
        for i, (k, b) in enumerate(rem):
            if k.size == width * b.size:
                tail = rec(rem[:i] + rem[i + 1:], b.size)
                if tail is not None:
                    return [k.reshape(width, b.size), b] + tail
        return None
    out = rec(list(pairs), int(fan_in))
    if out is None:
        raise ValueError('cannot chain the layers of the weight file from the given input width')
    return out


def load_keras_weights(path, fan_in):
    """[W1, b1, ...] of a reference ``actor_*.h5`` / ``critic_*.h5`` / ``target_critic_*.h5`` file."""
    return chain_layers(read_datasets(path), fan_in)
