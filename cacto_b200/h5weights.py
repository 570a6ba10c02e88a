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
    with open(path, 'rb') as f:
        blob = f.read()
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


# ------------------------------------------------------------------------------------------------------------------ writer
# Keras-2.11 `model.save_weights('x.h5')` layout (RL.py:191-195), written without h5py as a classic HDF5 file: superblock
# version 0, version-1 object headers, symbol-table groups (v1 B-tree node + SNOD + local heap) -- the structures the reference's
# own files use (e.g. Results */NNs/N_try_*/critic_0.h5, whose tree this mirrors):
#     /                      attrs layer_names [L], backend, keras_version
#     /<layer>/              attr  weight_names ['<layer>/kernel:0', '<layer>/bias:0']
#     /<layer>/<layer>/kernel:0   float32 (in, out), contiguous        /<layer>/<layer>/bias:0   float32 (out,)
# String attributes are fixed-length null-padded (h5py hands them to Keras as bytes, which it decodes); Keras'
# load_weights_from_hdf5_group matches layers by ORDER of `layer_names` among the layers that have weights, so the weightless
# layers (input, LeakyReLU) are omitted.  Not verifiable against h5py here (not installed): checked by a structural parser
# (read_tree, which also walks the reference's files) and by round-tripping through load_keras_weights.
_UNDEF = 0xFFFFFFFFFFFFFFFF
_F32_TYPE = bytes.fromhex('11201f000400000000002000170800177f000000')          # IEEE f32 little-endian (as h5py writes it)


def _pad8(b):
    return b + b'\0' * (-len(b) % 8)


def _msg(mtype, body):
    body = _pad8(body)
    return struct.pack('<HHB3x', mtype, len(body), 0) + body


def _ohdr(msgs):
    data = b''.join(msgs)
    return struct.pack('<BxHII4x', 1, len(msgs), 1, len(data)) + data


def _str_attr(name, values, scalar=False):
    vals = [v.encode() if isinstance(v, str) else bytes(v) for v in values]
    size = max(1, max(len(v) for v in vals))
    dt = struct.pack('<BBBBI', 0x13, 0x01, 0, 0, size)                           # class 3 string, null-padded ASCII
    ds = struct.pack('<BBB5x', 1, 0, 0) if scalar else struct.pack('<BBB5xQ', 1, 1, 0, len(vals))
    nm = name.encode() + b'\0'
    body = struct.pack('<BxHHH', 1, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + b''.join(v.ljust(size, b'\0') for v in vals)
    return _msg(0x0C, body)


class _Group:
    """Sizes and serialisation of one symbol-table group: object header | local heap | B-tree node | symbol node."""
    BTREE, SNOD = 24 + 33 * 8 + 32 * 8, 8 + 8 * 40

    def __init__(self, names, attrs):
        assert len(names) <= 8, 'one symbol node holds at most 2 * leaf K = 8 links'
        self.names = sorted(names)
        self.attrs = attrs
        heap, self.off = b'\0' * 8, {}
        for n in self.names:
            self.off[n] = len(heap)
            heap += _pad8(n.encode() + b'\0')
        self.heap_data = heap
        self.ohdr_size = len(_ohdr([_msg(0x11, b'\0' * 16)] + attrs))
        self.size = self.ohdr_size + 32 + len(heap) + self.BTREE + self.SNOD

    def place(self, addr):
        self.addr = addr
        self.heap = addr + self.ohdr_size
        self.btree = self.heap + 32 + len(self.heap_data)
        self.snod = self.btree + self.BTREE
        return addr + self.size

    def serialise(self, child_addr):
        out = _ohdr([_msg(0x11, struct.pack('<QQ', self.btree, self.heap))] + self.attrs)
        out += b'HEAP' + struct.pack('<B3xQQQ', 0, len(self.heap_data), _UNDEF, self.heap + 32) + self.heap_data
        keys = struct.pack('<Q', 0) + struct.pack('<Q', self.snod) + struct.pack('<Q', self.off[self.names[-1]] if self.names else 0)
        out += (b'TREE' + struct.pack('<BBHQQ', 0, 0, 1, _UNDEF, _UNDEF) + keys).ljust(self.BTREE, b'\0')
        ents = b''.join(struct.pack('<QQII16x', self.off[n], child_addr[n], 0, 0) for n in self.names)
        out += (b'SNOD' + struct.pack('<BxH', 1, len(self.names)) + ents).ljust(self.SNOD, b'\0')
        assert len(out) == self.size
        return out


def save_keras_weights(path, weights, layer_names=None, backend='tensorflow', keras_version='2.11.0'):
    """Write [W1, b1, W2, b2, ...] (kernels (in, out), Keras order) as a Keras-2.11 save_weights HDF5 file.
    ``layer_names``: one name per (kernel, bias) pair; default dense, dense_1, ... (the actor's names in the reference's files)."""
    weights = [np.ascontiguousarray(np.asarray(w, dtype='<f4')) for w in weights]
    assert len(weights) % 2 == 0
    L = len(weights) // 2
    if layer_names is None:
        layer_names = ['dense' if l == 0 else f'dense_{l}' for l in range(L)]
    assert len(set(layer_names)) == L == len(layer_names)
    root = _Group(layer_names, [_str_attr('layer_names', layer_names), _str_attr('backend', [backend], scalar=True),
                                _str_attr('keras_version', [keras_version], scalar=True)])
    outer = {n: _Group([n], [_str_attr('weight_names', [f'{n}/kernel:0', f'{n}/bias:0'])]) for n in layer_names}
    inner = {n: _Group(['kernel:0', 'bias:0'], []) for n in layer_names}

    def dataset_header(arr, data_addr):
        ds = struct.pack('<BBB5x', 1, arr.ndim, 1) + b''.join(struct.pack('<Q', d) for d in arr.shape) * 2          # dims, max dims
        return _ohdr([_msg(0x01, ds), _msg(0x03, _F32_TYPE), _msg(0x05, bytes.fromhex('0202020100000000')),
                      _msg(0x08, struct.pack('<BBQQ', 3, 1, data_addr, arr.nbytes))])
    dsize = {(n, k): len(dataset_header(weights[2 * l + j], 0)) for l, n in enumerate(layer_names) for j, k in enumerate(('kernel:0', 'bias:0'))}
    # ---- addresses: superblock | root group | per layer: outer group, inner group, two dataset headers | raw data
    addr = root.place(96)
    dhdr = {}
    for n in layer_names:
        addr = outer[n].place(addr)
        addr = inner[n].place(addr)
        for k in ('kernel:0', 'bias:0'):
            dhdr[(n, k)] = addr
            addr += dsize[(n, k)]
    daddr = {}
    for l, n in enumerate(layer_names):
        for j, k in enumerate(('kernel:0', 'bias:0')):
            addr = (addr + 7) // 8 * 8
            daddr[(n, k)] = addr
            addr += weights[2 * l + j].nbytes
    eof = addr
    # ---- bytes
    sb = (b'\x89HDF\r\n\x1a\n' + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0) + struct.pack('<QQQQ', 0, _UNDEF, eof, _UNDEF)
          + struct.pack('<QQII', 0, root.addr, 1, 0) + struct.pack('<QQ', root.btree, root.heap))
    assert len(sb) == 96
    blob = bytearray(eof)
    blob[:96] = sb
    blob[root.addr:root.addr + root.size] = root.serialise({n: outer[n].addr for n in layer_names})
    for l, n in enumerate(layer_names):
        blob[outer[n].addr:outer[n].addr + outer[n].size] = outer[n].serialise({n: inner[n].addr})
        blob[inner[n].addr:inner[n].addr + inner[n].size] = inner[n].serialise({k: dhdr[(n, k)] for k in ('kernel:0', 'bias:0')})
        for j, k in enumerate(('kernel:0', 'bias:0')):
            h = dataset_header(weights[2 * l + j], daddr[(n, k)])
            blob[dhdr[(n, k)]:dhdr[(n, k)] + len(h)] = h
            blob[daddr[(n, k)]:daddr[(n, k)] + weights[2 * l + j].nbytes] = weights[2 * l + j].tobytes()
    with open(path, 'wb') as f:
        f.write(bytes(blob))


# ------------------------------------------------------------------------------------------------------------------ structural reader
def read_tree(path):
    """Walk a classic-format HDF5 file the way libhdf5 does (superblock -> root symbol-table entry -> object headers -> B-tree ->
    symbol nodes -> local heap) and return {'attrs': {path: {name: value}}, 'datasets': {path: float32 array}}.  Handles what Keras
    weight files contain: fixed- or variable-length string attributes, contiguous little-endian float32 datasets."""
    with open(path, 'rb') as f:
        b = f.read()
    assert b[:8] == b'\x89HDF\r\n\x1a\n' and b[8] == 0 and b[13] == 8 and b[14] == 8, 'not a superblock-v0 HDF5 file with 8-byte offsets'

    def messages(addr):
        assert b[addr] == 1
        nmsg, = struct.unpack('<H', b[addr + 2:addr + 4])
        hsz, = struct.unpack('<I', b[addr + 8:addr + 12])
        blocks, out = [(addr + 16, addr + 16 + hsz)], []
        while blocks:
            pos, end = blocks.pop(0)
            while pos + 8 <= end and len(out) < nmsg:
                t, sz = struct.unpack('<HH', b[pos:pos + 4])
                body = b[pos + 8:pos + 8 + sz]
                out.append((t, body))
                if t == 0x10:
                    ca, cl = struct.unpack('<QQ', body[:16])
                    blocks.append((ca, ca + cl))
                pos += 8 + sz
        return out

    def heap_name(heap, off):
        assert b[heap:heap + 4] == b'HEAP'
        daddr, = struct.unpack('<Q', b[heap + 24:heap + 32])
        return b[daddr + off:b.index(b'\0', daddr + off)].decode()

    def global_heap_object(addr, index):
        assert b[addr:addr + 4] == b'GCOL'
        pos = addr + 16
        while True:
            idx, _, _, size = struct.unpack('<HHIQ', b[pos:pos + 16])
            if idx == index:
                return b[pos + 16:pos + 16 + size]
            assert idx != 0, 'global heap object not found'
            pos += 16 + (size + 7) // 8 * 8

    def entries(btree, heap):
        out = []

        def node(a):
            assert b[a:a + 4] == b'TREE'
            _, level, used = struct.unpack('<BBH', b[a + 4:a + 8])
            pos = a + 24
            for _ in range(used):
                child, = struct.unpack('<Q', b[pos + 8:pos + 16])
                pos += 16
                if level > 0:
                    node(child)
                else:
                    assert b[child:child + 4] == b'SNOD'
                    n, = struct.unpack('<H', b[child + 6:child + 8])
                    for i in range(n):
                        e = child + 8 + 40 * i
                        lno, oh = struct.unpack('<QQ', b[e:e + 16])
                        out.append((heap_name(heap, lno), oh))
        node(btree)
        return out

    def attr(body):
        ver = body[0]
        nsz, dtsz, dssz = struct.unpack('<HHH', body[2:8])
        pad = (lambda x: (x + 7) // 8 * 8) if ver == 1 else (lambda x: x)
        p = 8
        name = body[p:p + nsz].split(b'\0')[0].decode(); p += pad(nsz)
        dt = body[p:p + dtsz]; p += pad(dtsz)
        ds = body[p:p + dssz]; p += pad(dssz)
        data = body[p:]
        rank = ds[1]
        dims = struct.unpack('<%dQ' % rank, ds[8:8 + 8 * rank]) if rank else ()
        n = int(np.prod(dims)) if rank else 1
        cls, size = dt[0] & 15, struct.unpack('<I', dt[4:8])[0]
        if cls == 3:
            vals = [data[i * size:(i + 1) * size].rstrip(b'\0').decode() for i in range(n)]
        elif cls == 9:                                      # variable-length strings: (length, global heap address, object index)
            vals = []
            for i in range(n):
                ln, ga, gi = struct.unpack('<IQI', data[16 * i:16 * i + 16])
                vals.append(global_heap_object(ga, gi)[:ln].decode())
        else:
            vals = []
        return name, (vals if rank else (vals[0] if vals else None))

    tree = {'attrs': {}, 'datasets': {}}

    def walk(addr, prefix):
        msgs = messages(addr)
        tree['attrs'][prefix or '/'] = dict(attr(body) for t, body in msgs if t == 0x0C)
        st = [body for t, body in msgs if t == 0x11]
        if st:
            bt, hp = struct.unpack('<QQ', st[0][:16])
            for name, oh in entries(bt, hp):
                walk(oh, prefix + '/' + name)
        else:
            ds = [body for t, body in msgs if t == 0x01][0]
            dims = struct.unpack('<%dQ' % ds[1], ds[8:8 + 8 * ds[1]])
            lay = [body for t, body in msgs if t == 0x08][0]
            assert lay[0] == 3 and lay[1] == 1, 'contiguous version-3 layout expected'
            a, size = struct.unpack('<QQ', lay[2:18])
            tree['datasets'][prefix] = np.frombuffer(b[a:a + size], '<f4').reshape(dims).copy()
    root, = struct.unpack('<Q', b[64:72])
    walk(root, '')
    return tree


def load_keras_weights_by_tree(path):
    """[W1, b1, ...] in the order of the file's ``layer_names`` attribute (Keras' own loading rule), via read_tree."""
    t = read_tree(path)
    out = []
    for n in t['attrs']['/']['layer_names']:
        for w in t['attrs'].get('/' + n, {}).get('weight_names', []):
            out.append(t['datasets']['/' + n + '/' + w])
    return out
