"""Minimal pure-Python reader for the Keras HDF5 weight files FlowKet ships (experiments/weights/*.h5) and
writes with `model.save_weights` (callbacks/checkpoint.py, experiments/train.py:128-129).

h5py is not a dependency.  Supported subset (what Keras 2.1.6-tf / h5py 2.x produce for weight files): superblock
version 0, version-1 object headers (with continuation blocks), symbol-table groups (v1 B-tree + local heap),
contiguous little-endian float32/float64 datasets of any rank.  Anything else raises ValueError."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class _H5(object):
    def __init__(self, path):
        with open(path, 'rb') as f:
            self.b = f.read()
        if self.b[:8] != b'\x89HDF\r\n\x1a\n':
            raise ValueError('%s is not an HDF5 file' % path)
        if self.b[8] != 0 or self.b[13] != 8 or self.b[14] != 8:
            raise ValueError('unsupported HDF5 superblock (need version 0 with 8-byte offsets)')
        self.root_header = struct.unpack('<Q', self.b[64:72])[0]

    # ---- object headers --------------------------------------------------------------------------------
    def messages(self, addr):
        """[(type, payload bytes)] of a version-1 object header, following continuation messages."""
        b = self.b
        if b[addr] != 1:
            raise ValueError('unsupported object header version %d' % b[addr])
        nmsg = struct.unpack('<H', b[addr + 2:addr + 4])[0]
        size = struct.unpack('<I', b[addr + 8:addr + 12])[0]
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            pos, remaining = blocks.pop(0)
            end = pos + remaining
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize = struct.unpack('<HH', b[pos:pos + 4])
                payload = b[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x0010:   # continuation
                    caddr, clen = struct.unpack('<QQ', payload[:16])
                    blocks.append((caddr, clen))
                out.append((mtype, payload))
        return out

    # ---- groups ------------------------------------------------------------------------------------------
    def group_entries(self, btree_addr, heap_addr):
        b = self.b
        assert b[heap_addr:heap_addr + 4] == b'HEAP'
        heap_data = struct.unpack('<Q', b[heap_addr + 24:heap_addr + 32])[0]
        res = []

        def walk(addr):
            assert b[addr:addr + 4] == b'TREE', 'bad B-tree node'
            node_type, level, used = b[addr + 4], b[addr + 5], struct.unpack('<H', b[addr + 6:addr + 8])[0]
            assert node_type == 0
            pos = addr + 24          # after left/right siblings
            pos += 8                 # key 0
            for _ in range(used):
                child = struct.unpack('<Q', b[pos:pos + 8])[0]
                pos += 16            # child pointer + next key
                if level > 0:
                    walk(child)
                else:
                    assert b[child:child + 4] == b'SNOD'
                    n = struct.unpack('<H', b[child + 6:child + 8])[0]
                    for i in range(n):
                        e = child + 8 + 40 * i
                        name_off, hdr = struct.unpack('<QQ', b[e:e + 16])
                        s = heap_data + name_off
                        name = b[s:b.index(b'\x00', s)].decode()
                        res.append((name, hdr))
        walk(btree_addr)
        return res

    def walk(self, addr, prefix=''):
        """yield (path, ndarray) for every dataset below the object header at `addr`"""
        msgs = self.messages(addr)
        stab = [p for t, p in msgs if t == 0x0011]
        if stab:
            btree, heap = struct.unpack('<QQ', stab[0][:16])
            for name, hdr in self.group_entries(btree, heap):
                for item in self.walk(hdr, prefix + '/' + name):
                    yield item
            return
        shape = dtype = data_addr = None
        for t, p in msgs:
            if t == 0x0001:      # dataspace
                version, rank = p[0], p[1]
                off = 8 if version == 1 else 4
                shape = struct.unpack('<%dQ' % rank, p[off:off + 8 * rank])
            elif t == 0x0003:    # datatype
                cls = p[0] & 0x0F
                size = struct.unpack('<I', p[4:8])[0]
                if cls != 1 or size not in (4, 8) or (p[1] & 1):
                    raise ValueError('unsupported datatype (class %d, size %d)' % (cls, size))
                dtype = np.dtype('<f%d' % size)
            elif t == 0x0008:    # data layout
                version = p[0]
                if version == 3:
                    if p[1] != 1:
                        raise ValueError('only contiguous datasets are supported (layout class %d)' % p[1])
                    data_addr = struct.unpack('<Q', p[2:10])[0]
                elif version in (1, 2):
                    rank, cls = p[1], p[2]
                    if cls != 1:
                        raise ValueError('only contiguous datasets are supported')
                    data_addr = struct.unpack('<Q', p[8:16])[0]
                else:
                    raise ValueError('unsupported layout version %d' % version)
        if shape is None or dtype is None or data_addr is None:
            return
        n = int(np.prod(shape)) if len(shape) else 1
        if data_addr == UNDEF:
            arr = np.zeros(shape, dtype)
        else:
            arr = np.frombuffer(self.b, dtype=dtype, count=n, offset=data_addr).reshape(shape).copy()
        yield prefix, arr


def read_h5_datasets(path):
    """{'/group/.../name': ndarray} for every dataset in the file"""
    h = _H5(path)
    return dict(h.walk(h.root_header))


def read_keras_weights(path, weight_specs):
    """Weights in the order of `machine.weight_specs()` (names like 'weight_normalization_3/kernel:0');
    Keras stores them at '/<layer>/<layer>/<weight>' ."""
    data = read_h5_datasets(path)
    out = []
    for name, shape, _ in weight_specs:
        layer, weight = name.split('/')
        key = '/%s/%s/%s' % (layer, layer, weight)
        if key not in data:
            cands = [k for k in data if k.endswith('/' + name)]
            if len(cands) != 1:
                raise KeyError('weight %s not found in %s' % (name, path))
            key = cands[0]
        arr = data[key]
        if tuple(arr.shape) != tuple(shape):
            raise ValueError('weight %s: file has shape %s, machine expects %s' % (name, arr.shape, shape))
        out.append(arr.astype(np.float32))
    return out
