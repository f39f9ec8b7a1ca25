"""Probe: per-phase clock64 timeline of one CTA of the tensor-core forward kernel (needs a -DFK_TC_TRACE build)."""
import ctypes
import numpy as np
import torch
from flowket_b200 import _lib
from flowket_b200 import Input, Model, FK_ENGINE_TC
from flowket_b200.machines import ConvNetAutoregressive2D

net = ConvNetAutoregressive2D(Input(shape=(10, 10), dtype='int8'), depth=20, num_of_channels=32, seed=0).device_net()
rng = np.random.RandomState(0)
sigma = net.to_sigma(rng.choice([-1, 1], size=(148 * 3 * 8, 10, 10)).astype(np.int8))
net.log_psi(sigma, engine=FK_ENGINE_TC)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_longlong * (3 * 40 * 16))()
lib.fk_tc_trace_read(buf)
t = np.array(buf[:], dtype=np.int64).reshape(3, 40, 16)[:, :38, :]
t0 = t[:, 0, 0].min()
segs = [('wait1', 0, 2), ('epi1', 2, 3), ('wait2', 3, 5), ('epi2', 5, 6), ('wait3', 6, 8), ('epi3', 8, 9)]
print('mean cycles per segment (blocks 2..35), per pipeline:')
for p in range(3):
    print(p, ' '.join('%s=%5.0f' % (n, (t[p, 2:36, e] - t[p, 2:36, s]).mean()) for n, s, e in segs),
          ' block period=%.0f' % np.diff(t[p, 2:36, 0]).mean())
for p in range(3):
    print(p, 'issue duration ph1=%.0f ph2=%.0f ph3=%.0f ; commit->epilogue sees it: %.0f %.0f %.0f' % tuple(
        [(t[p, 2:36, 12 + k] - t[p, 2:36, 9 + k]).mean() for k in (1, 2, 3)] +
        [(t[p, 2:36, e] - t[p, 2:36, 12 + k]).mean() for k, e in ((1, 2), (2, 5), (3, 8))]))
print('timeline of blocks 10..11 (cycles since start): start, mma1 done, epi1 done, mma2 done, epi2 done, mma3 done, epi3 done')
for b in range(10, 12):
    for p in range(3):
        print(b, p, ' '.join('%6d' % (t[p, b, i] - t0) for i in (0, 2, 3, 5, 6, 8, 9)))
