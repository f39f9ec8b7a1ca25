"""GPU probe: clock64 timeline of the tensor-core sampler (CTA 0, site (5, 5), all blocks); needs flowket_b200/libflowket_b200_trace.so
built with -DFK_TS_TRACE."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from flowket_b200 import _lib
_lib.load(os.path.join(os.path.dirname(_lib.LIB_PATH), 'libflowket_b200_trace.so'))
from flowket_b200 import Input, FK_ENGINE_TC
from flowket_b200.machines import ConvNetAutoregressive2D
lib = _lib.require_cuda()
net = ConvNetAutoregressive2D(Input(shape=(10, 10), dtype='int8'), depth=20, num_of_channels=32, seed=0).device_net()
for _ in range(2):
    net.sample(1024, seed=3, engine=FK_ENGINE_TC)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (40 * 16))()
lib.fk_ts_trace_read(buf)
t = np.array(buf[:], dtype=np.int64).reshape(40, 16)[:38, :9]
names = ['requests', 'wait w/xa', 'MMA1', 'epi1', 'MMA2', 'epi2', 'wait c', 'MMA3', 'epi3']
d = np.diff(t, axis=1)
print('per-block cycles at site (5,5): mean over blocks 2..36')
for k in range(8):
    print('  %-10s -> %-10s %7.0f' % (names[k], names[k + 1], d[2:37, k].mean()))
print('  block period %.0f cycles' % np.diff(t[2:37, 0]).mean())
