"""GPU probe: clock64 timeline of tc_dw_kernel in Jacobian-row mode (needs flowket_b200/libflowket_b200_trace.so, a -DFK_DW_TRACE build)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from flowket_b200 import _lib
_lib.load(os.path.join(os.path.dirname(_lib.LIB_PATH), 'libflowket_b200_trace.so'))
from flowket_b200 import Input, FK_ENGINE_TC
from flowket_b200.machines import ConvNetAutoregressive2D
lib = _lib.require_cuda()
net = ConvNetAutoregressive2D(Input(shape=(10, 10), dtype='int8'), depth=20, num_of_channels=32, seed=0).device_net()
B = 1024
sg = net.to_sigma(np.random.RandomState(0).choice([-1, 1], size=(B, 10, 10)).astype(np.int8))
P = net.num_params
nkb = (P + 63) // 64
X = torch.empty((nkb, 2 * B, 64), dtype=torch.bfloat16, device=sg.device)
wsb = lib.fk_jacobian_rows_tc_workspace_bytes(net.handle, B)
ws = torch.empty(wsb, dtype=torch.uint8, device=sg.device)
for _ in range(2):
    _lib.check(lib.fk_jacobian_rows_tc(net.handle, sg.data_ptr(), B, X.data_ptr(), 2 * B, 0, B, ws.data_ptr(), wsb, _lib.stream_ptr()))
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (64 * 8))()
lib.fk_dw_trace_read(buf)
t = np.array(buf[:], dtype=np.int64).reshape(64, 8)
t0 = t[0, 0]
print('item: issuer start | tfree seen | full seen | issued+committed || drainer: done+sfree seen | drain end || worker 0: sfull seen | row written   (cycles since start)')
for i in range(8, 24):
    print(i, ' '.join('%8d' % (x - t0) for x in t[i, :8]))
d = t[8:60]
print('means: wait tfree %.0f, wait full %.0f, issue %.0f, issue->done seen %.0f, flush %.0f, item period %.0f' % (
    (d[:, 1] - d[:, 0]).mean(), (d[:, 2] - d[:, 1]).mean(), (d[:, 3] - d[:, 2]).mean(), (d[:, 4] - d[:, 3]).mean(),
    (d[:, 5] - d[:, 4]).mean(), np.diff(d[:, 0]).mean()))
