"""GPU tool: time fk_jacobian_rows_tc on the headline machine (args: B wn)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flowket_b200 import Input, Model, _lib, FK_ENGINE_TC
from flowket_b200.machines import ConvNetAutoregressive2D
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
wn = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lib = _lib.require_cuda()
inp = Input(shape=(10, 10), dtype='int8')
m = ConvNetAutoregressive2D(inp, depth=20, num_of_channels=32, weights_normalization=bool(wn), seed=0)
net = m.device_net()
sg = net.sample(B, seed=3, engine=FK_ENGINE_TC)
P = net.num_params
nkb = (P + 63) // 64
X = torch.empty((nkb, 2 * B, 64), dtype=torch.bfloat16, device=sg.device)
wsb = lib.fk_jacobian_rows_tc_workspace_bytes(net.handle, B)
ws = torch.empty(wsb, dtype=torch.uint8, device=sg.device)
for it in range(3):
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    _lib.check(lib.fk_jacobian_rows_tc(net.handle, sg.data_ptr(), B, X.data_ptr(), 2 * B, 0, B, ws.data_ptr(), wsb, _lib.stream_ptr()))
    t1.record()
    torch.cuda.synchronize()
    print('jacobian rows B=%d wn=%d: %.2f ms (%.1f us per sample)' % (B, wn, t0.elapsed_time(t1), 1e3 * t0.elapsed_time(t1) / B), flush=True)
