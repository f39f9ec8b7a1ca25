"""GPU tool: launches of fk_sr_gram_xxt on a synthetic panel-major X (for ncu / timing sweeps).
args: R K reps"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flowket_b200 import _lib
lib = _lib.require_cuda()
R = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
K = int(sys.argv[2]) if len(sys.argv) > 2 else 106752
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device('cuda:0')
nkb = (K + 63) // 64
X = torch.empty((nkb, R, 64), dtype=torch.bfloat16, device=dev)
step = max(1, nkb // 16)
for k in range(0, nkb, step):
    X[k:k + step] = (torch.randn((min(step, nkb - k), R, 64), device=dev) * 0.3).to(torch.bfloat16)
G = torch.empty((R, R), dtype=torch.float32, device=dev)
wsb = lib.fk_sr_gram_xxt_workspace_bytes(R)
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
for i in range(reps):
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    _lib.check(lib.fk_sr_gram_xxt(X.data_ptr(), R, K, R, 1, 0, 1.0, G.data_ptr(), R, ws.data_ptr(), wsb, _lib.stream_ptr()))
    t1.record()
    torch.cuda.synchronize()
    nt = (R + 255) // 256
    ms = t0.elapsed_time(t1)
    print('R %d K %d super %s: %.2f ms, %.0f TFLOP/s (triangle)' % (R, K, os.environ.get('FK_GRAM2_SUPER', '8'), ms, 2.0 * (nt * (nt + 1) / 2) * 65536 * K / ms / 1e9), flush=True)
