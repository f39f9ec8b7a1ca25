import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from flowket_b200 import Input, FK_ENGINE_TC, FK_ENGINE_FP32
from flowket_b200.machines import ConvNetAutoregressive2D
for L in (12, 16, 10):
    net = ConvNetAutoregressive2D(Input(shape=(L, L), dtype='int8'), depth=20, num_of_channels=32, seed=0).device_net()
    rng = np.random.RandomState(0)
    n = 65536
    sigma = net.to_sigma(rng.choice([-1, 1], size=(n, L, L)).astype(np.int8))
    a = net.log_psi(sigma[:2048], engine=FK_ENGINE_TC); b = net.log_psi(sigma[:2048], engine=FK_ENGINE_FP32)
    err = (a - b).abs().max().item()
    net.log_psi(sigma, engine=FK_ENGINE_TC); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); net.log_psi(sigma, engine=FK_ENGINE_TC); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts)
    print('L=%d  %.3f M psi/s  %.1f TFLOP/s  max |dlogpsi| vs fp32 %.3e' % (L, n / t / 1e3, n * 2 * 844288 * L * L / (t * 1e-3) / 1e12, err))
