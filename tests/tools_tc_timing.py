"""Timing probe (not a test): tc_log_psi throughput vs n, with CUDA events."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from flowket_b200 import Input, Model, FK_ENGINE_TC
from flowket_b200.machines import ConvNetAutoregressive2D
inp = Input(shape=(10, 10), dtype='int8')
machine = ConvNetAutoregressive2D(inp, depth=20, num_of_channels=32, seed=0)
net = machine.device_net()
rng = np.random.RandomState(0)
for n in [296, 2960, 27464, 29600, 65536, 27464, 131072, 262144, 65536]:
    sigma = net.to_sigma(rng.choice([-1, 1], size=(n, 10, 10)).astype(np.int8))
    net.log_psi(sigma, engine=FK_ENGINE_TC)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); net.log_psi(sigma, engine=FK_ENGINE_TC); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts)
    print('n=%7d  %8.3f ms  %.3f M cfg/s  %.1f TFLOP/s (algorithmic)  iters/CTA=%.1f  us/iter=%.1f' % (
        n, t, n / t / 1e3, n * 168.8576e6 / (t * 1e-3) / 1e12, n / 2 / 148, t * 1e3 / np.ceil(n / 2 / 148)))
