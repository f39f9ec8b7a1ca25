"""Timing probe (not a test): fk_sample throughput vs batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flowket_b200 import Input, Model
from flowket_b200.machines import ConvNetAutoregressive2D
inp = Input(shape=(10, 10), dtype='int8')
machine = ConvNetAutoregressive2D(inp, depth=20, num_of_channels=32, seed=0)
net = machine.device_net()
import sys
engine = 1 if (len(sys.argv) > 1 and sys.argv[1] == 'tc') else 0
for B in [1024, 4096, 8192, 16384, 32768]:
    net.sample(B, seed=1, engine=engine)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); net.sample(B, seed=2, engine=engine); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts)
    print('B=%6d  %8.2f ms  %8.0f samples/s  %.2f TFLOP/s' % (B, t, B / t * 1e3, B * 168.8576e6 / (t * 1e-3) / 1e12))
