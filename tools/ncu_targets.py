"""One launch of every hot kernel of the north-star step at a profiler-friendly size (ncu replays a kernel ~40 times):
sampler, local energy on both tensor-core engines, bf16 Jacobian rows, Gram, centring, solve, X^T w.
Usage (under gpurun):  ncu --set full --clock-control none --import-source on -k regex:<pattern> -o gpurun_out/<name> \
                           python tools/ncu_targets.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from flowket_b200 import Input, Model, FK_ENGINE_TC, FK_ENGINE_TC_EXACT
from flowket_b200.machines import ConvNetAutoregressive2D
from flowket_b200.operators import Heisenberg
from flowket_b200.optimizers import StochasticReconfiguration

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
inp = Input(shape=(10, 10), dtype='int8')
m = ConvNetAutoregressive2D(inp, depth=20, num_of_channels=32, seed=0)
model = Model(inputs=inp, outputs=m.predictions)
model.engine = FK_ENGINE_TC
net = m.device_net()
op = Heisenberg(hilbert_state_shape=[10, 10], pbc=False)
sg = net.sample(B, seed=3, engine=FK_ENGINE_TC)
e, _, n16 = net.local_energy(op.device_desc(), sg, engine=FK_ENGINE_TC)
ex, _, nx = net.local_energy(op.device_desc(), sg, engine=FK_ENGINE_TC_EXACT)
sr = StochasticReconfiguration(model, sample_space=True)
delta = sr.compute_update(sg, ex)
torch.cuda.synchronize()
print('ncu targets ran: B %d, %d psi evaluations, |delta| %.3e, timings %s' % (B, nx, float(delta.norm()), sr.last_timings_ms))
