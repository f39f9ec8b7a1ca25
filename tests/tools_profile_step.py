"""Profiling driver (not a test): one VMC step at a reduced batch so that ncu replays stay short."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from flowket_b200 import Input, Model, FK_ENGINE_TC, FK_ENGINE_FP32
from flowket_b200.machines import ConvNetAutoregressive2D
from flowket_b200.operators import Heisenberg
from flowket_b200.samplers import FastAutoregressiveSampler
from flowket_b200.observables.monte_carlo import Observable

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
engine = FK_ENGINE_TC if (len(sys.argv) < 3 or sys.argv[2] == 'tc') else FK_ENGINE_FP32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
inp = Input(shape=(10, 10), dtype='int8')
machine = ConvNetAutoregressive2D(inp, depth=20, num_of_channels=32, seed=0)
model = Model(inputs=inp, outputs=machine.predictions)
model.engine = engine
cond = Model(inputs=inp, outputs=machine.conditional_log_probs)
net = machine.device_net()
sampler = FastAutoregressiveSampler(cond, B, seed=1)
obs = Observable(Heisenberg(hilbert_state_shape=[10, 10], pbc=False))
for _ in range(steps):
    sigma = sampler.next_device()
    eloc = obs.local_values_device(model, sigma)
    y = (torch.conj(eloc - eloc.mean()) / B).to(torch.complex64)
    g = net.grad_weighted(net.to_sigma(sigma), y)
    torch.cuda.synchronize()
print('ok', float(eloc.real.mean()))
