"""Profiling driver (not a test): one pass of each tensor-core kernel at a reduced batch (for ncu)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from flowket_b200 import Input, Model, FK_ENGINE_TC
from flowket_b200.machines import ConvNetAutoregressive2D
from flowket_b200.operators import Heisenberg
from flowket_b200.samplers import FastAutoregressiveSampler
from flowket_b200.observables.monte_carlo import Observable

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
inp = Input(shape=(10, 10), dtype='int8')
machine = ConvNetAutoregressive2D(inp, depth=20, num_of_channels=32, seed=0)
model = Model(inputs=inp, outputs=machine.predictions)
cond = Model(inputs=inp, outputs=machine.conditional_log_probs)
model.engine = cond.engine = FK_ENGINE_TC
net = machine.device_net()
sampler = FastAutoregressiveSampler(cond, B, seed=1)
obs = Observable(Heisenberg(hilbert_state_shape=[10, 10], pbc=False))
for _ in range(2):
    sigma = sampler.next_device()
    eloc = obs.local_values_device(model, sigma)
    y = (torch.conj(eloc - eloc.mean()) / B).to(torch.complex64)
    g = net.grad_weighted(net.to_sigma(sigma), y, engine=FK_ENGINE_TC)
    torch.cuda.synchronize()
print('ok', float(eloc.real.mean()), float(g.norm()))
