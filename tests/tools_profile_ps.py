"""Profiling driver (not a test): tensor-core per-sample Jacobians of the headline machine (for ncu launch lists)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from flowket_b200 import Input, FK_ENGINE_TC
from flowket_b200.machines import ConvNetAutoregressive2D

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
net = ConvNetAutoregressive2D(Input(shape=(10, 10), dtype='int8'), depth=20, num_of_channels=32, seed=0).device_net()
sigma = net.to_sigma(np.random.RandomState(0).choice([-1, 1], size=(B, 10, 10)).astype(np.int8))
for _ in range(2):
    O_re, O_im = net.grad_per_sample(sigma, imag=True, engine=FK_ENGINE_TC)
    torch.cuda.synchronize()
print('ok', float(O_re.norm()), float(O_im.norm()))
