"""Probe: the TC kernel must be bitwise deterministic run-to-run (a race would show up here)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from flowket_b200 import Input, Model, FK_ENGINE_TC
from flowket_b200.machines import ConvNetAutoregressive2D
inp = Input(shape=(10, 10), dtype='int8')
machine = ConvNetAutoregressive2D(inp, depth=20, num_of_channels=32, seed=0)
net = machine.device_net()
rng = np.random.RandomState(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
sigma = net.to_sigma(rng.choice([-1, 1], size=(n, 10, 10)).astype(np.int8))
ref = net.log_psi(sigma, engine=FK_ENGINE_TC).clone()
bad = 0
for i in range(int(sys.argv[2]) if len(sys.argv) > 2 else 10):
    out = net.log_psi(sigma, engine=FK_ENGINE_TC)
    bad += int((torch.view_as_real(out) != torch.view_as_real(ref)).any())
print('nondeterministic runs:', bad, 'finite:', bool(torch.isfinite(torch.view_as_real(ref)).all()))
