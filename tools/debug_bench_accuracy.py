"""GPU debug tool: why does bench.py's `accuracy` field disagree (0.51) when smoke() agrees to 6e-7?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from flowket_b200 import Input, Model, FK_ENGINE_FP32, FK_ENGINE_TC
from flowket_b200.machines import ConvNetAutoregressive2D
from flowket_b200.operators import Heisenberg

inp = Input(shape=(10, 10), dtype='int8')
machine = ConvNetAutoregressive2D(inp, depth=20, num_of_channels=32, seed=0)
net = machine.device_net()
op = Heisenberg(hilbert_state_shape=[10, 10], pbc=False)
run = bench.cpu_step(8, seed=2, weights=machine.get_weights())
sg = net.to_sigma(run['sigma'])
want = torch.as_tensor(run['local_values']).to(sg.device)
for count in (True, False):
    got, _, _ = net.local_energy(op.device_desc(), sg, engine=FK_ENGINE_FP32, count=count)
    print('count', count, 'err', float(((got - want).abs().max() / want.abs().max()).item()))
print('want', run['local_values'][:4])
print('got ', got[:4].cpu().numpy())
print('sigma dtype', run['sigma'].dtype, 'unique', np.unique(run['sigma']), 'sum per sample', run['sigma'].reshape(8, -1).sum(1))
from oracle import nets, operators as oops, local_energy as oeloc
spec = nets.Conv2DSpec(10, 10, 20, 32)
p64 = [torch.from_numpy(np.asarray(w, np.float64)) for w in machine.get_weights()]
lv64 = oeloc.local_values(oops.OracleOperator('heisenberg', (10, 10), pbc=False), lambda c: nets.log_psi_numpy(spec, p64, c), run['sigma'].astype(np.float64))
print('oracle fp64', np.asarray(lv64).reshape(-1)[:4])
