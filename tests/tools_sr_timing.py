"""Timing probe (not a test): sample-space stochastic reconfiguration on the headline machine (10x10, depth 20, 32 ch)."""
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
from flowket_b200.optimizers import StochasticReconfiguration

inp = Input(shape=(10, 10), dtype='int8')
machine = ConvNetAutoregressive2D(inp, depth=20, num_of_channels=32, seed=0)
model = Model(inputs=inp, outputs=machine.predictions)
cond = Model(inputs=inp, outputs=machine.conditional_log_probs)
model.engine = cond.engine = FK_ENGINE_TC
obs = Observable(Heisenberg(hilbert_state_shape=[10, 10], pbc=False))
for B in [int(a) for a in sys.argv[1:]] or [1024, 4096]:
    sampler = FastAutoregressiveSampler(cond, B, seed=1)
    sigma = sampler.next_device()
    eloc = obs.local_values_device(model, sigma)
    sr = StochasticReconfiguration(model, diag_shift=0.05, sample_space=True, gram_dtype='bf16', jacobian_chunk=512)
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        delta = sr.compute_update(sigma, eloc.cpu().numpy())
        e1.record()
        torch.cuda.synchronize()
        print(sr.last_timings_ms)
        print('B=%5d rep %d: SR update %.1f ms (jacobian %.1f, solve %.1f)  |delta| = %.4e  peak mem %.1f GB' % (
            B, rep, e0.elapsed_time(e1), sr.last_timings_ms['jacobian'], sr.last_timings_ms['solve'], float(delta.norm()),
            torch.cuda.max_memory_allocated() / 1e9), flush=True)
    del sr, delta
    torch.cuda.empty_cache()
