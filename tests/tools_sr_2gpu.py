"""Probe (torchrun, 2 ranks, NCCL): distributed complex SR on J1J2 6x6 == the single-process update on the same global batch.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/tools_sr_2gpu.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from flowket_b200 import Input, Model
from flowket_b200.machines import ComplexValuesSimpleConvNetAutoregressive1D
from flowket_b200.operators import J1J2
from flowket_b200.samplers import FastAutoregressiveSampler
from flowket_b200.optimization import VariationalMonteCarlo, DistributedVariationalMonteCarlo
from flowket_b200.optimizers import ComplexValuesStochasticReconfiguration

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl')
B = 256


def build(batch, offset, distributed):
    inp = Input(shape=(36,), dtype='int8')
    m = ComplexValuesSimpleConvNetAutoregressive1D(inp, depth=5, num_of_channels=16, seed=0)
    model = Model(inp, m.predictions)
    cond = Model(inp, m.conditional_log_probs)
    op = J1J2(hilbert_state_shape=[6, 6], j2=0.5, pbc=False)
    sampler = FastAutoregressiveSampler(cond, batch, seed=11, sample_offset=offset)
    vmc = (DistributedVariationalMonteCarlo if distributed else VariationalMonteCarlo)(model, op, sampler)
    sr = ComplexValuesStochasticReconfiguration(model, diag_shift=0.05, iterative_solver=False, distributed=distributed)
    return vmc, sr


vmc, sr = build(B, rank * B, True)
sigma, y = vmc.next_batch()
delta = sr.compute_update(sigma, y)
gathered = [torch.empty_like(delta) for _ in range(world)]
dist.all_gather(gathered, delta)
if rank == 0:
    assert all(torch.equal(g, gathered[0]) or (g - gathered[0]).abs().max() < 1e-6 for g in gathered), 'ranks disagree'
    vmc1, sr1 = build(B * world, 0, False)
    sigma1, y1 = vmc1.next_batch()
    want = sr1.compute_update(sigma1, y1)
    rel = float((delta - want).abs().max() / want.abs().max())
    print('distributed SR (2 x %d samples, NCCL) vs single process (%d samples): max rel diff %.3e, energy %.6f vs %.6f' % (
        B, B * world, rel, vmc.current_energy.real, vmc1.current_energy.real), flush=True)
    assert rel < 2e-3
dist.barrier()
dist.destroy_process_group()
