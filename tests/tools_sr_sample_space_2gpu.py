"""Probe (torchrun, G ranks, NCCL): sample-space SR with the batch sharded over the ranks (all-to-all re-shard of the
Jacobian rows, partial Grams on every GPU, allreduce, replicated Cholesky) == the single-process update on the same
global batch, for the real-parameter 2-D machine.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 \\
      tests/tools_sr_sample_space_2gpu.py [--depth 20 --batch 1024 --lattice 10]"""
import argparse
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from flowket_b200 import Input, Model, FK_ENGINE_TC
from flowket_b200.machines import ConvNetAutoregressive2D
from flowket_b200.operators import Heisenberg
from flowket_b200.samplers import FastAutoregressiveSampler
from flowket_b200.optimization import VariationalMonteCarlo, DistributedVariationalMonteCarlo
from flowket_b200.optimizers import StochasticReconfiguration

ap = argparse.ArgumentParser()
ap.add_argument('--depth', type=int, default=4)
ap.add_argument('--batch', type=int, default=128, help='samples per rank')
ap.add_argument('--lattice', type=int, default=6)
ap.add_argument('--skip_single', action='store_true', help='timing only (the global batch may not fit one GPU)')
ap.add_argument('--shared_cholesky', action='store_true', help='also time the solve with the factorisation shared between the ranks')
args = ap.parse_args()
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl')
L, B = args.lattice, args.batch


def build(batch, offset, distributed):
    inp = Input(shape=(L, L), dtype='int8')
    m = ConvNetAutoregressive2D(inp, depth=args.depth, num_of_channels=32, seed=0)
    model, cond = Model(inp, m.predictions), Model(inp, m.conditional_log_probs)
    model.engine = cond.engine = FK_ENGINE_TC
    op = Heisenberg(hilbert_state_shape=[L, L], pbc=False)
    sampler = FastAutoregressiveSampler(cond, batch, seed=11, sample_offset=offset)
    vmc = (DistributedVariationalMonteCarlo if distributed else VariationalMonteCarlo)(model, op, sampler)
    sr = StochasticReconfiguration(model, diag_shift=0.05, sample_space=True, gram_dtype='bf16', distributed=distributed)
    return vmc, sr


vmc, sr = build(B, rank * B, True)
vmc.next_batch()
for _ in range(2):
    delta = sr.compute_update(vmc.current_batch_device, vmc.current_local_energy)
gathered = [torch.empty_like(delta) for _ in range(world)]
dist.all_gather(gathered, delta)
if rank == 0:
    assert all((g - gathered[0]).abs().max() == 0 for g in gathered), 'ranks disagree'
    print('sharded sample-space SR, %d x %d samples, P = %d: %s' % (
        world, B, delta.numel(), {k: round(v, 2) for k, v in sr.last_timings_ms.items()}), flush=True)
if args.shared_cholesky:
    sr.shared_cholesky = True
    for _ in range(2):
        delta_shared = sr.compute_update(vmc.current_batch_device, vmc.current_local_energy)
    if rank == 0:
        print('shared Cholesky: %s, relative difference to the replicated solve %.3e' % (
            {k: round(v, 2) for k, v in sr.last_timings_ms.items()}, float((delta_shared - delta).norm() / delta.norm())), flush=True)
if rank == 0:
    if not args.skip_single:
        vmc1, sr1 = build(B * world, 0, False)
        vmc1.next_batch()
        want = sr1.compute_update(vmc1.current_batch_device, vmc1.current_local_energy)
        rel = float((delta - want).norm() / want.norm())
        print('vs single process (%d samples): relative difference %.3e' % (B * world, rel), flush=True)
        assert rel < 3e-2
dist.barrier()
dist.destroy_process_group()
