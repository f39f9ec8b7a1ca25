"""Probe (torchrun, G >= 2 ranks, NCCL): the split solve of the sharded sample-space SR step -- one rank factors the SR
matrix while the others evaluate local energies, samples gathered and dealt over the ranks
(flowket_b200/optimizers/sample_space_sr.py) -- returns the update of the plain sharded step (every rank evaluates its own
samples and factors), identically on every rank, for every way of dealing the samples; also through
StochasticReconfiguration.step_generator and through the forced fp64 re-solve.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 \\
      tests/tools_split_solve.py [--depth 20 --batch 512 --lattice 10 --engine tc_exact]"""
import argparse
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from flowket_b200 import Input, Model, FK_ENGINE_TC, FK_ENGINE_TC_EXACT
from flowket_b200.machines import ConvNetAutoregressive2D
from flowket_b200.operators import Heisenberg
from flowket_b200.samplers import FastAutoregressiveSampler
from flowket_b200.optimization import DistributedVariationalMonteCarlo
from flowket_b200.optimizers import StochasticReconfiguration

ap = argparse.ArgumentParser()
ap.add_argument('--depth', type=int, default=4)
ap.add_argument('--batch', type=int, default=128, help='samples per rank')
ap.add_argument('--lattice', type=int, default=6)
ap.add_argument('--engine', default='tc_exact', choices=['tc', 'tc_exact'])
args = ap.parse_args()
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl')
L, B = args.lattice, args.batch

inp = Input(shape=(L, L), dtype='int8')
m = ConvNetAutoregressive2D(inp, depth=args.depth, num_of_channels=32, seed=0)
model, cond = Model(inp, m.predictions), Model(inp, m.conditional_log_probs)
model.engine = FK_ENGINE_TC_EXACT if args.engine == 'tc_exact' else FK_ENGINE_TC
cond.engine = FK_ENGINE_TC
op = Heisenberg(hilbert_state_shape=[L, L], pbc=False)
sampler = FastAutoregressiveSampler(cond, B, seed=11, sample_offset=rank * B)
vmc = DistributedVariationalMonteCarlo(model, op, sampler)
sr = StochasticReconfiguration(model, diag_shift=0.05, sample_space=True, distributed=True)


def same_on_every_rank(t, what):
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t.contiguous())
    assert all(bool((p == parts[0]).all()) for p in parts), '%s: the ranks disagree' % what


vmc.next_batch()
sigma, eloc = vmc.current_batch_device, vmc.current_local_energy_device
want = sr.compute_update(sigma, eloc).clone()                 # plain: own samples, replicated factorisation
plain_ms = dict(sr.last_timings_ms)
same_on_every_rank(want, 'plain update')
pipe = sr._pipeline
fn = vmc.local_energy_function()
# the sharded update against ONE process holding the whole global batch (rank-major order = the global sample order)
sig_all = torch.empty((world * B,) + tuple(sigma.shape[1:]), dtype=sigma.dtype, device=sigma.device)
dist.all_gather_into_tensor(sig_all, sigma.contiguous())
eloc_all = torch.empty(world * B, dtype=eloc.dtype, device=eloc.device)
dist.all_gather_into_tensor(eloc_all, eloc.contiguous())
rel_single = 0.0
if rank == 0:
    sr1 = StochasticReconfiguration(model, diag_shift=0.05, sample_space=True, distributed=False)
    single = sr1.compute_update(sig_all, eloc_all)
    rel_single = float((want - single).norm() / single.norm())
    assert rel_single < 2e-3, rel_single      # (bf16 rows, fp32 Gram accumulated in a different order: partial Grams per rank)
dist.barrier()
report = []
for rho in (None, 0.0, 0.3, 10.0):      # measured split, equal shares, a small solver share, an empty solver share
    for _ in range(2):
        if rho is not None:
            pipe.split_rho, pipe.split_kappa, pipe._split_events = rho, 1.0, None
        got = sr.compute_update(sigma, fn)
    assert pipe._split_events is not None, 'the split solve did not run'
    same_on_every_rank(got, 'split update')
    mine = sr.last_local_energy
    assert float((mine - eloc).abs().max()) <= 2e-5 * float(eloc.abs().max()), 'local energies of the dealt samples differ'
    rel = float((got - want).norm() / want.norm())
    report.append((rho, list(pipe.split_counts), rel, {k: round(v, 2) for k, v in sr.last_timings_ms.items()}))
    assert rel < 1e-3, (rho, rel)
# forced fp64 re-solve (the solver rank solves, everybody applies)
got64 = pipe.refine_with_fp64()
same_on_every_rank(got64, 'fp64 re-solve')
rel64 = float((got64 - want).norm() / want.norm())
assert rel64 < 1e-3, rel64
# through the generator: samples from the generator, local energies inside the update, statistics handed back
params0 = m.flat_params_device().clone()
sr.step_generator(vmc)
e_split, var_split = complex(vmc.current_energy), float(vmc.current_local_energy_variance)
assert vmc.current_local_energy.shape == (B,)
assert abs(e_split.real) > 0 and var_split >= 0
moved = float((m.flat_params_device() - params0).abs().max())
assert moved > 0
if rank == 0:
    print('plain sharded step: %s' % {k: round(v, 2) for k, v in plain_ms.items()}, flush=True)
    print('plain sharded step vs ONE process on the global batch of %d samples: relative difference of the update %.2e' % (world * B, rel_single), flush=True)
    for rho, counts, rel, tm in report:
        print('split solve, rho %s: samples per rank %s, relative difference of the update %.2e, %s' % (rho, counts, rel, tm), flush=True)
    print('fp64 re-solve through the solver rank: relative difference %.2e' % rel64, flush=True)
    print('step_generator: energy %.4f%+.4fj, variance %.3f' % (e_split.real, e_split.imag, var_split), flush=True)
    print('SPLIT SOLVE OK', flush=True)
dist.barrier()
dist.destroy_process_group()
