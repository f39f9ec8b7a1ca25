#!/usr/bin/env python
"""Timings of the other BASELINE.json configurations (bench.py covers configs[2]); one JSON line per configuration.

  python benchmarks/configs.py            # all of: cfg1 (exact 4x4 Ising + 1-D 16), cfg2 (Heisenberg 1-D 20),
                                          #         cfg4 (J1J2 6x6, complex 1-D machine + SR), cfg5 (12x12 / 16x16 sweep)
  python benchmarks/configs.py --only cfg5 --max-batch 65536
  python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 benchmarks/configs.py --only cfg5
                                          # the sweep of cfg5 with every total batch sharded over G GPUs (one rank per GPU, NCCL)
These are parity-test workloads first (tests/), measured here for completeness; times are CUDA-event times, best of 3 after
2 warm-up iterations (1 + 1 for batches >= 16384), max over the ranks under torchrun (the gradient phase then includes its
allreduce)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowket_b200 import Input, Model, FK_ENGINE_TC, FK_ENGINE_FP32  # noqa: E402
from flowket_b200.machines import ConvNetAutoregressive2D, SimpleConvNetAutoregressive1D, \
    ComplexValuesSimpleConvNetAutoregressive1D  # noqa: E402
from flowket_b200.operators import Heisenberg, Ising, J1J2  # noqa: E402
from flowket_b200.samplers import FastAutoregressiveSampler  # noqa: E402
from flowket_b200.observables.monte_carlo import Observable  # noqa: E402
from flowket_b200.optimization import ExactVariational  # noqa: E402
from flowket_b200.optimizers import ComplexValuesStochasticReconfiguration  # noqa: E402


def timed(fn, reps=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


def _world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def _max_over_ranks(ms):
    import torch.distributed as dist
    world, _ = _world()
    if world == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def vmc_phases(machine, inp, operator, B, engine, sr=None, eloc_engines=None):
    """sample / E_loc / weighted gradient of a total batch B (sharded over the ranks under torchrun)"""
    import torch.distributed as dist
    world, rank = _world()
    assert B % world == 0
    Bl = B // world
    model = Model(inp, machine.predictions)
    cond = Model(inp, machine.conditional_log_probs)
    model.engine = cond.engine = engine
    net = machine.device_net()
    sampler = FastAutoregressiveSampler(cond, Bl, seed=1, sample_offset=rank * Bl)
    obs = Observable(operator)
    reps, warm = (1, 1) if B >= 16384 else (3, 2)
    t_s, sigma = timed(lambda: sampler.next_device(), reps, warm)
    t_e, eloc = timed(lambda: obs.local_values_device(model, sigma), reps, warm)
    n_conn = float(obs.last_num_connections)
    esum = torch.stack([eloc.real.sum(), eloc.imag.sum(), torch.tensor(n_conn, dtype=torch.float64, device=eloc.device)])
    if world > 1:
        dist.all_reduce(esum)
    emean = torch.complex(esum[0], esum[1]) / B
    y = (torch.conj(eloc - emean) / B).to(torch.complex64)

    def grad():
        tc_grad = engine == FK_ENGINE_TC and net.lib.fk_grad_weighted_tc_workspace_bytes(net.handle, 1) >= 0
        g = net.grad_weighted(net.to_sigma(sigma), y, engine=FK_ENGINE_TC if tc_grad else FK_ENGINE_FP32)
        if world > 1:
            dist.all_reduce(g)
        return g
    t_g, _ = timed(grad, reps, warm)
    t_s, t_e, t_g = _max_over_ranks(t_s), _max_over_ranks(t_e), _max_over_ranks(t_g)
    res = {'batch': B, 'n_gpus': world, 'sample_ms': t_s, 'eloc_ms': t_e, 'grad_ms': t_g, 'samples_per_s': B / t_s * 1e3,
           'eloc_evals_per_s': B / t_e * 1e3, 'psi_evals_per_s': float(esum[2]) / t_e * 1e3,
           'step_samples_per_s': B / (t_s + t_e + t_g) * 1e3, 'energy': [float(emean.real), float(emean.imag)]}
    for name, eng in (eloc_engines or {}).items():
        model.engine = eng
        t_x, _ = timed(lambda: obs.local_values_device(model, sigma), reps, warm)
        res['eloc_ms_' + name] = _max_over_ranks(t_x)
    model.engine = engine
    if sr is not None:
        t_sr, _ = timed(lambda: sr(model).compute_update(sigma, (torch.conj(eloc - eloc.mean()) / B).cpu().numpy()), reps=2, warm=1)
        res['sr_update_ms'] = t_sr
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default='all', choices=['all', 'cfg1', 'cfg2', 'cfg4', 'cfg5'])
    ap.add_argument('--max-batch', type=int, default=16384, help='largest total batch of the cfg5 sweep (SURVEY: up to 65536)')
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0'))))
        assert args.only == 'cfg5', 'under torchrun only the sharded sweep (--only cfg5) is defined'
    rank = int(os.environ.get('RANK', '0'))
    want = lambda name: args.only in ('all', name)
    out = []
    if want('cfg1'):
        # cfg 1: exact gradient, Ising 4x4 OBC h=3 (2-D machine depth 5 / 32) and the named script's 1-D 16-site machine
        inp = Input(shape=(4, 4))
        m = ConvNetAutoregressive2D(inp, depth=5, num_of_channels=32, weights_normalization=False, seed=0)
        ev = ExactVariational(Model(inp, m.predictions), Ising(hilbert_state_shape=[4, 4], pbc=False, h=3.0), 2 ** 12)
        t, _ = timed(ev.machine_updated, reps=2, warm=1)
        out.append({'config': 'cfg1: Ising 4x4 OBC h=3, ConvNetAutoregressive2D d5 c32, ExactVariational.machine_updated (2^16 states, '
                              'tables resident on the device)', 'ms': t, 'energy': float(ev.energy_observable.current_energy.real),
                    'wave_function_ms': 1e3 * (ev.wave_function_update_end_time - ev.machine_updated_start_time),
                    'local_energy_ms': 1e3 * (ev.local_energy_update_end_time - ev.wave_function_update_end_time)})
        inp = Input(shape=(16,))
        m = SimpleConvNetAutoregressive1D(inp, depth=7, num_of_channels=32, seed=0)
        ev = ExactVariational(Model(inp, m.predictions), Ising(hilbert_state_shape=[16], pbc=False, h=3.0), 2 ** 12)
        t, _ = timed(ev.machine_updated, reps=2, warm=1)
        out.append({'config': "cfg1': Ising 1-D 16 OBC h=3, SimpleConvNetAutoregressive1D d7 c32, ExactVariational.machine_updated",
                    'ms': t, 'energy': float(ev.energy_observable.current_energy.real)})
    if want('cfg2'):
        # cfg 2: Heisenberg 1-D 20 PBC, 1-D machine depth 8 / 64, batch 1024
        inp = Input(shape=(20,))
        m = SimpleConvNetAutoregressive1D(inp, depth=8, num_of_channels=64, max_dilation_rate=4, weights_normalization=False, seed=0)
        r = vmc_phases(m, inp, Heisenberg(hilbert_state_shape=[20], pbc=True), 1024, FK_ENGINE_FP32)
        r['config'] = 'cfg2: Heisenberg 1-D 20 PBC, SimpleConvNetAutoregressive1D d8 c64, batch 1024 (fp32 engine; incremental 1-D sampler)'
        out.append(r)
    if want('cfg4'):
        # cfg 4: J1J2 6x6 OBC j2=0.5, complex 1-D machine over the raster-flattened lattice + complex SR
        class FlatJ1J2(object):   # the operator acts on the 6x6 lattice; the machine sees the 36 raster-flattened sites
            def __init__(self):
                self.op = J1J2((6, 6), j2=0.5)
                self.hilbert_state_shape = (36,)
                self.max_number_of_local_connections = self.op.max_number_of_local_connections

            def device_desc(self):
                return self.op.device_desc()
        for depth, ch, iterative in ((5, 16, False), (8, 32, True)):
            inp = Input(shape=(36,))
            m = ComplexValuesSimpleConvNetAutoregressive1D(inp, depth=depth, num_of_channels=ch, max_dilation_rate=4, seed=0)
            r = vmc_phases(m, inp, FlatJ1J2(), 1024, FK_ENGINE_FP32,
                           sr=lambda model: ComplexValuesStochasticReconfiguration(model, iterative_solver=iterative, diag_shift=0.05))
            r['config'] = ('cfg4: J1J2 6x6 OBC j2=0.5, ComplexValuesSimpleConvNetAutoregressive1D d%d c%d (flattened lattice), batch 1024, '
                           'complex SR (%s)' % (depth, ch, 'conjugate gradient, reference defaults' if iterative else 'direct'))
            out.append(r)
    if want('cfg5'):
        # cfg 5: Heisenberg 12x12 / 16x16 OBC sweep, 2-D machine depth 20 / 32; + the 10x10 lattice of cfg 3 for comparison
        batches = [b for b in (1024, 4096, 16384, 65536) if b <= args.max_batch and b % world == 0]
        for L in (12, 16):
            for B in batches:
                inp = Input(shape=(L, L))
                m = ConvNetAutoregressive2D(inp, depth=20, num_of_channels=32, seed=0)
                r = vmc_phases(m, inp, Heisenberg(hilbert_state_shape=[L, L], pbc=False), B, FK_ENGINE_TC)
                r['config'] = ('cfg5: Heisenberg %dx%d OBC, ConvNetAutoregressive2D d20 c32, total batch %d on %d GPU(s) (fp16 tensor-core '
                               'sampler and local energy; gradient: fp32 engine, the tensor-core backward covers one-tile lattices)'
                               % (L, L, B, world))
                out.append(r)
                del m
                torch.cuda.empty_cache()
    if rank == 0:
        for r in out:
            print(json.dumps(r), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
