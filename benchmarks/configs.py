#!/usr/bin/env python
"""Timings of the other BASELINE.json configurations (bench.py covers configs[2]); one JSON line per configuration.

  python benchmarks/configs.py            # all of: cfg1 (exact 4x4 Ising + 1-D 16), cfg2 (Heisenberg 1-D 20),
                                          #         cfg4 (J1J2 6x6, complex 1-D machine + SR), cfg5 (12x12 / 16x16 sweep)
These are parity-test workloads first (tests/), measured here for completeness; times are CUDA-event times of one
process on one GPU, best of 3 after 2 warm-up iterations."""
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


def vmc_phases(machine, inp, operator, B, engine, sr=None):
    model = Model(inp, machine.predictions)
    cond = Model(inp, machine.conditional_log_probs)
    model.engine = cond.engine = engine
    net = machine.device_net()
    sampler = FastAutoregressiveSampler(cond, B, seed=1)
    obs = Observable(operator)
    t_s, sigma = timed(lambda: sampler.next_device())
    t_e, eloc = timed(lambda: obs.local_values_device(model, sigma))
    y = (torch.conj(eloc - eloc.mean()) / B).to(torch.complex64)
    t_g, _ = timed(lambda: net.grad_weighted(net.to_sigma(sigma), y))
    res = {'batch': B, 'sample_ms': t_s, 'eloc_ms': t_e, 'grad_ms': t_g, 'samples_per_s': B / t_s * 1e3,
           'eloc_evals_per_s': B / t_e * 1e3, 'psi_evals_per_s': obs.last_num_connections / t_e * 1e3,
           'energy': [float(eloc.real.mean()), float(eloc.imag.mean())]}
    if sr is not None:
        t_sr, _ = timed(lambda: sr(model).compute_update(sigma, (torch.conj(eloc - eloc.mean()) / B).cpu().numpy()), reps=2, warm=1)
        res['sr_update_ms'] = t_sr
    return res


def main():
    out = []
    # cfg 1: exact gradient, Ising 4x4 OBC h=3 (2-D machine depth 5 / 32) and the named script's 1-D 16-site machine
    inp = Input(shape=(4, 4))
    m = ConvNetAutoregressive2D(inp, depth=5, num_of_channels=32, weights_normalization=False, seed=0)
    ev = ExactVariational(Model(inp, m.predictions), Ising(hilbert_state_shape=[4, 4], pbc=False, h=3.0), 2 ** 12)
    t, _ = timed(ev.machine_updated, reps=2, warm=1)
    out.append({'config': 'cfg1: Ising 4x4 OBC h=3, ConvNetAutoregressive2D d5 c32, ExactVariational.machine_updated (2^16 states)',
                'ms': t, 'energy': float(ev.energy_observable.current_energy.real)})
    inp = Input(shape=(16,))
    m = SimpleConvNetAutoregressive1D(inp, depth=7, num_of_channels=32, seed=0)
    ev = ExactVariational(Model(inp, m.predictions), Ising(hilbert_state_shape=[16], pbc=False, h=3.0), 2 ** 12)
    t, _ = timed(ev.machine_updated, reps=2, warm=1)
    out.append({'config': "cfg1': Ising 1-D 16 OBC h=3, SimpleConvNetAutoregressive1D d7 c32, ExactVariational.machine_updated",
                'ms': t, 'energy': float(ev.energy_observable.current_energy.real)})
    # cfg 2: Heisenberg 1-D 20 PBC, 1-D machine depth 8 / 64, batch 1024
    inp = Input(shape=(20,))
    m = SimpleConvNetAutoregressive1D(inp, depth=8, num_of_channels=64, max_dilation_rate=4, weights_normalization=False, seed=0)
    r = vmc_phases(m, inp, Heisenberg(hilbert_state_shape=[20], pbc=True), 1024, FK_ENGINE_FP32)
    r['config'] = 'cfg2: Heisenberg 1-D 20 PBC, SimpleConvNetAutoregressive1D d8 c64, batch 1024 (fp32 engine; incremental 1-D sampler)'
    out.append(r)
    # cfg 4: J1J2 6x6 OBC j2=0.5, complex 1-D machine over the raster-flattened lattice + complex SR
    inp = Input(shape=(36,))
    m = ComplexValuesSimpleConvNetAutoregressive1D(inp, depth=5, num_of_channels=16, max_dilation_rate=4, seed=0)

    class FlatJ1J2(object):   # the operator acts on the 6x6 lattice; the machine sees the 36 raster-flattened sites
        def __init__(self):
            self.op = J1J2((6, 6), j2=0.5)
            self.hilbert_state_shape = (36,)
            self.max_number_of_local_connections = self.op.max_number_of_local_connections

        def device_desc(self):
            return self.op.device_desc()
    r = vmc_phases(m, inp, FlatJ1J2(), 1024, FK_ENGINE_FP32,
                   sr=lambda model: ComplexValuesStochasticReconfiguration(model, iterative_solver=False, diag_shift=0.05))
    r['config'] = 'cfg4: J1J2 6x6 OBC j2=0.5, ComplexValuesSimpleConvNetAutoregressive1D d5 c16 (flattened lattice), batch 1024, complex SR (direct)'
    out.append(r)
    # cfg 5: Heisenberg 12x12 / 16x16 OBC sweep, 2-D machine depth 20 / 32
    for (L, batches) in [(12, [1024, 4096, 16384]), (16, [1024, 4096, 16384])]:
        for B in batches:
            inp = Input(shape=(L, L))
            m = ConvNetAutoregressive2D(inp, depth=20, num_of_channels=32, seed=0)
            r = vmc_phases(m, inp, Heisenberg(hilbert_state_shape=[L, L], pbc=False), B, FK_ENGINE_TC)
            r['config'] = 'cfg5: Heisenberg %dx%d OBC, ConvNetAutoregressive2D d20 c32, batch %d (tensor-core engines; gradient fp32)' % (L, L, B)
            out.append(r)
            del m
            torch.cuda.empty_cache()
    for r in out:
        print(json.dumps(r))


if __name__ == '__main__':
    main()
