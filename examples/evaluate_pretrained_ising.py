#!/usr/bin/env python
"""Evaluate the reference's pretrained 12x12 transverse-field Ising machines with the symmetrised wave function, the way
experiments/run_evaluation.py:12-26 + experiments/ising_runner.py do: samples from the base network, E_loc with
psi_sym = up/down-invariant(D4-invariant(psi)) (16 forwards per configuration), mini-batches of 256, means over the
mini-batches.  Published (experiments/README.md:38-47, 2^15 samples):

  Gamma   energy          variance   |Mz|
  2      -346.9817926     0.00125    0.783929189
  2.5    -395.6618438     0.00421    0.57235633
  3      -457.0420317     0.000821   0.1622390747
  3.5    -524.5172088     0.000647   0.1106881036
  4      -593.5389339     0.000777   0.09679073758

  python examples/evaluate_pretrained_ising.py [--gamma 3] [--num_of_samples 32768] [--engine tc|fp32]
The weights are the committed fixtures tests/golden/ising_12x12_gamma*_keras_weights.npz (exported from the reference's
experiments/weights/*.h5 by oracle/make_golden.py; `--weights_path some.h5` reads a Keras file directly)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flowket_b200 import Input, Model, FK_ENGINE_TC, FK_ENGINE_FP32  # noqa: E402
from flowket_b200.machines import ConvNetAutoregressive2D, make_2d_obc_invariants, make_up_down_invariant  # noqa: E402
from flowket_b200.operators import Ising  # noqa: E402
from flowket_b200.samplers import FastAutoregressiveSampler  # noqa: E402
from flowket_b200.observables.monte_carlo import Observable  # noqa: E402

PUBLISHED = {2.0: (-346.9817926, 0.00125, 0.783929189), 2.5: (-395.6618438, 0.00421, 0.57235633),
             3.0: (-457.0420317, 0.000821, 0.1622390747), 3.5: (-524.5172088, 0.000647, 0.1106881036),
             4.0: (-593.5389339, 0.000777, 0.09679073758)}


def fixture(gamma):
    tag = ('%g' % gamma).replace('.', '_')
    return os.path.join(ROOT, 'tests', 'golden', 'ising_12x12_gamma%s_keras_weights.npz' % tag)


def evaluate(gamma, num_of_samples, engine, mini_batch_size=256, weights_path=None, symmetrise=True, seed=0, chunk=2048):
    inp = Input(shape=(12, 12), dtype='int8')
    machine = ConvNetAutoregressive2D(inp, depth=10, num_of_channels=32)
    model = Model(inp, machine.predictions)
    cond = Model(inp, machine.conditional_log_probs)
    if weights_path:
        model.load_weights(weights_path)
    else:
        with np.load(fixture(gamma)) as f:
            model.set_weights([f['w%04d' % i] for i in range(len(f.files))])
    model.engine = cond.engine = engine
    psi = model
    if symmetrise:
        psi = make_up_down_invariant(inp, make_2d_obc_invariants(inp, model))
    obs = Observable(Ising(hilbert_state_shape=[12, 12], pbc=False, h=gamma))
    sampler = FastAutoregressiveSampler(cond, chunk, seed=seed)
    energies, variances, mz = [], [], []
    for _ in range(num_of_samples // chunk):
        sigma = sampler.next_device()
        if symmetrise:
            eloc = obs.local_values_device_generic(psi.predict_device, sigma)
        else:
            eloc = obs.local_values_device(model, sigma)
        e = eloc.real.reshape(-1, mini_batch_size)           # the reference averages per mini-batch statistics
        energies.append(e.mean(dim=1).cpu().numpy())
        variances.append(e.var(dim=1, unbiased=False).cpu().numpy())
        mz.append(sigma.reshape(sigma.shape[0], -1).float().mean(dim=1).abs().reshape(-1, mini_batch_size).mean(dim=1).cpu().numpy())
    energies, variances, mz = np.concatenate(energies), np.concatenate(variances), np.concatenate(mz)
    return {'gamma': gamma, 'samples': int(num_of_samples), 'energy': float(energies.mean()),
            'energy_stderr': float(np.sqrt(variances.mean() / num_of_samples)), 'variance': float(variances.mean()),
            'abs_mz': float(mz.mean()), 'published': dict(zip(('energy', 'variance', 'abs_mz'), PUBLISHED.get(gamma, (None,) * 3))),
            'symmetrised': bool(symmetrise), 'engine': 'tc' if engine == FK_ENGINE_TC else 'fp32'}


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--gamma', type=float, nargs='*', default=[2.0, 2.5, 3.0, 3.5, 4.0])
    ap.add_argument('--num_of_samples', type=int, default=2 ** 15)
    ap.add_argument('--engine', default='tc', choices=['tc', 'fp32'])
    ap.add_argument('--weights_path', default=None)
    ap.add_argument('--no_symmetrisation', action='store_true')
    args = ap.parse_args()
    for g in args.gamma:
        r = evaluate(g, args.num_of_samples, FK_ENGINE_TC if args.engine == 'tc' else FK_ENGINE_FP32,
                     weights_path=args.weights_path, symmetrise=not args.no_symmetrisation)
        print(json.dumps(r), flush=True)
