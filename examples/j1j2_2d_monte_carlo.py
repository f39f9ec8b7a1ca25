#!/usr/bin/env python
"""J1-J2 model (j2 = 0.5, OBC) with Monte-Carlo gradients -- the flowket_b200 counterpart of the reference's
examples/j1j2_2d_monte_carlo_4.py (4x4, Adam, exact diagonalisation -30.022227800323677) and of BASELINE.json configs[3]
(6x6, complex-valued 1-D machine + stochastic reconfiguration; that composition is not in the reference).

  python examples/j1j2_2d_monte_carlo.py                       # 4x4, ConvNetAutoregressive2D, Adam (the reference script)
  python examples/j1j2_2d_monte_carlo.py --lattice 6 --sr      # 6x6, complex 1-D machine over the flattened lattice, SR

Same calls as the reference: compile / fit_generator / default callbacks / BadEigenStateStopping / TerminateOnNaN /
set_sampler for the larger second-stage batch / evaluate with the D4-symmetrised wave function.
NOTE: written at the end of round 1 after the GPU budget was spent -- every component it uses is covered by the GPU
tests, the script as a whole has only been exercised up to the first device call."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowket_b200 import Input, Model  # noqa: E402
from flowket_b200.callbacks import TerminateOnNaN  # noqa: E402
from flowket_b200.callbacks.monte_carlo import TensorBoardWithGeneratorValidationData, \
    default_wave_function_stats_callbacks_factory, BadEigenStateStopping  # noqa: E402
from flowket_b200.evaluation import evaluate  # noqa: E402
from flowket_b200.operators import J1J2, FlattenedOperator  # noqa: E402
from flowket_b200.machines import ConvNetAutoregressive2D, ComplexValuesSimpleConvNetAutoregressive1D  # noqa: E402
from flowket_b200.machines import make_2d_obc_invariants  # noqa: E402
from flowket_b200.optimization import VariationalMonteCarlo, loss_for_energy_minimization  # noqa: E402
from flowket_b200.optimizers import Adam, ComplexValuesStochasticReconfiguration  # noqa: E402
from flowket_b200.samplers import FastAutoregressiveSampler  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--lattice', type=int, default=4)
ap.add_argument('--sr', action='store_true', help='complex 1-D machine + stochastic reconfiguration (configs[3])')
ap.add_argument('--depth', type=int, default=5)
ap.add_argument('--width', type=int, default=32)
ap.add_argument('--lr', type=float, default=1e-3)
ap.add_argument('--batch_size', type=int, default=2 ** 10)
ap.add_argument('--steps_per_epoch', type=int, default=2 ** 8)
ap.add_argument('--epochs', type=int, nargs=2, default=[60, 80], help='last epoch of the first / second stage')
ap.add_argument('--log_dir', default=None)
args = ap.parse_args()

L = args.lattice
true_ground_state_energy = -30.022227800323677 if L == 4 else None
if args.sr:
    hilbert_state_shape = (L * L,)          # the complex machine is 1-D: raster-flattened lattice
    inputs = Input(shape=hilbert_state_shape, dtype='int8')
    convnet = ComplexValuesSimpleConvNetAutoregressive1D(inputs, depth=args.depth, num_of_channels=args.width,
                                                         max_dilation_rate=4)
else:
    hilbert_state_shape = (L, L)
    inputs = Input(shape=hilbert_state_shape, dtype='int8')
    convnet = ConvNetAutoregressive2D(inputs, depth=args.depth, num_of_channels=args.width, weights_normalization=False)
model = Model(inputs=inputs, outputs=convnet.predictions)
conditional_log_probs_model = Model(inputs=inputs, outputs=convnet.conditional_log_probs)

if args.sr:
    optimizer = ComplexValuesStochasticReconfiguration(model, lr=args.lr * 10, diag_shift=0.05, iterative_solver=False)
else:
    optimizer = Adam(lr=args.lr, beta_1=0.9, beta_2=0.999)
model.compile(optimizer=optimizer, loss=loss_for_energy_minimization)
model.summary()
operator = J1J2(hilbert_state_shape=[L, L], j2=0.5, pbc=False)
if args.sr:
    operator = FlattenedOperator(operator)     # the 1-D machine sees the raster-flattened lattice
sampler = FastAutoregressiveSampler(conditional_log_probs_model, args.batch_size)
monte_carlo_generator = VariationalMonteCarlo(model, operator, sampler)

callbacks = default_wave_function_stats_callbacks_factory(monte_carlo_generator, log_in_batch_or_epoch=False,
                                                          true_ground_state_energy=true_ground_state_energy)
callbacks.append(TerminateOnNaN())
early_stopping = None
if true_ground_state_energy is not None:
    early_stopping = BadEigenStateStopping(true_ground_state_energy)
    callbacks.append(early_stopping)
if args.log_dir:
    callbacks.append(TensorBoardWithGeneratorValidationData(log_dir=args.log_dir, generator=monte_carlo_generator,
                                                            update_freq='epoch'))
model.fit_generator(monte_carlo_generator.to_generator(), steps_per_epoch=args.steps_per_epoch, epochs=args.epochs[0],
                    callbacks=callbacks, max_queue_size=0, workers=0, verbose=1)
model.save_weights('before_increasing_batch_j1j2_%d.h5' % L)
if early_stopping is not None and early_stopping.stopped_epoch is not None:
    print('stopped at epoch %s because of a bad eigenstate' % early_stopping.stopped_epoch)
    sys.exit()

print('increasing the batch size to %d' % (args.batch_size * 8))
sampler = FastAutoregressiveSampler(conditional_log_probs_model, args.batch_size * 8)
monte_carlo_generator.set_sampler(sampler)
model.fit_generator(monte_carlo_generator.to_generator(), steps_per_epoch=args.steps_per_epoch, epochs=args.epochs[1],
                    initial_epoch=args.epochs[0], callbacks=callbacks, max_queue_size=0, workers=0, verbose=1)
model.save_weights('final_j1j2_%d.h5' % L)

if not args.sr:     # symmetrised evaluation needs the 2-D machine (D4 acts on the lattice)
    evaluation_inputs = Input(shape=hilbert_state_shape, dtype='int8')
    invariant_model = make_2d_obc_invariants(evaluation_inputs, model)
    generator = VariationalMonteCarlo(invariant_model, operator, sampler)
    keys = {'energy/energy': 'energy'}
    if true_ground_state_energy is not None:
        keys['energy/relative_error'] = 'relative_error'
    print(evaluate(generator, steps=20, callbacks=callbacks[:4], keys_to_progress_bar_mapping=keys))
