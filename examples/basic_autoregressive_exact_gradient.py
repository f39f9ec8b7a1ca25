#!/usr/bin/env python
"""Transverse-field Ising chain (16 sites, h = 3, OBC) trained with the *exact* energy gradient over all 2^16 states --
the script BASELINE.json configs[0] names (the reference's examples/basic_autoregressive_exact_gradient.py), line for
line with flowket_b200 objects: Input/Model, SimpleConvNetAutoregressive1D, Ising, ExactVariational, Adam wrapped by
convert_to_accumulate_gradient_optimizer, the exact callbacks, fit_generator.  Exact diagonalisation: -49.257706531889006.

  python examples/basic_autoregressive_exact_gradient.py --cycles 300 [--lattice 4 4]   # 4x4: ED -50.18662388277671
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowket_b200 import Input, Model  # noqa: E402
from flowket_b200.callbacks import TensorBoard  # noqa: E402
from flowket_b200.callbacks.exact import default_wave_function_callbacks_factory  # noqa: E402
from flowket_b200.machines import SimpleConvNetAutoregressive1D, ConvNetAutoregressive2D  # noqa: E402
from flowket_b200.operators import Ising  # noqa: E402
from flowket_b200.optimization import ExactVariational, loss_for_energy_minimization  # noqa: E402
from flowket_b200.optimizers import Adam, convert_to_accumulate_gradient_optimizer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--lattice', type=int, nargs='+', default=[16])
ap.add_argument('--cycles', type=int, default=300, help='parameter updates (full enumerations) per epoch')
ap.add_argument('--epochs', type=int, default=2)
ap.add_argument('--log_dir', default=None)
args = ap.parse_args()

hilbert_state_shape = list(args.lattice)
inputs = Input(shape=hilbert_state_shape, dtype='int8')
if len(hilbert_state_shape) == 1:
    convnet = SimpleConvNetAutoregressive1D(inputs, depth=7, num_of_channels=32, weights_normalization=False)
    true_ground_state_energy = -49.257706531889006 if hilbert_state_shape == [16] else None
else:
    convnet = ConvNetAutoregressive2D(inputs, depth=5, num_of_channels=32, weights_normalization=False)
    true_ground_state_energy = -50.18662388277671 if hilbert_state_shape == [4, 4] else None
model = Model(inputs=inputs, outputs=convnet.predictions)

batch_size = 2 ** 12
operator = Ising(h=3.0, hilbert_state_shape=hilbert_state_shape, pbc=False)
exact_variational = ExactVariational(model, operator, batch_size)
steps_per_epoch = args.cycles * exact_variational.num_of_batch_until_full_cycle

optimizer = Adam(lr=0.001, beta_1=0.9, beta_2=0.999)
convert_to_accumulate_gradient_optimizer(
    optimizer,
    update_params_frequency=exact_variational.num_of_batch_until_full_cycle,
    accumulate_sum_or_mean=True)
model.compile(optimizer=optimizer, loss=loss_for_energy_minimization)
model.summary()

callbacks = default_wave_function_callbacks_factory(exact_variational, true_ground_state_energy=true_ground_state_energy)
if args.log_dir:
    callbacks.append(TensorBoard(log_dir=args.log_dir, update_freq=exact_variational.num_of_batch_until_full_cycle))
model.fit_generator(exact_variational.to_generator(), steps_per_epoch=steps_per_epoch, epochs=args.epochs,
                    callbacks=callbacks, max_queue_size=0, workers=0, verbose=1)
model.save_weights('final_ising_exact_gradient')
