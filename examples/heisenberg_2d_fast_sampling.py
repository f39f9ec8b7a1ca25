#!/usr/bin/env python
"""Heisenberg 2-D 10x10 OBC with ConvNetAutoregressive2D and exact autoregressive sampling -- the flowket_b200
counterpart of the reference's examples/heisenberg_2d_horvod_multy_gpu_fast_sampling.py (same objects, same shape).

  python examples/heisenberg_2d_fast_sampling.py --steps 100                      # one GPU
  torchrun --nproc-per-node 8 examples/heisenberg_2d_fast_sampling.py --steps 100  # samples sharded over 8 GPUs
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from flowket_b200 import Input, Model, FK_ENGINE_TC, FK_ENGINE_FP32  # noqa: E402
from flowket_b200.machines import ConvNetAutoregressive2D  # noqa: E402
from flowket_b200.operators import Heisenberg  # noqa: E402
from flowket_b200.optimization import VariationalMonteCarlo, DistributedVariationalMonteCarlo  # noqa: E402
from flowket_b200.optimizers import Adam, Trainer  # noqa: E402
from flowket_b200.samplers import FastAutoregressiveSampler  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--batch_size', type=int, default=1024, help='global batch (split over the ranks)')
    ap.add_argument('--depth', type=int, default=20)
    ap.add_argument('--width', type=int, default=32)
    ap.add_argument('--lr', type=float, default=1e-3)
    ap.add_argument('--engine', default='tc', choices=['tc', 'fp32'])
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    if world > 1:
        dist.init_process_group('nccl')
    batch = (args.batch_size + world - 1) // world

    inputs = Input(shape=(10, 10), dtype='int8')
    convnet = ConvNetAutoregressive2D(inputs, depth=args.depth, num_of_channels=args.width, weights_normalization=False, seed=0)
    model = Model(inputs=inputs, outputs=convnet.predictions)
    conditional_log_probs_model = Model(inputs=inputs, outputs=convnet.conditional_log_probs)
    model.engine = conditional_log_probs_model.engine = FK_ENGINE_TC if args.engine == 'tc' else FK_ENGINE_FP32
    if world > 1:
        dist.broadcast(convnet.flat_params_device(), src=0)
        convnet.params_updated()
    sampler = FastAutoregressiveSampler(conditional_log_probs_model, batch, sample_offset=rank * batch)
    operator = Heisenberg(hilbert_state_shape=(10, 10), pbc=False)
    vmc_cls = DistributedVariationalMonteCarlo if world > 1 else VariationalMonteCarlo
    variational_monte_carlo = vmc_cls(model, operator, sampler)
    trainer = Trainer(model, variational_monte_carlo, Adam(lr=args.lr, beta_1=0.9, beta_2=0.999), distributed=world > 1)
    t0 = time.time()
    for step in range(args.steps):
        energy = trainer.train_step()
        if rank == 0 and (step % 10 == 0 or step == args.steps - 1):
            print('step %4d  energy %.4f  variance %.3f  (reference ground state -251.4624)  %.1f s' % (
                step, energy.real, variational_monte_carlo.current_local_energy_variance, time.time() - t0), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
