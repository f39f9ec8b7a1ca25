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

from flowket_b200 import Input, Model, FK_ENGINE_TC, FK_ENGINE_TC_EXACT, FK_ENGINE_FP32  # noqa: E402
from flowket_b200.machines import ConvNetAutoregressive2D  # noqa: E402
from flowket_b200.operators import Heisenberg  # noqa: E402
from flowket_b200.optimization import VariationalMonteCarlo, DistributedVariationalMonteCarlo  # noqa: E402
from flowket_b200.optimizers import Adam, StochasticReconfiguration, Trainer  # noqa: E402
from flowket_b200.samplers import FastAutoregressiveSampler  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, nargs='+', default=[100], help='updates per stage')
    ap.add_argument('--batch_size', type=int, nargs='+', default=[1024], help='global batch per stage (split over the ranks); '
                    'staged schedules like experiments/heisenberg_runner.py: --steps 2500 400 --batch_size 1024 4096')
    ap.add_argument('--beta_2', type=float, default=0.999, help='0.9 in the paper runs (experiments/train.py:52)')
    ap.add_argument('--checkpoint', default=None, help='path of a resumable checkpoint (written every 60 s and at the end)')
    ap.add_argument('--eval_samples', type=int, default=0, help='final energy estimate from this many fresh samples')
    ap.add_argument('--print_every', type=int, default=10)
    ap.add_argument('--depth', type=int, default=20)
    ap.add_argument('--width', type=int, default=32)
    ap.add_argument('--lr', type=float, default=1e-3)
    ap.add_argument('--engine', default='tc', choices=['tc', 'tc_exact', 'fp32'],
                    help='local-energy engine (the sampler and the gradient use the fp16 tensor-core engine unless fp32)')
    ap.add_argument('--optimizer', default='adam', choices=['adam', 'sr'],
                    help="sr: stochastic reconfiguration in sample space, device pipeline (north_star's target step)")
    ap.add_argument('--diag_shift', type=float, default=0.05)
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    if world > 1:
        dist.init_process_group('nccl')
    assert len(args.steps) == len(args.batch_size)
    batch = (args.batch_size[0] + world - 1) // world

    inputs = Input(shape=(10, 10), dtype='int8')
    convnet = ConvNetAutoregressive2D(inputs, depth=args.depth, num_of_channels=args.width, weights_normalization=False, seed=0)
    model = Model(inputs=inputs, outputs=convnet.predictions)
    conditional_log_probs_model = Model(inputs=inputs, outputs=convnet.conditional_log_probs)
    conditional_log_probs_model.engine = FK_ENGINE_FP32 if args.engine == 'fp32' else FK_ENGINE_TC
    model.engine = {'tc': FK_ENGINE_TC, 'tc_exact': FK_ENGINE_TC_EXACT, 'fp32': FK_ENGINE_FP32}[args.engine]
    if world > 1:
        dist.broadcast(convnet.flat_params_device(), src=0)
        convnet.params_updated()
    sampler = FastAutoregressiveSampler(conditional_log_probs_model, batch, sample_offset=rank * batch)
    operator = Heisenberg(hilbert_state_shape=(10, 10), pbc=False)
    vmc_cls = DistributedVariationalMonteCarlo if world > 1 else VariationalMonteCarlo
    variational_monte_carlo = vmc_cls(model, operator, sampler)
    if args.optimizer == 'sr':
        optimizer = StochasticReconfiguration(model, lr=args.lr, diag_shift=args.diag_shift, sample_space=True, distributed=world > 1)
    else:
        optimizer = Adam(lr=args.lr, beta_1=0.9, beta_2=args.beta_2)
    trainer = Trainer(model, variational_monte_carlo, optimizer, distributed=world > 1)
    if args.checkpoint and os.path.exists(args.checkpoint + '.npz'):
        trainer.load_checkpoint(args.checkpoint)
        if rank == 0:
            print('resumed from %s at update %d' % (args.checkpoint, len(trainer.history)), flush=True)
    resumed_updates = len(trainer.history)
    t0 = time.time()
    last_ckpt = time.time()
    done = 0
    for stage, (steps, global_batch) in enumerate(zip(args.steps, args.batch_size)):
        stage_batch = (global_batch + world - 1) // world
        if stage_batch != variational_monte_carlo.sampler.batch_size:   # staged schedule (experiments/train.py:99-101)
            new_sampler = variational_monte_carlo.sampler.copy_with_new_batch_size(stage_batch)
            new_sampler.sample_offset = rank * stage_batch
            variational_monte_carlo.set_sampler(new_sampler)
        for step in range(steps):
            done += 1
            if done <= resumed_updates:      # already applied before the checkpoint was written
                continue
            energy = trainer.train_step()
            if rank == 0 and (step % args.print_every == 0 or step == steps - 1):
                print('stage %d step %5d  batch %5d  energy %.4f  variance %.3f  (reference ground state -251.4624)  %.1f s' % (
                    stage, step, global_batch, energy.real, variational_monte_carlo.current_local_energy_variance, time.time() - t0),
                    flush=True)
            if args.checkpoint and rank == 0 and time.time() - last_ckpt > 60:
                trainer.save_checkpoint(args.checkpoint)
                last_ckpt = time.time()
    if args.checkpoint and rank == 0:
        trainer.save_checkpoint(args.checkpoint)
    if args.eval_samples and rank == 0:
        import numpy as np
        from flowket_b200.observables.monte_carlo import Observable
        obs = Observable(operator)
        ev = FastAutoregressiveSampler(conditional_log_probs_model, 8192, seed=99)
        vals = []
        for _ in range(max(1, args.eval_samples // 8192)):
            vals.append(obs.local_values_device(model, ev.next_device()).real.cpu().numpy())
        vals = np.concatenate(vals)
        print('evaluation: %d samples  energy %.4f +- %.4f  variance %.3f  (reference ground state -251.4624, published -251.4536)' % (
            len(vals), vals.mean(), vals.std() / np.sqrt(len(vals)), vals.var()), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
