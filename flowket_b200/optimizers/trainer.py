"""Thin host loop standing in for Keras `fit_generator` + `train_on_batch` on this path
(SURVEY.md section 3.1): pull a mini-batch (sigma, y) from the generator, differentiate
loss_for_energy_minimization on the device (fk_grad_weighted), allreduce, apply the update.

Keras' mean over the mini-batch (optimization/loss.py:4-5 + Keras reduction) is kept: grad = (1/mb) * sum_b ...
Gradients of the `update_params_frequency` mini-batches of one batch are summed before the update, like
convert_to_accumulate_gradient_optimizer (optimizers/accumulate_gradient_optimizer.py:54-82)."""
import numpy as np


def allreduce_sum_(tensor):
    """In-place sum over ranks (NCCL on CUDA tensors, gloo on CPU tensors); no-op without torch.distributed."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


class SGD(object):
    def __init__(self, lr=0.01):
        self.lr = lr

    def step(self, params, grad):
        params.add_(grad, alpha=-self.lr)


class Adam(object):
    """Adam with Keras defaults (epsilon 1e-7); the paper runs use beta_1 = beta_2 = 0.9 (experiments/train.py:52)."""

    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.lr, self.beta_1, self.beta_2, self.epsilon = lr, beta_1, beta_2, epsilon
        self.t, self.m, self.v = 0, None, None

    def step(self, params, grad):
        import torch
        if self.m is None:
            self.m, self.v = torch.zeros_like(params), torch.zeros_like(params)
        self.t += 1
        self.m.mul_(self.beta_1).add_(grad, alpha=1 - self.beta_1)
        self.v.mul_(self.beta_2).addcmul_(grad, grad, value=1 - self.beta_2)
        lr_t = self.lr * np.sqrt(1 - self.beta_2 ** self.t) / (1 - self.beta_1 ** self.t)
        params.addcdiv_(self.m, self.v.sqrt().add_(self.epsilon), value=-lr_t)


class Trainer(object):
    def __init__(self, model, generator, optimizer, distributed=False):
        self.model, self.generator, self.optimizer, self.distributed = model, generator, optimizer, distributed
        self.machine = model.machine
        self.history = []
        self.logs = []          # one dict per epoch, filled by the callbacks of fit()

    def gradient(self, x, y):
        """(1/mb) d/dtheta sum_b 2 Re(log psi_b y_b) on the device."""
        import torch
        net = self.machine.device_net()
        sigma = net.to_sigma(x)
        y_t = torch.as_tensor(np.asarray(y, np.complex64)) if not hasattr(y, 'is_cuda') else y
        return net.grad_weighted(sigma, y_t, engine=getattr(self.model, 'engine', 0)) / float(sigma.shape[0])

    def train_step(self):
        """One parameter update = `update_params_frequency` mini-batches of the generator."""
        gen = self.generator
        freq = getattr(gen, 'update_params_frequency', 1)
        grad = None
        for _ in range(freq):
            x, y = next(gen)
            g = self.gradient(x, y)
            grad = g if grad is None else grad.add_(g)
        if self.distributed:
            allreduce_sum_(grad)
        params = self.machine.flat_params_device()
        self.optimizer.step(params, grad)
        self.machine.params_updated()
        energy = getattr(gen, 'current_energy', None)
        self.history.append(energy)
        return energy

    def fit(self, steps, checkpoint_path=None, checkpoint_every_seconds=None, callbacks=(), steps_per_epoch=None,
            initial_step=0):
        """`steps` updates; with a path, a checkpoint every `checkpoint_every_seconds` of wall clock and at the end
        (CheckpointByTime, callbacks/checkpoint.py:9-73).

        `callbacks` (flowket_b200.callbacks) see what Keras' fit_generator shows the reference's callbacks
        (experiments/train.py:117-130): `on_batch_end(batch, logs)` after every update, `on_epoch_begin/end(epoch,
        logs)` every `steps_per_epoch` updates (default: one epoch = one update), `on_train_begin/end`; a callback
        that sets `model.stop_training` (BadEigenStateStopping) ends the loop.  The filled `logs` dicts are kept in
        `self.logs`."""
        import time
        last = time.time()
        steps_per_epoch = 1 if steps_per_epoch is None else int(steps_per_epoch)
        self.model.stop_training = False
        for cb in callbacks:
            cb.set_model(self.model)
            if hasattr(cb, 'set_trainer'):
                cb.set_trainer(self)
            cb.on_train_begin({})
        epoch_logs = {}
        for step in range(initial_step, initial_step + steps):
            epoch, batch = divmod(step, steps_per_epoch)
            if batch == 0:
                epoch_logs = {}
                for cb in callbacks:
                    cb.on_epoch_begin(epoch, epoch_logs)
            self.train_step()
            batch_logs = {}
            for cb in callbacks:
                cb.on_batch_end(batch, batch_logs)
            epoch_logs.update(batch_logs)
            if batch == steps_per_epoch - 1:
                for cb in callbacks:
                    cb.on_epoch_end(epoch, epoch_logs)
                if callbacks:
                    self.logs.append(dict(epoch_logs))
            if checkpoint_path and checkpoint_every_seconds is not None and time.time() - last >= checkpoint_every_seconds:
                self.save_checkpoint(checkpoint_path)
                last = time.time()
            if self.model.stop_training:
                break
        for cb in callbacks:
            cb.on_train_end({})
        if checkpoint_path:
            self.save_checkpoint(checkpoint_path)
        return self.history

    # ---- checkpoint / resume: weights + optimizer slots + sampler counter (callbacks/checkpoint.py:9-73 pickles the
    #      optimizer state next to save_weights; here one .npz) ------------------------------------------------------
    def save_checkpoint(self, path):
        state = {'params': self.machine.flat_params_device().cpu().numpy(), 'steps': np.int64(len(self.history))}
        opt = self.optimizer
        state['opt_t'] = np.int64(getattr(opt, 't', 0))
        for slot in ('m', 'v'):
            if getattr(opt, slot, None) is not None:
                state['opt_' + slot] = getattr(opt, slot).cpu().numpy()
        sampler = getattr(self.generator, 'sampler', None)
        if sampler is not None and hasattr(sampler, '_draws'):
            state['sampler_draws'] = np.int64(sampler._draws)
            state['sampler_seed'] = np.int64(sampler.seed)
        tmp = str(path) + '.tmp.npz'
        np.savez(tmp, **state)
        import os
        os.replace(tmp, str(path) if str(path).endswith('.npz') else str(path) + '.npz')

    def load_checkpoint(self, path):
        import torch
        with np.load(str(path) if str(path).endswith('.npz') else str(path) + '.npz') as f:
            params = self.machine.flat_params_device()
            params.copy_(torch.from_numpy(f['params']).to(params.device))
            self.machine.params_updated()
            opt = self.optimizer
            if hasattr(opt, 't'):
                opt.t = int(f['opt_t'])
            for slot in ('m', 'v'):
                if 'opt_' + slot in f.files:
                    setattr(opt, slot, torch.from_numpy(f['opt_' + slot]).to(params.device))
            sampler = getattr(self.generator, 'sampler', None)
            if sampler is not None and 'sampler_draws' in f.files:
                sampler._draws = int(f['sampler_draws'])
                sampler.seed = int(f['sampler_seed'])
            self.history = [None] * int(f['steps'])
        return self
