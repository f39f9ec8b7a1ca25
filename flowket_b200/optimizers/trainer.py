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


def convert_to_accumulate_gradient_optimizer(orig_optimizer, update_params_frequency, accumulate_sum_or_mean=True,
                                             ema_decay=0, use_horovod=False):
    """Same call as flowket/optimizers/accumulate_gradient_optimizer.py:13-88: gradients of `update_params_frequency`
    consecutive mini-batches are summed (or averaged) and the wrapped optimizer steps once per group; with `ema_decay`
    an exponential moving average of the parameters is tracked after every step and `set_weights_ema()` installs its
    bias-corrected value (:28-29,66-68).  `use_horovod` asks the Trainer to sum the accumulated gradient over the ranks
    (NCCL allreduce) before the step (:31-50).  Here the optimizer is only annotated; Trainer.train_on_batch does the
    accumulation on the device."""
    if update_params_frequency < 1:
        raise ValueError('update_params_frequency must be >= 1')
    orig_optimizer.update_params_frequency = int(update_params_frequency)
    orig_optimizer.accumulate_sum_or_mean = bool(accumulate_sum_or_mean)
    orig_optimizer.accumulated_iterations = 0
    orig_optimizer.ema_decay = float(ema_decay)
    orig_optimizer.total_iterations = 0
    orig_optimizer.params_ema = None
    orig_optimizer.use_horovod = bool(use_horovod)

    def set_update_params_frequency(frequency):
        orig_optimizer.update_params_frequency = int(frequency)

    def track_ema(params):
        if orig_optimizer.ema_decay <= 0:
            return
        if orig_optimizer.params_ema is None:
            orig_optimizer.params_ema = params.new_zeros(params.shape)
        orig_optimizer.total_iterations += 1
        orig_optimizer.params_ema.mul_(orig_optimizer.ema_decay).add_(params, alpha=1 - orig_optimizer.ema_decay)

    def set_weights_ema(machine=None):
        """params <- ema / (1 - decay^t); needs the machine the optimizer trains (remembered by the Trainer)."""
        machine = machine if machine is not None else getattr(orig_optimizer, '_machine', None)
        if machine is None or orig_optimizer.params_ema is None:
            raise RuntimeError('no moving average yet: train at least one step with ema_decay > 0')
        correction = 1.0 - orig_optimizer.ema_decay ** orig_optimizer.total_iterations
        machine.flat_params_device().copy_(orig_optimizer.params_ema / correction)
        machine.params_updated()

    orig_optimizer.set_update_params_frequency = set_update_params_frequency
    orig_optimizer.track_ema = track_ema
    orig_optimizer.set_weights_ema = set_weights_ema
    return orig_optimizer


class Trainer(object):
    def __init__(self, model, generator, optimizer, distributed=False):
        self.model, self.generator, self.optimizer, self.distributed = model, generator, optimizer, distributed
        self.machine = model.machine
        self.history = []
        self.logs = []          # one dict per epoch, filled by the callbacks of fit()
        self._accumulator, self._accumulated = None, 0

    def gradient(self, x, y):
        """(1/mb) d/dtheta sum_b 2 Re(log psi_b y_b) on the device."""
        import torch
        net = self.machine.device_net()
        sigma = net.to_sigma(x)
        y_t = torch.as_tensor(np.asarray(y, np.complex64)) if not hasattr(y, 'is_cuda') else y
        return net.grad_weighted(sigma, y_t, engine=getattr(self.model, 'engine', 0)) / float(sigma.shape[0])

    def _frequency(self):
        """mini-batches per parameter update: the optimizer's (convert_to_accumulate_gradient_optimizer) or the generator's"""
        freq = getattr(self.optimizer, 'update_params_frequency', None)
        if freq is None:
            freq = getattr(self.generator, 'update_params_frequency', 1)
        return int(freq)

    def train_on_batch(self, x, y):
        """Keras' train_on_batch under an accumulate-gradient optimizer (accumulate_gradient_optimizer.py:54-82): add this
        mini-batch's gradient to the accumulator; every `update_params_frequency`-th call allreduce (if distributed), step,
        reset.  Returns True when the parameters moved."""
        if hasattr(self.optimizer, 'compute_update'):
            return self._stochastic_reconfiguration_step(x, y)
        g = self.gradient(x, y)
        freq = self._frequency()
        if not getattr(self.optimizer, 'accumulate_sum_or_mean', True):
            g = g / float(freq)
        self._accumulator = g if self._accumulator is None else self._accumulator.add_(g)
        self._accumulated += 1
        if hasattr(self.optimizer, 'accumulated_iterations'):
            self.optimizer.accumulated_iterations += 1
        if self._accumulated % freq != 0:
            return False
        grad, self._accumulator, self._accumulated = self._accumulator, None, 0
        if self.distributed or getattr(self.optimizer, 'use_horovod', False):
            allreduce_sum_(grad)
        params = self.machine.flat_params_device()
        self.optimizer.step(params, grad)
        self.machine.params_updated()
        if getattr(self.optimizer, 'ema_decay', 0) > 0:
            self.optimizer._machine = self.machine
            self.optimizer.track_ema(params)
        return True

    def _split_solve_route(self):
        opt, gen = self.optimizer, self.generator
        if not (hasattr(opt, 'step_generator') and getattr(opt, 'distributed', False) and getattr(opt, 'split_solve', True)):
            return False
        if not (hasattr(gen, 'next_samples') and hasattr(getattr(gen, 'sampler', None), 'next_device')):
            return False
        try:
            import torch.distributed as dist
            world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        except Exception:
            world = 1
        from ..keras_shim import Model
        from ..observables.monte_carlo import Observable
        return world > 1 and self._frequency() == 1 and isinstance(getattr(gen, 'model', None), Model) and \
            isinstance(getattr(gen, 'energy_observable', None), Observable)

    def _stochastic_reconfiguration_step(self, x, y):
        """`model.compile(optimizer=ComplexValuesStochasticReconfiguration(model, ...))` (the reference passes its SR
        optimizer to Keras like any other, optimizers/stochastic_reconfiguration/optimizer.py:14-31): the optimizer owns
        the whole update -- per-sample Jacobians, SR system, solve, W <- W - lr * delta -- from the batch (sigma, y_true).
        SR needs the whole batch at once (the reference asserts the same through its Jacobian shapes)."""
        if self._frequency() != 1:
            raise ValueError('stochastic reconfiguration needs mini_batch_size == batch_size (one update per batch)')
        opt = self.optimizer
        if hasattr(opt, 'complex_jacobian'):
            opt.step(x, y)                                   # complex-parameter machines: takes y_true like the reference
        else:
            # real-parameter SR is written in terms of the centred local energies.  y_true = conj(E_loc - E) / B_global
            # carries them, but B_global != len(y) on a multi-GPU generator, so take them from the generator itself
            gen = self.generator
            if getattr(gen, 'current_local_energy', None) is not None and len(gen.current_local_energy) == len(y):
                e_centred = np.asarray(gen.current_local_energy) - gen.current_energy
            else:
                e_centred = np.conj(np.asarray(y)) * float(getattr(gen, 'global_batch_size', len(y)))
            opt.step(x, e_centred)
        return True

    def train_step(self):
        """One parameter update = `update_params_frequency` mini-batches of the generator."""
        gen = self.generator
        if self._split_solve_route():
            # sharded real-parameter SR: the optimizer evaluates the local energies itself, dealt over the ranks next to the
            # factorisation of the SR matrix (optimizers/sample_space_sr.py), and hands this rank's values back to the generator
            self.optimizer.step_generator(gen)
            energy = getattr(gen, 'current_energy', None)
            self.history.append(energy)
            return energy
        updated = False
        while not updated:
            x, y = next(gen)
            updated = self.train_on_batch(x, y)
        energy = getattr(gen, 'current_energy', None)
        self.history.append(energy)
        return energy

    def fit_generator(self, generator=None, steps_per_epoch=1, epochs=1, callbacks=(), initial_epoch=0, verbose=0,
                      **_keras_only):
        """Keras' Model.fit_generator as the reference scripts call it (examples/*.py, experiments/train.py:127-128):
        `steps_per_epoch` counts *mini-batches*, epochs run from `initial_epoch` to `epochs`, callbacks see
        on_batch_end(batch, logs) after every mini-batch and on_epoch_end(epoch, logs) after every epoch.
        max_queue_size / workers are accepted and ignored (the generator runs on this thread, like workers=0)."""
        gen = self.generator if generator is None else generator
        self.model.stop_training = False
        for cb in callbacks:
            cb.set_model(self.model)
            if hasattr(cb, 'set_trainer'):
                cb.set_trainer(self)
            cb.set_params({'epochs': epochs, 'steps': steps_per_epoch, 'verbose': verbose})
            cb.on_train_begin({})
        for epoch in range(initial_epoch, epochs):
            epoch_logs = {}
            for cb in callbacks:
                cb.on_epoch_begin(epoch, epoch_logs)
            for batch in range(steps_per_epoch):
                x, y = next(gen)
                batch_logs = {'batch': batch, 'size': len(x)}
                for cb in callbacks:
                    cb.on_batch_begin(batch, batch_logs)
                if self.train_on_batch(x, y):
                    self.history.append(getattr(self.generator, 'current_energy', None))
                for cb in callbacks:
                    cb.on_batch_end(batch, batch_logs)
                epoch_logs.update({k: v for k, v in batch_logs.items() if k not in ('batch', 'size')})
                if self.model.stop_training:
                    break
            for cb in callbacks:
                cb.on_epoch_end(epoch, epoch_logs)
            self.logs.append(dict(epoch_logs))
            if verbose:
                print('epoch %d/%d  %s' % (epoch + 1, epochs, '  '.join(
                    '%s %.6g' % (k, v) for k, v in sorted(epoch_logs.items()) if not k.startswith('times/'))), flush=True)
            if self.model.stop_training:
                break
        for cb in callbacks:
            cb.on_train_end({})
        return self.logs

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
        if self._accumulated:
            raise RuntimeError('checkpoint requested in the middle of a gradient accumulation window')
        if getattr(opt, 'params_ema', None) is not None:     # convert_to_accumulate_gradient_optimizer(ema_decay > 0)
            state['opt_params_ema'] = opt.params_ema.cpu().numpy()
        for counter in ('total_iterations', 'accumulated_iterations'):
            if hasattr(opt, counter):
                state['opt_' + counter] = np.int64(getattr(opt, counter))
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
            if 'opt_params_ema' in f.files:
                opt.params_ema = torch.from_numpy(f['opt_params_ema']).to(params.device)
            for counter in ('total_iterations', 'accumulated_iterations'):
                if 'opt_' + counter in f.files:
                    setattr(opt, counter, int(f['opt_' + counter]))
            sampler = getattr(self.generator, 'sampler', None)
            if sampler is not None and 'sampler_draws' in f.files:
                sampler._draws = int(f['sampler_draws'])
                sampler.seed = int(f['sampler_seed'])
            self.history = [None] * int(f['steps'])
        return self
