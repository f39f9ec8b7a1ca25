"""Metric callbacks of the VMC loop: the reference hangs these on Keras'
`fit_generator` (flowket/callbacks/monte_carlo/*.py, flowket/callbacks/exact/*.py); here they hang on
`Trainer.fit(callbacks=[...])` and `evaluation.evaluate`, with the same constructor arguments and the same
keys in `logs`.  Host-only bookkeeping: every number they report was produced on the device by the generator."""


class Callback(object):
    """The part of keras.callbacks.Callback these callbacks use."""

    def __init__(self, **_unused):
        self.model = None
        self.params = None

    def set_model(self, model):
        self.model = model

    def set_params(self, params):
        self.params = params

    def on_train_begin(self, logs=None):
        pass

    def on_train_end(self, logs=None):
        pass

    def on_epoch_begin(self, epoch, logs=None):
        pass

    def on_epoch_end(self, epoch, logs=None):
        pass

    def on_batch_begin(self, batch, logs=None):
        pass

    def on_batch_end(self, batch, logs=None):
        pass


class StatsCallback(Callback):
    """Shared gating of every stats callback of the reference: write into `logs` after each batch
    (`log_in_batch_or_epoch=True`) or after each epoch, and every `validation_period` epochs repeat the
    measurement on the validation generator under the 'val_' prefix (local_energy_stats.py:22-34,
    observable.py:20-32 follow the same pattern)."""

    def __init__(self, generator, validation_generator=None, log_in_batch_or_epoch=True, validation_period=1,
                 **kwargs):
        super(StatsCallback, self).__init__(**kwargs)
        self.generator = generator
        self.validation_generator = validation_generator
        self.log_in_batch_or_epoch = log_in_batch_or_epoch
        self.validation_period = validation_period

    def collect(self, logs, generator, prefix=''):
        raise NotImplementedError

    def batch_is_due(self, batch):
        return True

    def on_batch_end(self, batch, logs=None):
        if self.log_in_batch_or_epoch and self.batch_is_due(batch):
            self.collect({} if logs is None else logs, self.generator)

    def on_epoch_end(self, epoch, logs=None):
        logs = {} if logs is None else logs
        if not self.log_in_batch_or_epoch:
            self.collect(logs, self.generator)
        if self.validation_generator is not None and epoch % self.validation_period == 0:
            self.collect(logs, self.validation_generator, prefix='val_')


class TensorBoard(Callback):
    """Scalar logger with the constructor of flowket.callbacks.TensorBoard (callbacks/tensorboard.py; keras TensorBoard
    arguments are accepted).  The TensorFlow event-file writer, histograms and graph dumps of the reference are not
    rebuilt: every numeric entry of `logs` is appended to `<log_dir>/scalars.jsonl` as {"step": n, "tag": value, ...} --
    per batch (`update_freq` = 'batch' or an integer period in batches) or per epoch (`update_freq='epoch'`)."""

    def __init__(self, log_dir='./logs', update_freq='epoch', **_keras_arguments):
        super(TensorBoard, self).__init__()
        self.log_dir = log_dir
        self.update_freq = update_freq
        self._file = None
        self._batches_seen = 0

    def _write(self, step, logs):
        import json
        import os
        import numbers
        import numpy
        scalars = {k: float(numpy.real(v)) for k, v in (logs or {}).items()
                   if isinstance(v, (numbers.Number, numpy.number)) and k not in ('batch', 'size')}
        if not scalars:
            return
        if self._file is None:
            os.makedirs(self.log_dir, exist_ok=True)
            self._file = open(os.path.join(self.log_dir, 'scalars.jsonl'), 'a')
        self._file.write(json.dumps(dict(step=int(step), **scalars)) + '\n')
        self._file.flush()

    def on_batch_end(self, batch, logs=None):
        self._batches_seen += 1
        if self.update_freq == 'epoch':
            return
        period = 1 if self.update_freq == 'batch' else int(self.update_freq)
        if self._batches_seen % period == 0:
            self._write(self._batches_seen, logs)

    def on_epoch_end(self, epoch, logs=None):
        if self.update_freq == 'epoch':
            self._write(epoch, logs)

    def on_train_end(self, logs=None):
        if self._file is not None:
            self._file.close()
            self._file = None


class TerminateOnNaN(Callback):
    """keras.callbacks.TerminateOnNaN as the reference scripts use it (examples/j1j2_2d_monte_carlo_4.py:61): stop when
    the monitored quantity stops being finite.  Keras watches `logs['loss']`; this path has no scalar loss on the host, so
    the energy entries of `logs` are watched instead."""

    def __init__(self, keys=('energy/energy', 'energy/local_energy_variance', 'loss'), **kwargs):
        super(TerminateOnNaN, self).__init__(**kwargs)
        self.keys = tuple(keys)
        self.stopped = False

    def _check(self, logs):
        import numpy
        for key in self.keys:
            if key in (logs or {}) and not numpy.all(numpy.isfinite(logs[key])):
                self.stopped = True
                if self.model is not None:
                    self.model.stop_training = True

    def on_batch_end(self, batch, logs=None):
        self._check(logs)

    def on_epoch_end(self, epoch, logs=None):
        self._check(logs)


from .checkpoint import CheckpointByTime  # noqa: E402
from . import monte_carlo, exact  # noqa: E402,F401
from .monte_carlo import default_wave_function_stats_callbacks_factory  # noqa: E402,F401

__all__ = ['Callback', 'StatsCallback', 'TensorBoard', 'TerminateOnNaN', 'CheckpointByTime', 'monte_carlo', 'exact',
           'default_wave_function_stats_callbacks_factory']
