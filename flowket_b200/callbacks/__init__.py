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


from .checkpoint import CheckpointByTime  # noqa: E402
from . import monte_carlo, exact  # noqa: E402,F401
from .monte_carlo import default_wave_function_stats_callbacks_factory  # noqa: E402,F401

__all__ = ['Callback', 'StatsCallback', 'CheckpointByTime', 'monte_carlo', 'exact',
           'default_wave_function_stats_callbacks_factory']
