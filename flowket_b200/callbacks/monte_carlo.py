"""Monte-Carlo stats callbacks: same class names, constructor arguments and `logs` keys as
flowket/callbacks/monte_carlo/{local_energy_stats,observable,runtime_stats,mcmc_stats,generator_iterator,
bad_eigen_state_stopping}.py and the factory of flowket/callbacks/monte_carlo/__init__.py:10-29.
They read the attributes VariationalMonteCarlo publishes after every batch (current_energy,
current_local_energy_variance, current_local_energy, current_batch, wave_function, the three time stamps)."""
import time
import warnings

import numpy

from . import Callback, StatsCallback, TensorBoard


class LocalEnergyStats(StatsCallback):
    """energy/energy, energy/local_energy_variance[, energy/relative_error] (local_energy_stats.py:15-20)."""

    def __init__(self, generator, validation_generator=None, true_ground_state_energy=None,
                 log_in_batch_or_epoch=True, validation_period=1, **kwargs):
        super(LocalEnergyStats, self).__init__(generator, validation_generator, log_in_batch_or_epoch,
                                               validation_period, **kwargs)
        self.true_ground_state_energy = true_ground_state_energy

    def collect(self, logs, generator, prefix=''):
        energy = numpy.real(generator.current_energy)
        logs[prefix + 'energy/energy'] = energy
        logs[prefix + 'energy/local_energy_variance'] = numpy.real(generator.current_local_energy_variance)
        if self.true_ground_state_energy is not None:
            exact = self.true_ground_state_energy
            logs[prefix + 'energy/relative_error'] = (exact - energy) / exact

    add_energy_stats_to_logs = collect


class ObservableStats(StatsCallback):
    """observables/<name>: Monte-Carlo estimate of an observable on the generator's current batch
    (observable.py:16-18; the misspelt argument name is the reference's)."""

    def __init__(self, generator, observable, observabler_name, validation_generator=None,
                 log_in_batch_or_epoch=True, validation_period=1, **kwargs):
        super(ObservableStats, self).__init__(generator, validation_generator, log_in_batch_or_epoch,
                                              validation_period, **kwargs)
        self.observable = observable
        self.observable_name = observabler_name

    def collect(self, logs, generator, prefix=''):
        value = self.observable.estimate(generator.wave_function, generator.current_batch)[0]
        logs['%sobservables/%s' % (prefix, self.observable_name)] = numpy.real(value)

    add_observable_stats_to_logs = collect


class RuntimeStats(StatsCallback):
    """times/{sampling,local_energy,gradients,total} from the generator's time stamps (runtime_stats.py:11-16)."""

    def __init__(self, generator, log_in_batch_or_epoch=True, **kwargs):
        super(RuntimeStats, self).__init__(generator, None, log_in_batch_or_epoch, **kwargs)

    def collect(self, logs, generator, prefix=''):
        now = time.time()
        logs['times/sampling'] = generator.sampling_end_time - generator.start_time
        logs['times/local_energy'] = generator.local_energy_end_time - generator.sampling_end_time
        logs['times/gradients'] = now - generator.local_energy_end_time
        logs['times/total'] = now - generator.start_time

    def add_runtime_stats_to_logs(self, logs):
        self.collect(logs, self.generator)


class MCMCStats(StatsCallback):
    """mcmc/* chain diagnostics of a Metropolis-Hastings sampler (mcmc_stats.py:11-17)."""

    def __init__(self, generator, log_in_batch_or_epoch=True, **kwargs):
        super(MCMCStats, self).__init__(generator, None, log_in_batch_or_epoch, **kwargs)

    def collect(self, logs, generator, prefix=''):
        sampler = generator.sampler
        r_hat, _, correlations_sum, effective_sample_size = sampler.calc_r_hat_value(
            numpy.real(generator.current_local_energy))
        logs['mcmc/acceptance_ratio'] = sampler.acceptance_ratio
        logs['mcmc/energy_r_hat'] = r_hat
        logs['mcmc/energy_effective_sample_size'] = effective_sample_size
        logs['mcmc/energy_correlations_sum'] = correlations_sum

    def add_mcmc_logs(self, logs):
        self.collect(logs, self.generator)


class GeneratorIterator(Callback):
    """Advance a (validation) generator every `period` epochs (generator_iterator.py:4-12)."""

    def __init__(self, generator, period=1, **kwargs):
        super(GeneratorIterator, self).__init__(**kwargs)
        self.generator = generator
        self.period = period

    def on_epoch_end(self, epoch, logs=None):
        if epoch % self.period == 0:
            next(self.generator)


class BadEigenStateStopping(Callback):
    """Stop when the machine has collapsed onto an excited eigenstate: tiny local-energy variance while the
    energy is still far above the known upper bound (bad_eigen_state_stopping.py:17-36)."""

    def __init__(self, ground_state_energy_upper_bound, variance_tol=1e-2, relative_error_to_stop=0.1, min_epoch=10,
                 **kwargs):
        super(BadEigenStateStopping, self).__init__(**kwargs)
        self.ground_state_energy_upper_bound = ground_state_energy_upper_bound
        self.variance_tol = variance_tol
        self.relative_error_to_stop = relative_error_to_stop
        self.min_epoch = min_epoch
        self.stopped_epoch = None

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        for prefix in ('val_', ''):
            if prefix + 'energy/energy' in logs:
                energy = logs[prefix + 'energy/energy']
                variance = logs[prefix + 'energy/local_energy_variance']
                break
        else:
            warnings.warn("Can't find local energy stats, skipping bad eigen state early stopping", RuntimeWarning)
            return
        if epoch < self.min_epoch:
            return
        bound = self.ground_state_energy_upper_bound
        if (energy - bound) / numpy.abs(bound) > self.relative_error_to_stop and variance < self.variance_tol:
            if self.model is not None:
                self.model.stop_training = True
            self.stopped_epoch = epoch


class TensorBoardWithGeneratorValidationData(TensorBoard):
    """Constructor of callbacks/monte_carlo/tensorboard_with_generator_validation_data.py:5-17; the reference feeds the
    generator's current batch to Keras' histogram summaries, which this scalar logger does not write."""

    def __init__(self, generator, **kwargs):
        super(TensorBoardWithGeneratorValidationData, self).__init__(**kwargs)
        self.generator = generator


def default_wave_function_stats_callbacks_factory(generator, validation_generator=None, true_ground_state_energy=None,
                                                  log_in_batch_or_epoch=True, validation_period=1):
    """[GeneratorIterator,] LocalEnergyStats, sigma_z, abs_sigma_z, RuntimeStats -- the list every reference
    script passes to fit_generator (callbacks/monte_carlo/__init__.py:10-29)."""
    from ..observables.monte_carlo import SigmaZ, AbsSigmaZ
    shared = dict(validation_generator=validation_generator, log_in_batch_or_epoch=log_in_batch_or_epoch,
                  validation_period=validation_period)
    callbacks = []
    if validation_generator is not None:
        callbacks.append(GeneratorIterator(validation_generator, period=validation_period))
    callbacks.append(LocalEnergyStats(generator, true_ground_state_energy=true_ground_state_energy, **shared))
    callbacks.append(ObservableStats(generator, SigmaZ(), 'sigma_z', **shared))
    callbacks.append(ObservableStats(generator, AbsSigmaZ(), 'abs_sigma_z', **shared))
    callbacks.append(RuntimeStats(generator, log_in_batch_or_epoch=log_in_batch_or_epoch))
    return callbacks
