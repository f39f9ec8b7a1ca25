"""Callbacks of the exact-enumeration path (flowket/callbacks/exact/*.py): same names, arguments and `logs` keys.
With ExactVariational one pass over the 2^N states is `num_of_batch_until_full_cycle` mini-batches, so the
per-batch variants only report on cycle boundaries."""
import time

import numpy

from . import Callback, StatsCallback
from ..exact.utils import fdot


class _ExactStats(StatsCallback):
    cycle_phase = 0      # report when (batch + cycle_phase) is a multiple of the cycle length

    def __init__(self, exact_variational, log_in_batch_or_epoch=True, **kwargs):
        super(_ExactStats, self).__init__(exact_variational, None, log_in_batch_or_epoch, **kwargs)
        self.exact_variational = exact_variational

    def batch_is_due(self, batch):
        return (batch + self.cycle_phase) % self.exact_variational.num_of_batch_until_full_cycle == 0


class ExactLocalEnergy(_ExactStats):
    """energy/energy, energy/local_energy_variance[, energy/relative_error] (exact/local_energy.py:12-17)."""

    def __init__(self, exact_variational, true_ground_state_energy=None, log_in_batch_or_epoch=True, **kwargs):
        super(ExactLocalEnergy, self).__init__(exact_variational, log_in_batch_or_epoch, **kwargs)
        self.true_ground_state_energy = true_ground_state_energy

    def collect(self, logs, generator, prefix=''):
        observable = generator.energy_observable
        energy = numpy.real(observable.current_energy)
        logs['energy/energy'] = energy
        logs['energy/local_energy_variance'] = numpy.real(observable.current_local_energy_variance)
        if self.true_ground_state_energy is not None:
            logs['energy/relative_error'] = (self.true_ground_state_energy - energy) / self.true_ground_state_energy

    def add_energy_to_logs(self, logs):
        self.collect(logs, self.exact_variational)


class ExactSigmaZ(_ExactStats):
    """observables/{sigma_z,abs_sigma_z} = sum_s p(s) m(s) over all states (exact/sigma_z.py:14-22)."""

    def __init__(self, exact_variational, log_in_batch_or_epoch=True, **kwargs):
        super(ExactSigmaZ, self).__init__(exact_variational, log_in_batch_or_epoch, **kwargs)
        states = exact_variational.states
        magnetisation = states.reshape(states.shape[0], -1).sum(axis=1) / float(numpy.prod(states.shape[1:]))
        self._sigma_z_vals = magnetisation
        self._abs_sigma_z_vals = numpy.absolute(magnetisation)

    def collect(self, logs, generator, prefix=''):
        logs['observables/abs_sigma_z'] = fdot(self._abs_sigma_z_vals, generator.probs)
        logs['observables/sigma_z'] = fdot(self._sigma_z_vals, generator.probs)

    def add_sigma_z_logs(self, logs):
        self.collect(logs, self.exact_variational)


class ExactObservableCallback(_ExactStats):
    """observables/<name>: exact expectation of another operator in the current state (exact/observable.py:8-19)."""

    def __init__(self, exact_variational, operator, operator_name, log_in_batch_or_epoch=True, **kwargs):
        super(ExactObservableCallback, self).__init__(exact_variational, log_in_batch_or_epoch, **kwargs)
        from ..optimization.exact_variational import ExactObservable
        self.observable = ExactObservable(exact_variational, operator)
        self.operator_name = operator_name

    def collect(self, logs, generator, prefix=''):
        self.observable.update_local_energy()
        logs['observables/%s' % self.operator_name] = numpy.real(self.observable.current_energy)

    def add_observable_to_logs(self, logs):
        self.collect(logs, self.exact_variational)


class RuntimeStats(_ExactStats):
    """times/{wave_function_update,local_energy,gradients,total}; reported on the *last* mini-batch of a cycle
    (exact/runtime_stats.py:11-24)."""
    cycle_phase = 1

    def collect(self, logs, generator, prefix=''):
        now = time.time()
        logs['times/wave_function_update'] = generator.wave_function_update_end_time - \
            generator.machine_updated_start_time
        logs['times/local_energy'] = generator.local_energy_update_end_time - generator.wave_function_update_end_time
        logs['times/gradients'] = now - generator.local_energy_update_end_time
        logs['times/total'] = now - generator.machine_updated_start_time

    def add_runtime_stats_to_logs(self, logs):
        self.collect(logs, self.exact_variational)


class MachineUpdated(Callback):
    """Re-evaluate psi (and optionally the local energies) of all 2^N states after a parameter update
    (exact/machine_updated.py:4-23)."""

    def __init__(self, exact_variational, update_in_batch_or_epoch=True, update_local_energy=True, **kwargs):
        super(MachineUpdated, self).__init__(**kwargs)
        self.exact_variational = exact_variational
        self.update_in_batch_or_epoch = update_in_batch_or_epoch
        self.update_local_energy = update_local_energy

    def _refresh(self):
        if self.update_local_energy:
            self.exact_variational.machine_updated()
        else:
            self.exact_variational._update_wave_function_arrays()

    def on_batch_end(self, batch, logs=None):
        if self.update_in_batch_or_epoch:
            self._refresh()

    def on_epoch_end(self, epoch, logs=None):
        if not self.update_in_batch_or_epoch:
            self._refresh()


def default_wave_function_callbacks_factory(generator, true_ground_state_energy=None, log_in_batch_or_epoch=True):
    """exact/__init__.py:8-12."""
    return [ExactLocalEnergy(generator, true_ground_state_energy=true_ground_state_energy,
                             log_in_batch_or_epoch=log_in_batch_or_epoch),
            ExactSigmaZ(generator, log_in_batch_or_epoch=log_in_batch_or_epoch),
            RuntimeStats(generator, log_in_batch_or_epoch=log_in_batch_or_epoch)]
