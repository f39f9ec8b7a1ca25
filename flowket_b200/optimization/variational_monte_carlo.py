"""VariationalMonteCarlo: constructor, methods and attributes of
flowket/optimization/variational_monte_carlo.py:12-50.  sample -> E_loc stays on the device
(FastAutoregressiveSampler.next_device -> Observable.local_values_device); the host sees the same
ndarrays the reference returns."""
import functools
import time

import numpy

from .mini_batch_generator import MiniBatchGenerator
from ..observables.monte_carlo import Observable, BaseObservable


class VariationalMonteCarlo(MiniBatchGenerator):
    def __init__(self, model, operator, sampler, mini_batch_size=None, wave_function_evaluation_batch_size=None,
                 **kwargs):
        super(VariationalMonteCarlo, self).__init__(sampler.batch_size, mini_batch_size, **kwargs)
        if wave_function_evaluation_batch_size is None:
            wave_function_evaluation_batch_size = self.mini_batch_size
        self.model = model
        self.operator = operator
        self.sampler = sampler
        self.current_batch = None
        self.current_batch_device = None
        self.wave_function = functools.partial(self.model.predict, batch_size=wave_function_evaluation_batch_size)
        self.energy_observable = operator
        if not isinstance(self.energy_observable, BaseObservable):
            self.energy_observable = Observable(operator)

    def set_sampler(self, sampler, mini_batch_size=None):
        self.sampler = sampler
        return self.set_batch_size(sampler.batch_size, mini_batch_size)

    def _estimate(self):
        """(mean, var(Re), local values) -- device route when the observable and sampler support it."""
        obs = self.energy_observable
        if isinstance(obs, Observable) and self.current_batch_device is not None:
            from ..keras_shim import Model
            from ..machines.ensemble import EnsembleModel
            eloc = None
            if isinstance(self.model, Model):
                eloc = obs.local_values_device(self.model, self.current_batch_device)
            elif isinstance(self.model, EnsembleModel):
                # a symmetrised wave function must be evaluated through the ensemble's own predict (the reference
                # evaluates model.predict of the ensemble model), never through the base machine's fused kernel
                eloc = obs.local_values_device_generic(self.model.predict_device, self.current_batch_device)
            if eloc is not None:
                self.current_local_energy_device = eloc
                lv = eloc.cpu().numpy()
                return numpy.mean(lv), numpy.var(numpy.real(lv)), lv
        return obs.estimate(self.wave_function, self.current_batch)

    def _update_batch_local_energy(self):
        self.current_energy, self.current_local_energy_variance, self.current_local_energy = self._estimate()

    def loss_coefficients(self):
        return numpy.conj(self.current_local_energy - self.current_energy)

    def _draw(self):
        self.start_time = time.time()
        if hasattr(self.sampler, 'next_device'):
            self.current_batch_device = self.sampler.next_device()
            self.current_batch = self.current_batch_device.cpu().numpy()
        else:
            self.current_batch_device = None
            self.current_batch = next(self.sampler)
        self.sampling_end_time = time.time()

    # ---- split solve of the sharded SR step (optimizers/sample_space_sr.py): the optimizer evaluates the local energies
    # itself, dealt over the ranks next to the factorisation, and hands this rank's values back
    def next_samples(self):
        """sampling only -> the batch (host ndarray); the local energies follow through set_local_energy"""
        self._draw()
        self.current_local_energy = None
        return self.current_batch

    def local_energy_function(self):
        """the local energies of THIS generator's model and operator as a function of the samples (PerSampleLocalEnergy)"""
        return self.energy_observable.per_sample_device(self.model)

    def _accept_local_values(self, lv):
        self.current_energy, self.current_local_energy_variance, self.current_local_energy = \
            numpy.mean(lv), numpy.var(numpy.real(lv)), lv

    def set_local_energy(self, local_energy_device):
        """the local energies of the current batch, evaluated elsewhere (complex128 device tensor [batch])"""
        self.current_local_energy_device = local_energy_device
        self._accept_local_values(local_energy_device.cpu().numpy())
        self.local_energy_end_time = time.time()

    def next_batch(self):
        self._draw()
        self._update_batch_local_energy()
        self.local_energy_end_time = time.time()
        return self.current_batch, self.loss_coefficients() / self.batch_size
