"""Samplers that draw from an enumerated distribution over all 2^N states
(flowket/samplers/exact_sampler.py:7-38).  Host-only: the probabilities come from ExactVariational, whose
log psi table was filled by the CUDA forward."""
import numpy

from . import Sampler
from ..exact.utils import decimal_array_to_binary_array


class _EnumeratedSampler(Sampler):
    def __init__(self, input_size, batch_size, seed=None, **kwargs):
        super(_EnumeratedSampler, self).__init__(input_size, batch_size, **kwargs)
        self.rng = numpy.random.default_rng(seed)

    def _distribution(self):
        """-> (probabilities over state indices, number of spins)"""
        raise NotImplementedError

    def __next__(self):
        probs, number_of_spins = self._distribution()
        indices = self.rng.choice(probs.shape[0], size=self.batch_size, p=probs)
        spins = decimal_array_to_binary_array(indices, num_of_bits=number_of_spins)
        return spins.reshape((self.batch_size,) + self.input_size)


class ExactSampler(_EnumeratedSampler):
    """Draws from |psi|^2 of an ExactVariational (always its *current* `probs`, exact_sampler.py:15-20)."""

    def __init__(self, exact_variational, batch_size, **kwargs):
        super(ExactSampler, self).__init__(exact_variational.input_size, batch_size, **kwargs)
        self.exact_variational = exact_variational

    def _distribution(self):
        return self.exact_variational.probs, self.exact_variational.number_of_spins


class WaveFunctionSampler(_EnumeratedSampler):
    """Draws from a fixed, normalised vector of log-amplitudes (exact_sampler.py:23-38)."""

    def __init__(self, wave_function_vector, input_size, batch_size, **kwargs):
        super(WaveFunctionSampler, self).__init__(input_size, batch_size, **kwargs)
        self.wave_function_vector = numpy.asarray(wave_function_vector)
        self.log_probs = numpy.real(self.wave_function_vector) * 2.0
        self.probs = numpy.exp(self.log_probs)

    def _distribution(self):
        return self.probs, int(round(numpy.log2(self.probs.shape[0])))
