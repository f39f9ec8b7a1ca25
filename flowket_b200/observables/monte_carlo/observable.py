"""Observable protocol of flowket/observables/monte_carlo/observable.py:5-23: `local_values(psi, configurations)` gives one
(complex) value per configuration; `estimate` adds the Monte-Carlo mean and the variance of the real part -- the triple
VariationalMonteCarlo and the stats callbacks consume."""
import numpy


class BaseObservable(object):
    def local_values(self, wave_function, configurations):
        raise NotImplementedError

    def estimate(self, wave_function, configurations):
        values = numpy.asarray(self.local_values(wave_function, configurations))
        return values.mean(), numpy.real(values).var(), values


class LambdaObservable(BaseObservable):
    """observable given as a function (psi, configurations) -> per-configuration values (SigmaZ, AbsSigmaZ)"""

    def __init__(self, observable_function):
        self.observable_function = observable_function

    def local_values(self, wave_function, configurations):
        return self.observable_function(wave_function, configurations)
