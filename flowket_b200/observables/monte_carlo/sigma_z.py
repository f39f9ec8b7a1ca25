"""flowket/observables/monte_carlo/sigma_z.py:7-20 (host reductions; not on the device hot path)."""
import functools

import numpy

from .observable import LambdaObservable


def abs_sigma_z(wave_function, configurations):
    configurations = numpy.asarray(configurations)
    total = numpy.prod(configurations.shape[1:])
    return numpy.absolute(configurations.sum(axis=tuple(range(1, configurations.ndim)))) / total


def sigma_z(wave_function, configurations):
    configurations = numpy.asarray(configurations)
    total = numpy.prod(configurations.shape[1:])
    return configurations.sum(axis=tuple(range(1, configurations.ndim))) / total


AbsSigmaZ = functools.partial(LambdaObservable, observable_function=abs_sigma_z)
SigmaZ = functools.partial(LambdaObservable, observable_function=sigma_z)
