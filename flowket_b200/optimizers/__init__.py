from .trainer import Adam, SGD, Trainer, allreduce_sum_, convert_to_accumulate_gradient_optimizer
from .stochastic_reconfiguration import ComplexValuesStochasticReconfiguration, StochasticReconfiguration, \
    conjugate_gradient, sr_delta, real_sr_delta, sample_space_sr_delta, distributed_cholesky_solve
