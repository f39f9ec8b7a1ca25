from .abstract_machine import Machine, AutoregressiveMachine, AutoNormalizedAutoregressiveMachine
from .conv_net_autoregressive_2D import ConvNetAutoregressive2D
from .simple_conv_net_autoregressive_1D import SimpleConvNetAutoregressive1D
from .complex_values_simple_conv_net_autoregressive_1D import ComplexValuesSimpleConvNetAutoregressive1D
from .ensemble import (EnsembleModel, make_2d_obc_invariants, make_up_down_invariant, make_pbc_invariants,
                       probabilistic_ensemble_op, average_ensemble_op)
