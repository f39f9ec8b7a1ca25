"""flowket_b200 -- B200-native VMC inner loop behind the FlowKet API.

    from flowket_b200 import Input, Model
    from flowket_b200.machines import ConvNetAutoregressive2D
    from flowket_b200.operators import Heisenberg
    from flowket_b200.samplers import FastAutoregressiveSampler
    from flowket_b200.optimization import VariationalMonteCarlo

The arithmetic of the hot path lives in libflowket_b200.so (hand-written sm_100a CUDA, C ABI in
include/flowket_b200.h); there is no CPU fallback."""
from .keras_shim import Input, Model
from ._lib import FlowketB200Error, FK_ENGINE_FP32, FK_ENGINE_TC, FK_ENGINE_TC_EXACT

__all__ = ['Input', 'Model', 'FlowketB200Error', 'FK_ENGINE_FP32', 'FK_ENGINE_TC', 'FK_ENGINE_TC_EXACT']
