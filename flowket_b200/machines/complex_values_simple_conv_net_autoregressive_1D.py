"""ComplexValuesSimpleConvNetAutoregressive1D: same constructor as
flowket/machines/complex_values_simple_conv_net_autoregressive_1D.py:24-37.  Complex weights are stored as
(real, imag) pairs with W = real - i*imag (layers/complex/base_layer.py:18-35); on the device a complex conv is
one real conv over channels [Re | Im] with the block matrix [[Wr, Wi], [-Wi, Wr]]."""
from .abstract_machine import AutoNormalizedAutoregressiveMachine
from .. import _lib


class ComplexValuesSimpleConvNetAutoregressive1D(AutoNormalizedAutoregressiveMachine):
    def __init__(self, keras_input_layer, depth, num_of_channels, kernel_size=3, use_dilation=True,
                 max_dilation_rate=None, activation='lncosh', use_float64_ops=False, seed=None, **kwargs):
        if use_float64_ops:
            raise NotImplementedError('complex128 layers are not implemented on the B200 path')
        if not (activation == 'lncosh' or getattr(activation, '__name__', '') == 'lncosh'):
            raise NotImplementedError('only the default lncosh activation is implemented on the B200 path')
        if len(keras_input_layer.shape) != 1:
            raise ValueError('needs a 1-D spin input')
        self.depth = depth
        self.num_of_channels = num_of_channels
        self.kernel_size = kernel_size
        self.use_dilation = use_dilation
        self.max_dilation_rate = max_dilation_rate
        self.activation = 'lncosh'
        self.use_float64_ops = use_float64_ops
        self.exponential_norm = False
        self._seed = seed
        super(ComplexValuesSimpleConvNetAutoregressive1D, self).__init__(keras_input_layer, **kwargs)

    def weight_specs(self):
        specs = []
        C = self.num_of_channels
        shapes = [(self.kernel_size, 1 if i == 0 else C, C) for i in range(self.depth - 1)] + [(1, C, 2)]
        for idx, shape in enumerate(shapes):
            name = 'complex_conv1d' if idx == 0 else 'complex_conv1d_%d' % idx
            specs.append((name + '/kernel_real:0', shape, 'glorot_normal'))
            specs.append((name + '/kernel_imag:0', shape, 'neg_glorot_normal'))
            specs.append((name + '/bias_real:0', (shape[-1],), 'zeros'))
            specs.append((name + '/bias_imag:0', (shape[-1],), 'zeros'))
        return specs

    def _create_args(self):
        (N,) = self.keras_input_layer.shape
        max_dil = self.max_dilation_rate if (self.use_dilation and self.max_dilation_rate) else 0
        return (_lib.FK_NET_CCONV1D, 1, N, self.depth, self.num_of_channels, self.kernel_size, max_dil, 0)
