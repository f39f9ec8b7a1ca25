"""SimpleConvNetAutoregressive1D: same constructor as flowket/machines/simple_conv_net_autoregressive_1D.py:26-40."""
from .abstract_machine import AutoNormalizedAutoregressiveMachine
from .. import _lib


class SimpleConvNetAutoregressive1D(AutoNormalizedAutoregressiveMachine):
    def __init__(self, keras_input_layer, depth, num_of_channels, kernel_size=3, use_dilation=True,
                 add_skip_connections=False, max_dilation_rate=None, activation='relu', weights_normalization=True,
                 should_expand_input_dim=True, seed=None, **kwargs):
        if activation != 'relu':
            raise NotImplementedError('only the default relu activation is implemented on the B200 path')
        if not should_expand_input_dim or len(keras_input_layer.shape) != 1:
            raise NotImplementedError('SimpleConvNetAutoregressive1D needs a 1-D spin input')
        self.depth = depth
        self.num_of_channels = num_of_channels
        self.kernel_size = kernel_size
        self.use_dilation = use_dilation
        self.add_skip_connections = add_skip_connections
        self.max_dilation_rate = max_dilation_rate
        self.activation = activation
        self.weights_normalization = weights_normalization
        self.exponential_norm = False   # WeightNormalization default in this machine (linear g)
        self.should_expand_input_dim = should_expand_input_dim
        self._seed = seed
        super(SimpleConvNetAutoregressive1D, self).__init__(keras_input_layer, **kwargs)

    def weight_specs(self):
        specs = []
        wn = self.weights_normalization
        shapes = [(self.kernel_size, 1 if i == 0 else self.num_of_channels, self.num_of_channels)
                  for i in range(self.depth - 2)] + [(1, self.num_of_channels, 4)]
        for idx, shape in enumerate(shapes):
            base = 'weight_normalization' if wn else 'conv1d'
            name = base if idx == 0 else '%s_%d' % (base, idx)
            specs.append((name + '/kernel:0', shape, 'glorot_uniform'))
            specs.append((name + '/bias:0', (shape[-1],), 'zeros'))
            if wn:
                specs.append((name + '/g:0', (shape[-1],), 'wn_g:%d' % (len(specs) - 2)))
        return specs

    def _create_args(self):
        (N,) = self.keras_input_layer.shape
        flags = 0
        if self.weights_normalization:
            flags |= _lib.FK_FLAG_WEIGHT_NORM
        if self.add_skip_connections:
            flags |= _lib.FK_FLAG_SKIP
        max_dil = self.max_dilation_rate if (self.use_dilation and self.max_dilation_rate) else 0
        return (_lib.FK_NET_CONV1D, 1, N, self.depth, self.num_of_channels, self.kernel_size, max_dil, flags)
