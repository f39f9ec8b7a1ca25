"""ConvNetAutoregressive2D: same constructor as flowket/machines/conv_net_autoregressive_2D.py:11-22
(+ abstract_machine.py:14,50 kwargs); the PixelCNN-style two-stack masked conv net is executed by
hand-written sm_100a kernels (flowket_b200/csrc)."""
from .abstract_machine import AutoNormalizedAutoregressiveMachine
from .. import _lib


class ConvNetAutoregressive2D(AutoNormalizedAutoregressiveMachine):
    def __init__(self, keras_input_layer, depth, num_of_channels, kernel_size=3, strides=1, activation='relu',
                 weights_normalization=True, exponential_norm=True, seed=None, **kwargs):
        if strides != 1:
            raise NotImplementedError('strides != 1')
        if activation != 'relu':
            raise NotImplementedError('only the default relu activation is implemented on the B200 path')
        if len(keras_input_layer.shape) != 2:
            raise ValueError('ConvNetAutoregressive2D needs a 2-D input, got shape %r' % (keras_input_layer.shape,))
        self.depth = depth
        self.num_of_channels = num_of_channels
        self.kernel_size = kernel_size
        self.padding = kernel_size - 1
        self.strides = strides
        self.activation = activation
        self.weights_normalization = weights_normalization
        self.exponential_norm = exponential_norm
        self._seed = seed
        super(ConvNetAutoregressive2D, self).__init__(keras_input_layer, **kwargs)

    def weight_specs(self):
        k, C = self.kernel_size, self.num_of_channels
        specs = []

        def conv(idx, kh, kw, cin, cout, wn):
            base = 'weight_normalization' if wn else 'conv2d'
            name = base if idx == 0 else '%s_%d' % (base, idx)
            specs.append((name + '/kernel:0', (kh, kw, cin, cout), 'glorot_uniform'))
            specs.append((name + '/bias:0', (cout,), 'zeros'))
            if wn:
                specs.append((name + '/g:0', (cout,), 'wn_g:%d' % (len(specs) - 2)))

        idx = 0
        for b in range(2 * self.depth - 2):
            cin = 1 if b == 0 else C
            for (kh, kw, ci, co) in [(k, k, cin, C), (1, k, cin, C), (1, 1, C, C // 2), (1, 1, C, C // 2), (k, k, C, C)]:
                conv(idx, kh, kw, ci, co, self.weights_normalization)
                idx += 1
        conv(idx, 1, 1, C, 4, False)   # the head is never weight-normalised (conv_net_autoregressive_2D.py:73)
        return specs

    def _create_args(self):
        H, W = self.keras_input_layer.shape
        flags = 0
        if self.weights_normalization:
            flags |= _lib.FK_FLAG_WEIGHT_NORM
            if self.exponential_norm:
                flags |= _lib.FK_FLAG_EXP_NORM
        return (_lib.FK_NET_CONV2D, H, W, self.depth, self.num_of_channels, self.kernel_size, 0, flags)
