"""Shared helpers for the parity tests: build the same machine in the product (flowket_b200) and in the
oracle (oracle.nets) with identical weights."""
import numpy as np
import torch

from oracle import nets


def make_pair(kind, shape, depth, channels, seed=0, bias_scale=0.2, dtype=torch.float64, **kw):
    """-> (product Model for predictions, product Model for conditional log probs, oracle spec, oracle params)"""
    from flowket_b200 import Input, Model
    from flowket_b200.machines import ConvNetAutoregressive2D, SimpleConvNetAutoregressive1D, \
        ComplexValuesSimpleConvNetAutoregressive1D
    inp = Input(shape=shape, dtype='int8')
    if kind == 'conv2d':
        wn = kw.get('weights_normalization', True)
        machine = ConvNetAutoregressive2D(inp, depth=depth, num_of_channels=channels, weights_normalization=wn, seed=seed)
        spec = nets.Conv2DSpec(shape[0], shape[1], depth, channels, weights_normalization=wn)
    elif kind == 'conv1d':
        wn = kw.get('weights_normalization', True)
        machine = SimpleConvNetAutoregressive1D(inp, depth=depth, num_of_channels=channels, weights_normalization=wn,
                                                max_dilation_rate=kw.get('max_dilation_rate'),
                                                add_skip_connections=kw.get('add_skip_connections', False), seed=seed)
        spec = nets.Conv1DSpec(shape[0], depth, channels, weights_normalization=wn,
                               max_dilation_rate=kw.get('max_dilation_rate'),
                               add_skip_connections=kw.get('add_skip_connections', False))
    else:
        machine = ComplexValuesSimpleConvNetAutoregressive1D(inp, depth=depth, num_of_channels=channels,
                                                             max_dilation_rate=kw.get('max_dilation_rate'), seed=seed)
        spec = nets.ComplexConv1DSpec(shape[0], depth, channels, max_dilation_rate=kw.get('max_dilation_rate'))
    rng = np.random.RandomState(seed + 1000)
    weights = machine.initial_weights(seed)
    # exercise biases and weight-norm gains: perturb everything that is not a kernel
    for i, (name, shp, init) in enumerate(machine.weight_specs()):
        if 'bias' in name:
            weights[i] = (rng.normal(size=shp) * bias_scale).astype(np.float32)
        elif name.endswith('/g:0'):
            weights[i] = (weights[i] + rng.normal(size=shp) * 0.1).astype(np.float32)
    machine.set_weights(weights)
    params = [torch.from_numpy(w.astype(np.float64)).to(dtype) for w in weights]
    assert [tuple(p.shape) for p in params] == [tuple(p.shape) for p in nets.init_params(spec)]
    model = Model(inputs=inp, outputs=machine.predictions)
    cond_model = Model(inputs=inp, outputs=machine.conditional_log_probs)
    return model, cond_model, spec, params


def random_sigma(n, shape, seed=0):
    return np.random.RandomState(seed).choice([-1, 1], size=(n,) + tuple(shape)).astype(np.int8)
