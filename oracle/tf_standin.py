"""Eager stand-in for the slice of `tensorflow` / `tensorflow.keras` that the reference's MACHINE-BUILDING code touches.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Used by oracle/make_golden.py (`machines`) in the build container to run
the reference's own `ConvNetAutoregressive2D`, `SimpleConvNetAutoregressive1D` and `ComplexValuesSimpleConvNetAutoregressive1D`
classes (flowket/machines/*.py) and their custom layers (flowket/deepar/layers/*.py, flowket/layers/complex/*.py) *unmodified*:
which layers exist, in which order, with which paddings, shifts, dilations, skip connections, weight-normalisation formula,
complex-convolution composition, activation, head, log-space normalisation and one-hot combination is all decided by the
reference's code.  What this file supplies is only the meaning of the primitive operations, each the textbook one:

  * tensors are torch CPU tensors and every op runs immediately (a Keras "input layer" is simply a batch of spins);
  * `Conv1D/Conv2D` = cross-correlation, channels-last, 'valid', HWIO kernels, bias added (Keras semantics);
  * `ZeroPadding`, `Add`, `Concatenate`, `Activation('relu')`, `Lambda` = what their names say;
  * `Layer.add_weight` hands out the next array of an injected weight list (creation order), so a run is a pure function of
    (weights, spins); initialisers are ignored;
  * float32 / complex64 requests are served in float64 / complex128 so that the golden values carry no rounding of their own.

  * while a machine is built with RECORDING on, every layer notes which layers produced its inputs and the shapes involved
    (Keras' `inbound_nodes` / `input_shape` / `output_shape`), which is all the reference's graph analysis
    (deepar/graph_analysis/*.py) and FastAutoregressiveSampler read; `K.placeholder` is the concrete batch size and
    `K.function` returns what was already computed;
  * `tf.multinomial` -- the fast sampler's only source of randomness, unseeded in the reference -- is inverse-CDF sampling
    with injected uniforms: category 0 iff p0 > u (the explicit-uniform rule of deepar/samplers/autoregressive.py:37-44).

Nothing here knows anything about autoregressive machines."""
import contextlib
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

_WEIGHTS = []          # injected weights, consumed by add_weight in creation order
_CREATED = []          # (name, shape) of every weight handed out


def inject_weights(arrays, requires_grad=False):
    """`requires_grad`: the handed-out weights are autograd leaves, so the reference's forward (all torch ops here) can be
    differentiated with respect to them -- returns the leaves in injection order"""
    del _WEIGHTS[:]
    del _CREATED[:]
    leaves = [torch.as_tensor(np.asarray(a, dtype=np.float64)).clone().requires_grad_(requires_grad) for a in arrays]
    _WEIGHTS.extend(leaves)
    return leaves


def created_weights():
    return list(_CREATED)


# ---- dtypes ---------------------------------------------------------------------------------------------------------------
class DType(object):
    def __init__(self, name, torch_dtype):
        self.name, self.torch_dtype = name, torch_dtype
        self.is_complex = torch_dtype.is_complex

    def __eq__(self, other):
        return isinstance(other, DType) and other.name == self.name

    def __hash__(self):
        return hash(self.name)


float32, float64 = DType('float32', torch.float64), DType('float64', torch.float64)
complex64, complex128 = DType('complex64', torch.complex128), DType('complex128', torch.complex128)
int32, int64, int8 = DType('int32', torch.int64), DType('int64', torch.int64), DType('int8', torch.int64)
_BY_NAME = {d.name: d for d in (float32, float64, complex64, complex128, int32, int64, int8)}


def _torch_dtype(dtype):
    if dtype is None:
        return None
    if isinstance(dtype, torch.dtype):
        return dtype
    if isinstance(dtype, str):
        return _BY_NAME[dtype].torch_dtype
    return dtype.torch_dtype


torch.Tensor.get_shape = lambda self: self.shape           # TF 1.x spelling used by deepar/layers/masking.py


class Opaque(object):
    """a tensor hidden from numpy: the fast sampler fills object arrays with `numpy.full(shape, fill_value=zeros)`, which only
    keeps `zeros` as ONE object per cell if it is not array-like (a symbolic TF tensor is not)"""

    def __init__(self, tensor):
        self.tensor = tensor


def _t(x):
    if isinstance(x, Opaque):
        return x.tensor
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(x)


# ---- graph recording (for the reference's graph analysis: which layer feeds which) ---------------------------------------
RECORDING = [False]
_COUNTERS = {}


class Node(object):
    def __init__(self, inbound_layers, tensor_indices):
        self.inbound_layers, self.tensor_indices = inbound_layers, tensor_indices


def _shape_of(x):
    return (None,) + tuple(_t(x).shape[1:])


# ---- tensorflow.* -----------------------------------------------------------------------------------------------------------
def cast(x, dtype, name=None):
    x, dt = _t(x), _torch_dtype(dtype)
    if x.dtype.is_complex and not dt.is_complex:
        x = x.real
    return x.to(dt)


def complex_(real, imag, name=None):
    return torch.complex(_t(real).to(torch.float64), _t(imag).to(torch.float64))


def reshape(x, shape, name=None):
    return _t(x).reshape(tuple(int(s) for s in shape))


def unstack(x, axis=0, name=None):
    return list(torch.unbind(_t(x), dim=axis))


def stack(xs, axis=0, name=None):
    return torch.stack([_t(x) for x in xs], dim=axis)


def concat(xs, axis, name=None):
    return torch.cat([_t(x) for x in xs], dim=axis)


def slice_(x, begin, size, name=None):
    idx = tuple(slice(b, None if s == -1 else b + s) for b, s in zip(begin, size))
    return _t(x)[idx]


def zeros_like(x, dtype=None, name=None):
    return torch.zeros_like(_t(x), dtype=_torch_dtype(dtype) if not isinstance(dtype, torch.dtype) else dtype)


def one_hot(indices, depth, on_value=1.0, off_value=0.0, axis=-1, name=None):
    assert axis == -1
    hot = F.one_hot(_t(indices).to(torch.int64), depth).to(torch.float64)
    return hot * on_value + (1.0 - hot) * off_value


def _reduce(fn):
    def wrapped(x, axis=None, keepdims=False, name=None):
        x = _t(x)
        if axis is None:
            axis = list(range(x.dim()))
        return fn(x, dim=tuple(axis) if isinstance(axis, (list, tuple)) else axis, keepdim=keepdims)
    return wrapped


def _l2_normalize(x, axis=None, epsilon=1e-12, name=None):
    sq = (x * x).sum(dim=tuple(axis) if isinstance(axis, (list, tuple)) else axis, keepdim=True)
    return x * torch.rsqrt(torch.clamp(sq, min=epsilon))


def _bias_add(x, bias, name=None):
    return x + bias


def _conv(rank):
    def conv(x, kernel, strides=1, padding='valid', data_format=None, dilation_rate=1):
        assert padding == 'valid'
        strides = (strides,) * rank if isinstance(strides, int) else tuple(strides)
        dilation = (dilation_rate,) * rank if isinstance(dilation_rate, int) else tuple(dilation_rate)
        if rank == 1:
            return F.conv1d(x.permute(0, 2, 1), kernel.permute(2, 1, 0), stride=strides, dilation=dilation).permute(0, 2, 1)
        return F.conv2d(x.permute(0, 3, 1, 2), kernel.permute(3, 2, 0, 1), stride=strides, dilation=dilation).permute(0, 2, 3, 1)
    return conv


class TensorShape(object):
    def __init__(self, dims):
        self.dims = list(dims.dims if isinstance(dims, TensorShape) else dims)

    def as_list(self):
        return list(self.dims)


# ---- tensorflow.keras -----------------------------------------------------------------------------------------------------
class Layer(object):
    def __init__(self, name=None, dtype=None, **_ignored):
        if name is None:
            base = type(self).__name__.lower()
            _COUNTERS[base] = _COUNTERS.get(base, 0) + 1
            name = '%s_%d' % (base, _COUNTERS[base])
        self.name, self.built = name, False
        self.dtype = dtype
        self.input_spec = None
        self.inbound_nodes = []

    def add_weight(self, name=None, shape=None, initializer=None, dtype=None, trainable=True, **_ignored):
        assert _WEIGHTS, 'the injected weight list is exhausted at %r' % name
        w = _WEIGHTS.pop(0)
        assert tuple(w.shape) == tuple(int(s) for s in shape), (name, tuple(w.shape), tuple(shape))
        _CREATED.append((name, tuple(w.shape)))
        return w

    def build(self, input_shape=None):
        self.built = True

    def call(self, inputs, **kwargs):
        return inputs

    def compute_output_shape(self, input_shape):
        return input_shape

    def get_config(self):
        return {}

    def get_output_shape_at(self, index):
        return self.output_shape

    def __call__(self, inputs, **kwargs):
        many = isinstance(inputs, (list, tuple))
        inputs = [_t(i) for i in inputs] if many else _t(inputs)
        if not self.built:
            self.build([tuple(i.shape) for i in inputs] if many else tuple(inputs.shape))
            self.built = True
        if self.dtype is None:          # TF 1.x Keras: a layer without an explicit dtype takes the dtype of its first input
            self.dtype = (inputs[0] if many else inputs).dtype
        out = self.call(inputs, **kwargs)
        if RECORDING[0]:
            sources = inputs if many else [inputs]
            history = [getattr(i, '_keras_history', None) for i in sources]
            assert all(h is not None for h in history), 'input of %s was not produced by a layer' % self.name
            layers_in, indices = [h[0] for h in history], [h[1] for h in history]
            self.inbound_nodes.append(Node(layers_in if many else layers_in[0], indices if many else indices[0]))
            self.input = sources if many else inputs
            self.input_shape = [_shape_of(i) for i in sources] if many else _shape_of(inputs)
            self.output_shape = _shape_of(out)
            out._keras_history = (self, 0)
        return out


class InputLayer(Layer):
    """the layer behind a model input; `tensor` is the concrete batch the machine is built on"""

    def __init__(self, tensor, dtype='int8', name=None):
        super(InputLayer, self).__init__(name=name, dtype=dtype)
        self.output_shape = self.input_shape = _shape_of(tensor)
        self.input = tensor
        self.inbound_nodes = [Node([], [])]
        tensor._keras_history = (self, 0)


class GraphModel(object):
    """what the graph analysis reads from a Keras Model: the layers reachable from the output, their names, the shapes"""

    def __init__(self, inputs, outputs):
        self.input, self.output = inputs, outputs
        self.layers, seen, stack = [], set(), [outputs._keras_history[0]]
        while stack:
            layer = stack.pop()
            if id(layer) in seen:
                continue
            seen.add(id(layer))
            self.layers.append(layer)
            inbound = layer.inbound_nodes[-1].inbound_layers
            stack.extend(inbound if isinstance(inbound, list) else [inbound])
        self.input_names, self.output_names = [inputs._keras_history[0].name], [outputs._keras_history[0].name]
        self.input_shape, self.output_shape = _shape_of(inputs), _shape_of(outputs)

    def get_layer(self, name):
        return [layer for layer in self.layers if layer.name == name][0]


class Wrapper(Layer):
    def __init__(self, layer, **kwargs):
        super(Wrapper, self).__init__(**kwargs)
        self.layer = layer


class InputSpec(object):
    def __init__(self, **kwargs):
        self.__dict__.update(kwargs)


class Lambda(Layer):
    def __init__(self, function, **kwargs):
        super(Lambda, self).__init__(**kwargs)
        self.function = function

    def call(self, inputs, **kwargs):
        return self.function(inputs)


class Activation(Layer):
    def __init__(self, activation, **kwargs):
        super(Activation, self).__init__(**kwargs)
        self.activation = activation

    def call(self, inputs, **kwargs):
        if callable(self.activation):
            return self.activation(inputs)
        assert self.activation == 'relu', self.activation
        return torch.relu(inputs)


class Add(Layer):
    def call(self, inputs, **kwargs):
        out = inputs[0]
        for other in inputs[1:]:
            out = out + other
        return out


class Concatenate(Layer):
    def __init__(self, axis=-1, **kwargs):
        super(Concatenate, self).__init__(**kwargs)
        self.axis = axis

    def call(self, inputs, **kwargs):
        return torch.cat([_t(i) for i in inputs], dim=self.axis)


class ZeroPadding1D(Layer):
    def __init__(self, padding=1, **kwargs):
        super(ZeroPadding1D, self).__init__(**kwargs)
        self.padding = (padding, padding) if isinstance(padding, int) else tuple(padding)

    def call(self, x, **kwargs):                       # [n, steps, channels]
        return F.pad(x, (0, 0, self.padding[0], self.padding[1]))


class ZeroPadding2D(Layer):
    def __init__(self, padding=(1, 1), **kwargs):
        super(ZeroPadding2D, self).__init__(**kwargs)
        (self.top, self.bottom), (self.left, self.right) = [(p, p) if isinstance(p, int) else tuple(p) for p in padding]
        self.padding = ((self.top, self.bottom), (self.left, self.right))

    def call(self, x, **kwargs):                       # [n, rows, cols, channels]
        return F.pad(x, (0, 0, self.left, self.right, self.top, self.bottom))


class _Conv(Layer):
    rank = None

    def __init__(self, filters, kernel_size, strides=1, padding='valid', dilation_rate=1, use_bias=True, **kwargs):
        super(_Conv, self).__init__(**kwargs)
        assert padding == 'valid' and kwargs.get('activation') is None
        as_tuple = lambda v: (v,) * self.rank if isinstance(v, int) else tuple(v)     # noqa: E731
        self.filters, self.kernel_size = filters, as_tuple(kernel_size)
        self.strides, self.dilation_rate, self.use_bias = as_tuple(strides), as_tuple(dilation_rate), use_bias
        self.padding, self.activation = padding, None

    def build(self, input_shape):
        self.kernel = self.add_weight(name='kernel', shape=self.kernel_size + (int(input_shape[-1]), self.filters))
        self.bias = self.add_weight(name='bias', shape=(self.filters,)) if self.use_bias else None
        self.built = True

    def call(self, x, **kwargs):
        out = _conv(self.rank)(x, self.kernel, strides=self.strides, dilation_rate=self.dilation_rate)   # kernel read at call
        return out + self.bias if self.use_bias else out                                                   # time (weight norm)


class Conv1D(_Conv):
    rank = 1


class Conv2D(_Conv):
    rank = 2


class Initializer(object):
    pass


class _Unused(Layer):
    """layer types the reference registers topologies for but these machines never instantiate"""


# ---- sampling primitives of the fast sampler ------------------------------------------------------------------------------
BATCH = [0]              # value of the batch-size placeholder
UNIFORMS = []            # one [B] array of uniforms per multinomial call, consumed in call order
MULTINOMIAL_CALLS = [0]


def _multinomial(logits, num_samples, output_dtype=None, name=None):
    """category 0 iff exp(logits[:, 0]) / sum exp(logits) > u -- inverse-CDF sampling of a two-class distribution with the
    injected uniforms (the explicit-uniform rule of deepar/samplers/autoregressive.py:37-44)"""
    logits = _t(logits).to(torch.float64)
    assert num_samples == 1 and logits.shape[-1] == 2
    p0 = torch.softmax(logits, dim=-1)[:, 0]
    u = torch.as_tensor(UNIFORMS.pop(0)) if UNIFORMS else torch.full_like(p0, 0.5)
    MULTINOMIAL_CALLS[0] += 1
    return (~(p0 > u)).to(torch.int64).reshape(-1, 1)


class _Shape(list):
    def as_list(self):
        return list(self)


class Variable(torch.Tensor):
    """a weight as the optimizer code sees it: `.shape.as_list()` exists (TF 1.x), arithmetic is torch's"""

    @staticmethod
    def __new__(cls, data):
        return torch.Tensor._make_subclass(cls, torch.as_tensor(data))

    @property
    def shape(self):
        return _Shape(super(Variable, self).shape)


def _matmul(a, b, adjoint_a=False, transpose_a=False, name=None):
    if adjoint_a:
        a = a.conj().transpose(-1, -2)
    elif transpose_a:
        a = a.transpose(-1, -2)
    return a @ b


class Model(object):
    """a finished eager computation: `output` is already a tensor"""

    def __init__(self, inputs=None, outputs=None, **_ignored):
        self.input, self.output = inputs, outputs


def install():
    """register the stand-in under the module names the reference imports; returns the names so that the caller can
    remove them again"""
    def module(name, **attrs):
        mod = types.ModuleType(name)
        mod.__dict__.update(attrs)
        sys.modules[name] = mod
        return mod
    scope = lambda *a, **k: contextlib.nullcontext('scope')     # noqa: E731
    math = types.SimpleNamespace(
        real=lambda x, name=None: _t(x).real if _t(x).dtype.is_complex else _t(x), imag=lambda x, name=None: _t(x).imag,
        abs=lambda x, name=None: torch.abs(_t(x)), exp=lambda x, name=None: torch.exp(_t(x)),
        log=lambda x, name=None: torch.log(_t(x).to(torch.float64) if not _t(x).dtype.is_complex else _t(x)),
        atan2=lambda y, x, name=None: torch.atan2(_t(y), _t(x)), sqrt=lambda x, name=None: torch.sqrt(_t(x)),
        multiply=lambda x, y, name=None: _t(x) * y,
        reduce_sum=_reduce(torch.sum), reduce_mean=_reduce(torch.mean), reduce_logsumexp=_reduce(torch.logsumexp))
    nn = types.SimpleNamespace(relu=lambda x, name=None: torch.relu(x), l2_normalize=_l2_normalize, bias_add=_bias_add)
    linalg = types.SimpleNamespace(norm=lambda x, axis=None: torch.linalg.vector_norm(x, dim=axis),
                                   cholesky=torch.linalg.cholesky, solve=torch.linalg.solve,
                                   cholesky_solve=lambda chol, rhs: torch.cholesky_solve(rhs, chol))
    backend = module('tensorflow.keras.backend', int_shape=lambda x: tuple(x.shape), update=lambda ref, value: (ref, value),
                     placeholder=lambda dtype=None, shape=None: BATCH[0],
                     function=lambda inputs, outputs: (lambda feed: [_t(o) for o in outputs]),
                     name_scope=scope, variable=lambda value, **k: value, epsilon=lambda: 1e-7,
                     expand_dims=lambda x, axis=-1: _t(x).unsqueeze(axis), cast=cast, conv1d=_conv(1), conv2d=_conv(2),
                     set_floatx=lambda name: None)
    layers = module('tensorflow.keras.layers', Layer=Layer, Wrapper=Wrapper, InputSpec=InputSpec, Lambda=Lambda,
                    Activation=Activation, Add=Add, Concatenate=Concatenate, ZeroPadding1D=ZeroPadding1D,
                    ZeroPadding2D=ZeroPadding2D, Conv1D=Conv1D, Conv2D=Conv2D, Input=None, Dense=None,
                    **{n: type(n, (_Unused,), {}) for n in ('Conv3D', 'ZeroPadding3D', 'Average', 'Subtract', 'Multiply', 'Maximum',
                                                              'Minimum', 'LeakyReLU', 'ELU', 'ThresholdedReLU', 'Softmax')})
    # images are [n, rows, cols, channels]; rot90 turns counter-clockwise like numpy / torch
    image = types.SimpleNamespace(rot90=lambda x, k=1, name=None: torch.rot90(x, k, dims=(1, 2)),
                                  flip_left_right=lambda x: torch.flip(x, dims=(2,)))
    initializers = module('tensorflow.keras.initializers', Initializer=Initializer)
    models = module('tensorflow.keras.models', Model=Model)
    module('tensorflow.keras.optimizers', Optimizer=object)
    keras = module('tensorflow.keras', backend=backend, layers=layers, initializers=initializers, models=models)
    py_keras = module('tensorflow.python.keras', backend=backend, layers=layers)
    public = lambda mod: {k: v for k, v in mod.__dict__.items() if not k.startswith('__')}     # noqa: E731
    module('tensorflow.python.keras.layers', **public(layers))
    module('tensorflow.python.keras.backend', **public(backend))
    def jacobian(y, xs, use_pfor=False):
        """d y[b, ...] / d x for every x in xs -> list of [B, *x.shape] (tf.python.ops.parallel_for.gradients.jacobian);
        rows by reverse-mode autograd over the torch graph the stand-in built"""
        y = _t(y)
        flat = y.reshape(y.shape[0], -1)
        assert flat.shape[1] == 1, 'per-sample scalar outputs only'
        rows = [torch.autograd.grad(flat[b, 0], xs, retain_graph=True) for b in range(flat.shape[0])]
        return [torch.stack([rows[b][i] for b in range(len(rows))]) for i in range(len(xs))]
    pfor_gradients = types.SimpleNamespace(jacobian=jacobian)
    pfor = module('tensorflow.python.ops.parallel_for', gradients=pfor_gradients)
    module('tensorflow.python.ops.parallel_for.gradients', jacobian=jacobian)
    ops = module('tensorflow.python.ops', parallel_for=pfor)
    python = module('tensorflow.python', keras=py_keras, ops=ops)
    engine = module('tensorflow.python.keras.engine')
    module('tensorflow.python.keras.engine.input_layer', InputLayer=InputLayer)
    py_keras.engine = engine
    module('tensorflow', math=math, nn=nn, linalg=linalg, keras=keras, python=python, complex=complex_, cast=cast,
           image=image, expand_dims=lambda x, axis=-1, name=None: _t(x).unsqueeze(axis),
           shape=lambda x, name=None: list(_t(x).shape), matmul=_matmul, conj=lambda x, name=None: torch.conj(_t(x)).resolve_conj(),
           eye=lambda n, dtype=None, name=None: torch.eye(int(n), dtype=_torch_dtype(dtype)), squeeze=lambda x, axis=None: _t(x).squeeze(),
           control_dependencies=lambda deps: contextlib.nullcontext(), stop_gradient=lambda x, name=None: x, no_op=lambda: None,
           __version__='1.13.1', zeros=lambda shape, dtype=None, name=None: Opaque(torch.zeros(tuple(int(v) for v in shape), dtype=_torch_dtype(dtype))),
           as_dtype=lambda d: d, multinomial=_multinomial, int8=int8,
           roll=lambda x, shift, axis, name=None: torch.roll(x, shifts=tuple(shift), dims=tuple(axis)),
           reshape=reshape, unstack=unstack, stack=stack, concat=concat, slice=slice_, zeros_like=zeros_like,
           one_hot=one_hot, name_scope=scope, TensorShape=TensorShape, float32=float32, float64=float64,
           complex64=complex64, complex128=complex128, int32=int32, int64=int64)
    return [k for k in sys.modules if k == 'tensorflow' or k.startswith('tensorflow.')]
