"""Host-side mirror of flowket/machines/abstract_machine.py:13-71: the same class hierarchy and properties,
with the Keras graph replaced by a handle to the CUDA layer program (fk_net_create)."""
import abc
import math

import numpy as np

from .. import _lib


class SymbolicOutput(object):
    """What `machine.predictions` / `.conditional_log_probs` return: a tag that Model(inputs, outputs) binds."""

    def __init__(self, machine, kind):
        self.machine, self.kind = machine, kind


class Machine(abc.ABC):
    def __init__(self, keras_input_layer, use_pfor=False):
        self.keras_input_layer = keras_input_layer
        self.use_pfor = use_pfor

    @property
    @abc.abstractmethod
    def predictions(self):
        pass

    def predictions_jacobian(self, params=None):
        """Machine.predictions_jacobian (abstract_machine.py:24-28): returns a callable
        sigma -> (d Re log psi / d theta [B,P], d Im log psi / d theta [B,P]) on the device."""
        return lambda sigma: self.device_net().grad_per_sample(self.device_net().to_sigma(sigma), imag=True)


class AutoregressiveMachine(Machine):
    @property
    def predictions(self):
        return SymbolicOutput(self, 'predictions')

    @property
    @abc.abstractmethod
    def conditional_log_wave_function(self):
        pass


class AutoNormalizedAutoregressiveMachine(AutoregressiveMachine):
    """Shared implementation of the three autoregressive conv machines."""

    net_kind = None

    def __init__(self, keras_input_layer, equal_up_down_spins=False, **kwargs):
        if equal_up_down_spins:
            raise NotImplementedError('equal_up_down_spins is outside the B200 hot path (SURVEY.md appendix A-11)')
        super(AutoNormalizedAutoregressiveMachine, self).__init__(keras_input_layer, **kwargs)
        self._device_net = None
        self._params = None
        self._dirty = True

    # ---- symbolic outputs (same property names as the reference) ------------------------------------
    @property
    def conditional_log_wave_function(self):
        return SymbolicOutput(self, 'conditional_log_wave_function')

    @property
    def conditional_log_probs(self):
        return SymbolicOutput(self, 'conditional_log_probs')

    @property
    def unnormalized_conditional_log_wave_function(self):
        return SymbolicOutput(self, 'unnormalized_conditional_log_wave_function')

    # ---- parameter layout ----------------------------------------------------------------------------
    @abc.abstractmethod
    def weight_specs(self):
        """[(name, shape, init)] in Keras layer-creation order; init in {'glorot_uniform','zeros','wn_g:<i>',
        'glorot_normal','neg_glorot_normal'}"""

    @abc.abstractmethod
    def _create_args(self):
        """(kind, H, W, depth, channels, kernel_size, max_dilation, flags)"""

    @property
    def input_shape(self):
        return tuple(self.keras_input_layer.shape)

    def count_params(self):
        return int(sum(int(np.prod(s)) for _, s, _ in self.weight_specs()))

    def initial_weights(self, seed=None):
        rng = np.random.RandomState(seed)
        specs = self.weight_specs()
        out = []
        for name, shape, init in specs:
            if init == 'zeros':
                w = np.zeros(shape, np.float32)
            elif init in ('glorot_uniform', 'glorot_normal', 'neg_glorot_normal'):
                receptive = int(np.prod(shape[:-2]))
                fan_in, fan_out = receptive * shape[-2], receptive * shape[-1]
                if init == 'glorot_uniform':
                    limit = math.sqrt(6.0 / (fan_in + fan_out))
                    w = rng.uniform(-limit, limit, size=shape).astype(np.float32)
                else:
                    w = rng.normal(0.0, math.sqrt(2.0 / (fan_in + fan_out)), size=shape).astype(np.float32)
                    if init.startswith('neg'):
                        w = -w   # ConjugateDecorator (layers/complex/initializers.py:36-47)
            elif init.startswith('wn_g:'):
                # CopyNormaInitializer (deepar/layers/wrappers.py:26-39): g = |v| or log(|v| + 1e-10)
                v = out[int(init.split(':')[1])]
                norm = np.sqrt((v.reshape(-1, v.shape[-1]).astype(np.float64) ** 2).sum(axis=0))
                w = (np.log(norm + 1e-10) if self.exponential_norm else norm).astype(np.float32)
            else:
                raise ValueError(init)
            out.append(w)
        return out

    # ---- weights -------------------------------------------------------------------------------------
    def get_weights(self):
        flat = self.flat_params_numpy()
        res, off = [], 0
        for _, shape, _ in self.weight_specs():
            n = int(np.prod(shape))
            res.append(flat[off:off + n].reshape(shape).copy())
            off += n
        return res

    def set_weights(self, weights):
        specs = self.weight_specs()
        if len(weights) != len(specs):
            raise ValueError('expected %d weight arrays, got %d' % (len(specs), len(weights)))
        for w, (name, shape, _) in zip(weights, specs):
            if tuple(np.shape(w)) != tuple(shape):
                raise ValueError('weight %s: expected shape %s, got %s' % (name, shape, np.shape(w)))
        self.set_flat_params(np.concatenate([np.asarray(w, np.float32).reshape(-1) for w in weights]))

    def flat_params_numpy(self):
        if self._params is None:
            self.set_weights(self.initial_weights(self._seed))
        p = self._params
        return p.detach().cpu().numpy().copy() if hasattr(p, 'detach') else np.array(p, copy=True)

    def set_flat_params(self, flat):
        """flat: numpy array or torch tensor of `count_params()` fp32 values."""
        try:
            import torch
        except ImportError:  # pragma: no cover
            torch = None
        if torch is not None and isinstance(flat, torch.Tensor):
            self._params = flat.detach().to(torch.float32).reshape(-1).clone()
        else:
            flat = np.asarray(flat, np.float32).reshape(-1)
            assert flat.size == self.count_params()
            self._params = torch.from_numpy(flat.copy()) if torch is not None else flat.copy()
        self._dirty = True

    def flat_params_device(self):
        """The live device copy (updated in place by the optimisers; call `params_updated()` afterwards)."""
        net = self.device_net()
        return self._params

    def params_updated(self):
        self._dirty = True

    # ---- device ---------------------------------------------------------------------------------------
    def device_net(self):
        """Creates the CUDA layer program on first use and keeps the device weights in sync."""
        if self._device_net is None:
            from .._device import DeviceNet
            self._device_net = DeviceNet(*self._create_args())
            assert self._device_net.num_params == self.count_params(), \
                (self._device_net.num_params, self.count_params())
        net = self._device_net
        if self._params is None:
            self.set_weights(self.initial_weights(self._seed))
        if not self._params.is_cuda:
            self._params = self._params.to(net.device)
        if self._dirty:
            net.set_params(self._params)
            self._dirty = False
        return net

    # ---- evaluation used by Model.predict ----------------------------------------------------------------
    def evaluate(self, kind, x, batch_size=None, engine=_lib.FK_ENGINE_FP32):
        net = self.device_net()
        sigma = net.to_sigma(x)
        n = sigma.shape[0]
        if kind == 'predictions':
            return net.log_psi(sigma, engine=engine, max_chunk=batch_size).reshape(n, 1)
        if kind == 'conditional_log_probs':
            return net.cond_log_probs(sigma).reshape((n,) + self.input_shape + (2,))
        raise NotImplementedError('Model output %r is not exposed by the B200 path' % kind)
