"""CPU restatement (torch-CPU, fp32 or fp64) of the reference's autoregressive machines.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Restates (file:line relative to /root/reference/src/flowket):
  * ConvNetAutoregressive2D            machines/conv_net_autoregressive_2D.py:24-74
  * SimpleConvNetAutoregressive1D      machines/simple_conv_net_autoregressive_1D.py:8-64
  * ComplexValuesSimpleConvNetAR1D     machines/complex_values_simple_conv_net_autoregressive_1D.py:11-59
  * DownShift / RightShift             deepar/layers/masking.py:8-19
  * WeightNormalization                deepar/layers/wrappers.py:26-39,123-134
  * VectorToComplexNumber              layers/complex/casting.py:16-24
  * normalize_in_log_space(norm 2)     deepar/layers/autoregressive.py:7-15
  * one-hot + combine conditionals     deepar/layers/one_hot.py:7-9, deepar/layers/autoregressive.py:18-22
  * conditional_log_probs = 2 Re       machines/abstract_machine.py:56-57
  * complex conv (4 real convs)        layers/complex/tensorflow_ops.py:7-13, layers/complex/conv.py:81-107
  * lncosh                             layers/complex/tensorflow_ops.py:79-85
  * conj-stored complex weights        layers/complex/base_layer.py:18-35

Keras semantics restated: NHWC activations, HWIO kernels, 'valid' convolution after
explicit zero padding, bias on every conv, glorot_uniform / zeros default init.

Parameters are a flat python list of torch tensors in *layer creation order*
(the order Keras numbers `weight_normalization_k` / `conv2d_k`):
  2-D net : for each block b: [v(kxk), x(1xk), xx(1x1), y(1x1), h(kxk)] each as
            (kernel HWIO, bias[, g]); then head (kernel[1,1,C,4], bias[4]).
  1-D net : for each hidden conv (kernel [k,cin,C], bias[, g]); head (kernel[1,C,4], bias[, g]).
  complex : for each conv (kernel_real, kernel_imag, bias_real, bias_imag), W = real - i*imag.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------
# specs
# ----------------------------------------------------------------------------------------
class Conv2DSpec(object):
    kind = 'conv2d'

    def __init__(self, H, W, depth, channels, kernel_size=3, weights_normalization=True,
                 exponential_norm=True):
        assert depth >= 2 and kernel_size % 2 == 1
        self.H, self.W, self.depth, self.C, self.k = H, W, depth, channels, kernel_size
        self.wn, self.exp_norm = weights_normalization, exponential_norm
        self.num_blocks = 2 * depth - 2

    @property
    def input_shape(self):
        return (self.H, self.W)

    def conv_shapes(self):
        """[(kh, kw, cin, cout, weight_normalised)] in creation order."""
        k, C = self.k, self.C
        res = []
        for b in range(self.num_blocks):
            cin = 1 if b == 0 else C
            res += [(k, k, cin, C, self.wn), (1, k, cin, C, self.wn), (1, 1, C, C // 2, self.wn),
                    (1, 1, C, C // 2, self.wn), (k, k, C, C, self.wn)]
        res.append((1, 1, C, 4, False))  # head is never weight-normalised (conv_net_autoregressive_2D.py:73)
        return res


class Conv1DSpec(object):
    kind = 'conv1d'

    def __init__(self, N, depth, channels, kernel_size=3, use_dilation=True, add_skip_connections=False,
                 max_dilation_rate=None, weights_normalization=True):
        assert depth >= 3
        self.N, self.depth, self.C, self.k = N, depth, channels, kernel_size
        self.use_dilation, self.skip, self.max_dil = use_dilation, add_skip_connections, max_dilation_rate
        self.wn, self.exp_norm = weights_normalization, False  # 1-D net: linear g (wrappers.py default)

    @property
    def input_shape(self):
        return (self.N,)

    def dilations(self, n_layers=None):
        """simple_conv_net_autoregressive_1D.py:49-60: dilation doubles only if max_dilation_rate is set."""
        n_layers = self.depth - 2 if n_layers is None else n_layers
        d, res = 1, []
        for _ in range(n_layers):
            res.append(d)
            if self.use_dilation and self.max_dil is not None and d < self.max_dil:
                d *= 2
        return res

    def conv_shapes(self):
        res = []
        for i in range(self.depth - 2):
            res.append((1, self.k, 1 if i == 0 else self.C, self.C, self.wn))
        res.append((1, 1, self.C, 4, self.wn))  # head goes through causal_conv_1d -> WN applies
        return res


class ComplexConv1DSpec(object):
    kind = 'cconv1d'

    def __init__(self, N, depth, channels, kernel_size=3, use_dilation=True, max_dilation_rate=None):
        assert depth >= 2
        self.N, self.depth, self.C, self.k = N, depth, channels, kernel_size
        self.use_dilation, self.max_dil = use_dilation, max_dilation_rate

    @property
    def input_shape(self):
        return (self.N,)

    def dilations(self):
        d, res = 1, []
        for _ in range(self.depth - 1):
            res.append(d)
            if self.use_dilation and self.max_dil is not None and d < self.max_dil:
                d *= 2
        return res

    def conv_shapes(self):
        res = []
        for i in range(self.depth - 1):
            res.append((1, self.k, 1 if i == 0 else self.C, self.C))
        res.append((1, 1, self.C, 2))
        return res


# ----------------------------------------------------------------------------------------
# initialisation (Keras defaults restated; synthetic random-init weights for benchmarks)
# ----------------------------------------------------------------------------------------
def init_params(spec, seed=0, dtype=torch.float32, bias_scale=0.0):
    """glorot_uniform kernels (limit sqrt(6/(fan_in+fan_out))), zero biases, WN g from CopyNormaInitializer.

    `bias_scale` > 0 draws non-zero biases (tests use it so bias handling is exercised)."""
    gen = torch.Generator().manual_seed(seed)
    params = []
    if spec.kind in ('conv2d', 'conv1d'):
        for (kh, kw, cin, cout, wn) in spec.conv_shapes():
            fan_in, fan_out = kh * kw * cin, kh * kw * cout
            limit = math.sqrt(6.0 / (fan_in + fan_out))
            shape = (kh, kw, cin, cout) if spec.kind == 'conv2d' else (kw, cin, cout)
            kernel = (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * limit
            bias = torch.randn((cout,), generator=gen, dtype=torch.float64) * bias_scale
            params += [kernel.to(dtype), bias.to(dtype)]
            if wn:
                norm = kernel.reshape(-1, cout).norm(dim=0)
                g = torch.log(norm + 1e-10) if spec.exp_norm else norm
                params.append(g.to(dtype))
    else:
        for (kh, kw, cin, cout) in spec.conv_shapes():
            fan_in, fan_out = kw * cin, kw * cout
            std = math.sqrt(2.0 / (fan_in + fan_out))
            kr = torch.randn((kw, cin, cout), generator=gen, dtype=torch.float64) * std
            ki = -torch.randn((kw, cin, cout), generator=gen, dtype=torch.float64) * std  # ConjugateDecorator
            br = torch.randn((cout,), generator=gen, dtype=torch.float64) * bias_scale
            bi = torch.randn((cout,), generator=gen, dtype=torch.float64) * bias_scale
            params += [kr.to(dtype), ki.to(dtype), br.to(dtype), bi.to(dtype)]
    return params


def num_params(spec):
    return sum(int(p.numel()) for p in init_params(spec))


def flatten_params(params):
    return torch.cat([p.reshape(-1) for p in params])


def unflatten_params(spec, flat, dtype=None):
    res, off = [], 0
    for p in init_params(spec):
        n = p.numel()
        t = flat[off:off + n].reshape(p.shape)
        res.append(t if dtype is None else t.to(dtype))
        off += n
    assert off == flat.numel()
    return res


# ----------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------
def _effective_kernel(kernel, g, exp_norm):
    """WeightNormalization._compute_weights: l2_normalize(v, axes != out) * (exp(g) | g)."""
    if g is None:
        return kernel
    cout = kernel.shape[-1]
    sq = (kernel.reshape(-1, cout) ** 2).sum(dim=0)
    inv = torch.rsqrt(torch.clamp(sq, min=1e-12))
    scale = torch.exp(g) if exp_norm else g
    return kernel * (inv * scale)


def _conv2d_nhwc(x, kernel_hwio, bias, pad, dilation=(1, 1)):
    """x [n,H,W,C]; pad = (top, bottom, left, right); 'valid' conv after explicit zero padding."""
    t, b, l, r = pad
    xn = F.pad(x.permute(0, 3, 1, 2), (l, r, t, b))
    w = kernel_hwio.permute(3, 2, 0, 1)
    y = F.conv2d(xn, w, bias, dilation=dilation)
    return y.permute(0, 2, 3, 1)


def _down_shift(x):   # axis 1 (rows / sequence), masking.py:8-19
    return torch.cat([torch.zeros_like(x[:, :1]), x[:, :-1]], dim=1)


def _right_shift(x):  # axis 2 (columns)
    return torch.cat([torch.zeros_like(x[:, :, :1]), x[:, :, :-1]], dim=2)


class _ParamReader(object):
    def __init__(self, params):
        self.params, self.i = params, 0

    def take(self, n):
        res = self.params[self.i:self.i + n]
        self.i += n
        return res


# ----------------------------------------------------------------------------------------
# ConvNetAutoregressive2D
# ----------------------------------------------------------------------------------------
def _conv2d_net_logits(spec, params, sigma):
    k, C = spec.k, spec.C
    pad = k - 1
    reader = _ParamReader(params)
    dtype = params[0].dtype

    def conv(x, kh, kw, padding):
        if spec.wn:
            kernel, bias, g = reader.take(3)
        else:
            (kernel, bias), g = reader.take(2), None
        return _conv2d_nhwc(x, _effective_kernel(kernel, g, spec.exp_norm), bias, padding)

    def block(v, h, mask=False):
        vp = conv(v, k, k, (pad, 0, pad // 2, pad // 2))
        x = torch.relu(conv(h, 1, k, (0, 0, pad, 0)))
        if mask:
            x = _right_shift(x)
        x = conv(x, 1, 1, (0, 0, 0, 0))
        y = conv(_down_shift(torch.relu(vp)), 1, 1, (0, 0, 0, 0))
        c = torch.cat([torch.relu(x), torch.relu(y)], dim=-1)
        hp = conv(c, k, k, (pad, 0, pad, 0))
        return vp, hp

    x = torch.as_tensor(sigma).to(dtype).unsqueeze(-1)
    v, h = block(x, x)
    v, h = torch.relu(v), torch.relu(h)
    for _ in range(spec.depth - 2):
        v_in, h_in = v, h
        v, h = block(v, h)
        v, h = torch.relu(v), torch.relu(h)
        v, h = block(v, h)
        v, h = torch.relu(v_in + v), torch.relu(h_in + h)
    _, x = block(v, h, mask=True)
    kernel, bias = reader.take(2)
    out = _conv2d_nhwc(torch.relu(x), kernel, bias, (0, 0, 0, 0))
    assert reader.i == len(params)
    return out  # [n,H,W,4]: channels 0,1 = Re (class +1, -1); 2,3 = Im


# ----------------------------------------------------------------------------------------
# SimpleConvNetAutoregressive1D
# ----------------------------------------------------------------------------------------
def _conv1d_net_logits(spec, params, sigma):
    reader = _ParamReader(params)
    dtype = params[0].dtype
    x = torch.as_tensor(sigma).to(dtype).unsqueeze(-1).unsqueeze(1)  # [n,1,N,1]

    def conv(x, dil):
        if spec.wn:
            kernel, bias, g = reader.take(3)
        else:
            (kernel, bias), g = reader.take(2), None
        kw = kernel.shape[0]
        padding = (kw - 1) * dil
        return _conv2d_nhwc(x, _effective_kernel(kernel, g, spec.exp_norm).unsqueeze(0), bias,
                            (0, 0, padding, 0), dilation=(1, dil))

    for i, dil in enumerate(spec.dilations()):
        skip = x if (spec.skip and i > 0) else None
        x = conv(x, dil)
        if skip is not None:
            x = x + skip
        x = torch.relu(x)
    x = _right_shift(x)  # DownShift along the sequence axis (= axis 2 in this [n,1,N,C] view)
    out = conv(x, 1)
    assert reader.i == len(params)
    return out[:, 0]  # [n,N,4]


# ----------------------------------------------------------------------------------------
# ComplexValuesSimpleConvNetAutoregressive1D
# ----------------------------------------------------------------------------------------
def lncosh(z):
    """layers/complex/tensorflow_ops.py:79-85 (log|.| + i*atan2 form of the complex log)."""
    a = z.real.abs()
    s = torch.exp(z - a) + torch.exp(-z - a)
    log_s = torch.complex(torch.log(s.abs()), torch.atan2(s.imag, s.real))
    return a - math.log(2.0) + log_s


def _cconv1d_net_cond(spec, params, sigma):
    reader = _ParamReader(params)
    rdtype = params[0].dtype
    cdtype = torch.complex64 if rdtype == torch.float32 else torch.complex128
    x = torch.as_tensor(sigma).to(rdtype).to(cdtype).unsqueeze(-1).unsqueeze(1)  # [n,1,N,1]

    def cconv(x, dil):
        kr, ki, br, bi = reader.take(4)
        wr, wi = kr.unsqueeze(0), -ki.unsqueeze(0)  # W = real - i*imag (base_layer.py:30-35)
        kw = kr.shape[0]
        padding = ((kw - 1) * dil)
        p = (0, 0, padding, 0)
        ac = _conv2d_nhwc(x.real, wr, None, p, (1, dil))
        bd = _conv2d_nhwc(x.imag, wi, None, p, (1, dil))
        ad = _conv2d_nhwc(x.real, wi, None, p, (1, dil))
        bc = _conv2d_nhwc(x.imag, wr, None, p, (1, dil))
        return torch.complex(ac - bd + br, ad + bc - bi)  # bias = br - i*bi

    for dil in spec.dilations():
        x = lncosh(cconv(x, dil))
    x = _right_shift(x)
    out = cconv(x, 1)
    assert reader.i == len(params)
    return out[:, 0]  # [n,N,2] complex


# ----------------------------------------------------------------------------------------
# public API
# ----------------------------------------------------------------------------------------
def unnormalized_conditional_log_wave_function(spec, params, sigma):
    """-> complex tensor [n, *input_shape, 2] (class 0 <-> sigma=+1, class 1 <-> sigma=-1)."""
    if spec.kind == 'conv2d':
        o = _conv2d_net_logits(spec, params, sigma)
        return torch.complex(o[..., 0:2], o[..., 2:4])
    if spec.kind == 'conv1d':
        o = _conv1d_net_logits(spec, params, sigma)
        return torch.complex(o[..., 0:2], o[..., 2:4])
    return _cconv1d_net_cond(spec, params, sigma)


def conditional_log_wave_function(spec, params, sigma):
    x = unnormalized_conditional_log_wave_function(spec, params, sigma)
    norm = 0.5 * torch.logsumexp(2.0 * x.real, dim=-1, keepdim=True)
    return torch.complex(x.real - norm, x.imag)


def conditional_log_probs(spec, params, sigma):
    return 2.0 * conditional_log_wave_function(spec, params, sigma).real


def log_psi(spec, params, sigma):
    """-> complex tensor [n]: sum over sites of cond_log_wf[site, (1 - sigma)//2]."""
    cond = conditional_log_wave_function(spec, params, sigma)
    s = torch.as_tensor(sigma)
    idx = ((1 - s.to(torch.int64)) // 2).unsqueeze(-1)
    sel_re = torch.gather(cond.real, -1, idx).squeeze(-1)
    sel_im = torch.gather(cond.imag, -1, idx).squeeze(-1)
    dims = tuple(range(1, s.dim()))
    return torch.complex(sel_re.sum(dim=dims), sel_im.sum(dim=dims))


def log_psi_numpy(spec, params, sigma, batch_size=None):
    """`model.predict`-shaped helper: ndarray[n,1] complex (complex64 for fp32 params)."""
    sigma = np.asarray(sigma)
    n = sigma.shape[0]
    batch_size = n if batch_size is None else batch_size
    out = []
    with torch.no_grad():
        for i in range(0, n, batch_size):
            out.append(log_psi(spec, params, sigma[i:i + batch_size]).numpy())
    return np.concatenate(out)[:, None] if out else np.zeros((0, 1), np.complex64)


def weighted_gradient(spec, params, sigma, y):
    """Gradient of sum_b 2 Re(log psi(sigma_b) * y_b) w.r.t. every parameter tensor (flat vector).

    This is `loss_for_energy_minimization` (optimization/loss.py:4-5) *summed* over the batch;
    Keras' additional 1/mini_batch mean is applied by the caller."""
    ps = [p.detach().clone().requires_grad_(True) for p in params]
    lp = log_psi(spec, ps, sigma)
    y = torch.as_tensor(y)
    loss = 2.0 * (lp.real * y.real.to(lp.real.dtype) - lp.imag * y.imag.to(lp.real.dtype)).sum()
    grads = torch.autograd.grad(loss, ps)
    return torch.cat([g.reshape(-1) for g in grads])


def per_sample_gradients(spec, params, sigma, part='real'):
    """O_b = d Re(log psi(sigma_b)) / d theta, [n, P] (Machine.predictions_jacobian,
    machines/abstract_machine.py:24-28). part='imag' gives d Im(log psi)/d theta."""
    rows = []
    sigma = np.asarray(sigma)
    for b in range(sigma.shape[0]):
        ps = [p.detach().clone().requires_grad_(True) for p in params]
        lp = log_psi(spec, ps, sigma[b:b + 1])
        target = lp.real.sum() if part == 'real' else lp.imag.sum()
        grads = torch.autograd.grad(target, ps)
        rows.append(torch.cat([g.reshape(-1) for g in grads]))
    return torch.stack(rows)
