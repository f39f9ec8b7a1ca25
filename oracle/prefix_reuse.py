"""Design oracle for the next step of the local-energy kernel: PREFIX REUSE.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Nothing here is on the product path; this module pins the
dependency analysis a row-trimmed CUDA forward will rely on, on the CPU, before any kernel is written.

A connected configuration sigma' of the Heisenberg / Ising / J1J2 operators differs from its sample sigma only at one or
two sites.  ConvNetAutoregressive2D (machines/conv_net_autoregressive_2D.py:33-74) is causal along rows: every convolution
pads the top only (3x3: two rows up; 1x3, 1x1: same row; DownShift: one row up, deepar/layers/masking.py:8-19), so with
r0 = first row in which sigma' differs from sigma

  * every activation of every layer in rows < r0 is identical for sigma and sigma';
  * rows >= r0 of a block can be recomputed from the *new* rows >= r0 plus a HALO of unchanged rows taken from the
    sample's own forward pass: rows r0-2, r0-1 of the block's vertical input and of its concat tensor, row r0-1 of
    relu(v');  nothing above row r0-2 is ever read;
  * log psi(sigma') = sum_{rows < r0} (the sample's conditional log-amplitudes, selected by sigma) +
                      sum_{rows >= r0} (recomputed conditionals, selected by sigma').

`forward_with_cache` runs the ordinary forward and keeps the three halo sources per block; `log_psi_from_row` recomputes
rows >= r0 touching only those halo rows.  tests/test_prefix_reuse_oracle.py checks it against the full forward
(oracle/nets.py) to fp64 round-off over all connections of random samples and reports the row-work that is saved."""
import numpy as np
import torch
import torch.nn.functional as F

from . import nets


def _kernels(spec, params):
    """effective (weight-normalised) kernels and biases in layer-creation order: per block V, X, XX, Y, H; then the head"""
    reader = nets._ParamReader(params)
    out = []
    nb = 2 * spec.depth - 2
    for _ in range(5 * nb):
        if spec.wn:
            kernel, bias, g = reader.take(3)
        else:
            (kernel, bias), g = reader.take(2), None
        out.append((nets._effective_kernel(kernel, g, spec.exp_norm), bias))
    kernel, bias = reader.take(2)
    out.append((kernel, bias))
    assert reader.i == len(params)
    return out


def _conv_rows(x_rows, kernel_hwio, bias, left, right):
    """'valid' convolution over a window of rows: no vertical padding (the caller supplies the halo rows), explicit
    horizontal zero padding like the reference.  x_rows [n, R, W, Cin] -> [n, R - kh + 1, W, Cout]"""
    xn = F.pad(x_rows.permute(0, 3, 1, 2), (left, right, 0, 0))
    return F.conv2d(xn, kernel_hwio.permute(3, 2, 0, 1), bias).permute(0, 2, 3, 1)


def _halo(cached, r0, rows):
    """rows [r0 - rows, r0) of a cached tensor [n, H, W, C]; rows above the lattice are the zero padding"""
    n, H, W, C = cached.shape
    lo = r0 - rows
    pieces = []
    if lo < 0:
        pieces.append(torch.zeros((n, min(-lo, rows), W, C), dtype=cached.dtype))
    if r0 > 0:
        pieces.append(cached[:, max(lo, 0):r0])
    return torch.cat(pieces, dim=1) if pieces else cached[:, :0]


def _block_rows(kernels5, v_new, h_new, r0, cache_b, k, last):
    """rows >= r0 of one block.  v_new / h_new: rows >= r0 of the block inputs (new configuration).
    cache_b: the sample's own (v_in, relu_vp, c) of this block, full height; only halo rows are read."""
    (kv, bv), (kx, bx), (kxx, bxx), (ky, by), (kh, bh) = kernels5
    pad = k - 1
    vp = _conv_rows(torch.cat([_halo(cache_b['v_in'], r0, pad), v_new], dim=1), kv, bv, pad // 2, pad // 2)
    x = torch.relu(_conv_rows(h_new, kx, bx, pad, 0))
    if last:
        x = nets._right_shift(x)
    x = _conv_rows(x, kxx, bxx, 0, 0)
    act_vp = torch.relu(vp)
    shifted = torch.cat([_halo(cache_b['relu_vp'], r0, 1), act_vp[:, :-1]], dim=1)       # DownShift: row i reads row i - 1
    y = _conv_rows(shifted, ky, by, 0, 0)
    c = torch.cat([torch.relu(x), torch.relu(y)], dim=-1)
    hp = _conv_rows(torch.cat([_halo(cache_b['c'], r0, pad), c], dim=1), kh, bh, pad, 0)
    return vp, hp


def forward_with_cache(spec, params, sigma):
    """Full forward of `sigma` [n, H, W]; returns (conditional log wave function [n, H, W, 2] complex, cache) where
    cache[b] = {'v_in', 'relu_vp', 'c'} are the tensors a later row-trimmed evaluation takes its halo rows from."""
    kernels = _kernels(spec, params)
    dtype = params[0].dtype
    nb = 2 * spec.depth - 2
    x0 = torch.as_tensor(sigma).to(dtype).unsqueeze(-1)
    cache = []

    def run_block(b, v, h, last=False):
        # the ordinary forward is the row-trimmed one with r0 = 0 (all halo rows are zero padding) -- but it must *fill* the
        # cache, so the intermediate tensors are recomputed here in the plain way
        (kv, bv), (kx, bx), (kxx, bxx), (ky, by), (kh, bh) = kernels[5 * b:5 * b + 5]
        pad = spec.k - 1
        vp = nets._conv2d_nhwc(v, kv, bv, (pad, 0, pad // 2, pad // 2))
        x = torch.relu(nets._conv2d_nhwc(h, kx, bx, (0, 0, pad, 0)))
        if last:
            x = nets._right_shift(x)
        x = nets._conv2d_nhwc(x, kxx, bxx, (0, 0, 0, 0))
        act_vp = torch.relu(vp)
        y = nets._conv2d_nhwc(nets._down_shift(act_vp), ky, by, (0, 0, 0, 0))
        c = torch.cat([torch.relu(x), torch.relu(y)], dim=-1)
        hp = nets._conv2d_nhwc(c, kh, bh, (pad, 0, pad, 0))
        cache.append({'v_in': v, 'relu_vp': act_vp, 'c': c})
        return vp, hp

    v, h = run_block(0, x0, x0)
    v, h = torch.relu(v), torch.relu(h)
    b = 1
    for _ in range(spec.depth - 2):
        v_in, h_in = v, h
        v, h = run_block(b, v, h)
        v, h = torch.relu(v), torch.relu(h)
        v, h = run_block(b + 1, v, h)
        v, h = torch.relu(v_in + v), torch.relu(h_in + h)
        b += 2
    _, x = run_block(nb - 1, v, h, last=True)
    kernel, bias = kernels[-1]
    logits = nets._conv2d_nhwc(torch.relu(x), kernel, bias, (0, 0, 0, 0))
    return _normalise(logits), cache


def _normalise(logits):
    re, im = logits[..., 0:2], logits[..., 2:4]
    return torch.complex(re - 0.5 * torch.logsumexp(2.0 * re, dim=-1, keepdim=True), im)


def _select(cond, sigma):
    idx = ((1 - torch.as_tensor(sigma).to(torch.int64)) // 2).unsqueeze(-1)
    return torch.complex(torch.gather(cond.real, -1, idx).squeeze(-1), torch.gather(cond.imag, -1, idx).squeeze(-1))


def log_psi_from_row(spec, params, sigma_new, r0, cache, cond_base, sigma_base):
    """log psi of configurations `sigma_new` [n, H, W] that agree with `sigma_base` on all rows < r0, recomputing only
    rows >= r0; `cache` / `cond_base` come from forward_with_cache(sigma_base)."""
    kernels = _kernels(spec, params)
    dtype = params[0].dtype
    nb = 2 * spec.depth - 2
    sigma_new = torch.as_tensor(sigma_new)
    x0 = sigma_new[:, r0:].to(dtype).unsqueeze(-1)
    k = spec.k

    v, h = _block_rows(kernels[0:5], x0, x0, r0, cache[0], k, False)
    v, h = torch.relu(v), torch.relu(h)
    b = 1
    for _ in range(spec.depth - 2):
        v_in, h_in = v, h
        v, h = _block_rows(kernels[5 * b:5 * b + 5], v, h, r0, cache[b], k, False)
        v, h = torch.relu(v), torch.relu(h)
        v, h = _block_rows(kernels[5 * (b + 1):5 * (b + 1) + 5], v, h, r0, cache[b + 1], k, False)
        v, h = torch.relu(v_in + v), torch.relu(h_in + h)
        b += 2
    _, x = _block_rows(kernels[5 * (nb - 1):5 * nb], v, h, r0, cache[nb - 1], k, True)
    kernel, bias = kernels[-1]
    cond_new = _normalise(_conv_rows(torch.relu(x), kernel, bias, 0, 0))            # rows >= r0
    prefix = _select(cond_base[:, :r0], torch.as_tensor(np.array(sigma_base))[:, :r0]).sum(dim=(1, 2)) if r0 > 0 else 0.0
    return prefix + _select(cond_new, sigma_new[:, r0:]).sum(dim=(1, 2))


def first_changed_row(sigma_base, sigma_new):
    """r0 per configuration: first row in which the two differ (H if they are equal)"""
    diff = (np.asarray(sigma_base) != np.asarray(sigma_new)).any(axis=-1)           # [n, H]
    H = diff.shape[1]
    return np.where(diff.any(axis=1), diff.argmax(axis=1), H)
