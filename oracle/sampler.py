"""CPU restatement of the reference's autoregressive samplers.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Restates (relative to /root/reference/src/flowket):
  * AutoregressiveSampler.__next__ (explicit-uniform rule, +-1 variant)   deepar/samplers/autoregressive.py:29-48,
                                                                        samplers/__init__.py:8-14
  * raster ordering                                                     deepar/ordering/raster.py:4-5
  * FastAutoregressiveSampler (each (layer, site) activation computed once, one matmul per
    conv per site)                deepar/samplers/fast_autoregressive.py:63-76,
                                  deepar/graph_analysis/convolutional_topology.py:22-30
The parity contract for "identical uniform draws": sigma_s = +1  <=>  float32 exp(log p(s, class 0)) > u_s,
sites visited in raster order, unsampled sites hold 0.
"""
import itertools

import numpy as np
import torch

from . import nets


def sample_with_uniforms(spec, params, uniforms, mini_batch_size=None):
    """N full forwards, one per site.  uniforms: float64 [B, *input_shape].  -> (int8 sigma, p0 [B,*shape])."""
    uniforms = np.asarray(uniforms)
    B = uniforms.shape[0]
    shape = tuple(spec.input_shape)
    batch = np.zeros((B,) + shape, dtype=np.float64)
    p0_all = np.zeros((B,) + shape, dtype=np.float64)
    mb = B if mini_batch_size is None else mini_batch_size
    with torch.no_grad():
        for site in itertools.product(*[range(d) for d in shape]):
            idx = (slice(None),) + site
            lp = np.concatenate([nets.conditional_log_probs(spec, params, batch[i:i + mb]).numpy()
                                 for i in range(0, B, mb)])
            p0 = np.exp(lp[idx + (0,)])
            p0_all[idx] = p0
            batch[idx] = 2.0 * (p0 > uniforms[idx]) - 1.0
    return batch.astype(np.int8), p0_all


class IncrementalSampler2D(object):
    """Cached per-site formulation for ConvNetAutoregressive2D: every conv is evaluated at one spatial
    location as `stack(deps)[mb, kh*kw*cin] @ kernel.reshape(kh*kw*cin, cout) + bias`, exactly the
    per-(layer, site) matmul of convolutional_topology.py:22-30.  Used as the CPU baseline of the
    fast sampler and to check incremental == full.

    Schedule (follows from the dependency graph the reference builds):
      site (i, j):  last block + head at (i, j)  ->  draw sigma(i, j)  ->  horizontal stack of blocks
                    0..nb-2 at (i, j);
      end of row i: vertical stack of row i for all blocks (it sees the whole row)."""

    def __init__(self, spec, params):
        assert spec.kind == 'conv2d' and spec.k == 3
        self.spec = spec
        self.dtype = params[0].dtype
        reader = nets._ParamReader(params)
        self.convs = []
        for (kh, kw, cin, cout, wn) in spec.conv_shapes():
            if wn:
                kernel, bias, g = reader.take(3)
            else:
                (kernel, bias), g = reader.take(2), None
            w = nets._effective_kernel(kernel, g, spec.exp_norm).reshape(kh * kw * cin, cout)
            self.convs.append((w, bias))

    def sample(self, uniforms):
        spec = self.spec
        H, W, C = spec.H, spec.W, spec.C
        nb = spec.num_blocks
        u = torch.as_tensor(np.asarray(uniforms))
        B = u.shape[0]
        dt = self.dtype
        OFF = 3  # zero border: rows/cols -3..-1 and one extra column on the right

        def zeros(c):
            return torch.zeros((B, H + OFF, W + OFF + 1, c), dtype=dt)

        vin = [zeros(1 if b == 0 else C) for b in range(nb)]   # input of the vertical conv of block b
        hin = [zeros(1 if b == 0 else C) for b in range(nb)]   # input of the horizontal conv of block b
        cc = [zeros(C) for _ in range(nb)]                      # concat tensor of block b
        relu_vp = [zeros(C) for _ in range(nb)]                 # relu(vertical conv output) of block b
        out = np.zeros((B, H, W), np.int8)
        p0_all = np.zeros((B, H, W))

        def gather(t, i, j, rows, cols):
            return torch.cat([t[:, i + OFF + a, j + OFF + b] for a in rows for b in cols], dim=-1)

        def store_next(buf, b, i, j, pre):
            """activation / residual wiring of _build_unnormalized_conditional_log_wave_function:
            block 0 -> relu; residual pairs (1,2), (3,4), ...: odd -> relu, even -> relu(pair input + pre)."""
            if b + 1 >= nb:
                return
            if b == 0 or b % 2 == 1:
                buf[b + 1][:, i + OFF, j + OFF] = torch.relu(pre)
            else:
                buf[b + 1][:, i + OFF, j + OFF] = torch.relu(buf[b - 1][:, i + OFF, j + OFF] + pre)

        def block_h(b, i, j):
            """horizontal stack of block b at (i, j) -> pre-activation h'(i, j)."""
            wx, bx = self.convs[5 * b + 1]
            wxx, bxx = self.convs[5 * b + 2]
            wy, by = self.convs[5 * b + 3]
            wh, bh = self.convs[5 * b + 4]
            if b == nb - 1:   # RightShift between the 1xk conv and the 1x1 conv: uses x1(i, j-1)
                if j > 0:
                    x1 = torch.relu(gather(hin[b], i, j - 1, [0], [-2, -1, 0]) @ wx + bx)
                else:
                    x1 = torch.zeros((B, C), dtype=dt)
            else:
                x1 = torch.relu(gather(hin[b], i, j, [0], [-2, -1, 0]) @ wx + bx)
            cx = torch.relu(x1 @ wxx + bxx)
            cy = torch.relu(relu_vp[b][:, i - 1 + OFF, j + OFF] @ wy + by)   # DownShift: row i-1 (zero for i=0)
            cc[b][:, i + OFF, j + OFF] = torch.cat([cx, cy], dim=-1)
            return gather(cc[b], i, j, [-2, -1, 0], [-2, -1, 0]) @ wh + bh

        def row_v(i):
            for b in range(nb):
                wv, bv = self.convs[5 * b]
                for j in range(W):
                    vp = gather(vin[b], i, j, [-2, -1, 0], [-1, 0, 1]) @ wv + bv
                    relu_vp[b][:, i + OFF, j + OFF] = torch.relu(vp)
                    store_next(vin, b, i, j, vp)

        wo, bo = self.convs[-1]
        with torch.no_grad():
            for i in range(H):
                for j in range(W):
                    logits = torch.relu(block_h(nb - 1, i, j)) @ wo + bo
                    re = logits[:, 0:2]
                    logp0 = 2.0 * (re[:, 0] - 0.5 * torch.logsumexp(2.0 * re, dim=-1))
                    p0 = torch.exp(logp0)
                    s = torch.where(p0 > u[:, i, j].to(dt), torch.ones_like(p0), -torch.ones_like(p0))
                    out[:, i, j] = s.numpy().astype(np.int8)
                    p0_all[:, i, j] = p0.numpy()
                    vin[0][:, i + OFF, j + OFF, 0] = s
                    hin[0][:, i + OFF, j + OFF, 0] = s
                    for b in range(nb - 1):
                        store_next(hin, b, i, j, block_h(b, i, j))
                row_v(i)
        return out, p0_all
