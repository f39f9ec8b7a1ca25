"""GPU tool: fk_jacobian_rows_tc (bf16, panel-major, weight-norm transform fused) against fk_grad_per_sample_tc rows;
the device sample-space SR pipeline against the torch route; timing on the headline machine (--big)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import numpy as np
import torch

from flowket_b200 import Input, Model, _lib, FK_ENGINE_TC
from flowket_b200.machines import ConvNetAutoregressive2D
from flowket_b200.optimizers.stochastic_reconfiguration import StochasticReconfiguration


def rows_from_panels(Xp, P):
    nkb, rld, _ = Xp.shape
    return Xp.permute(1, 0, 2).reshape(rld, nkb * 64)[:, :P]


def main():
    out = {}
    lib = _lib.require_cuda()
    for (H, W, depth, wn, B) in [(6, 6, 3, True, 96), (4, 5, 4, False, 64), (10, 10, 4, True, 128)]:
        inp = Input(shape=(H, W), dtype='int8')
        m = ConvNetAutoregressive2D(inp, depth=depth, num_of_channels=32, weights_normalization=wn, seed=2)
        model = Model(inputs=inp, outputs=m.predictions)
        model.engine = FK_ENGINE_TC
        flat = m.flat_params_device()
        g = torch.Generator(device='cpu').manual_seed(5)
        flat += (0.05 * torch.randn(flat.numel(), generator=g)).to(flat.device)
        m.params_updated()
        net = m.device_net()
        rng = np.random.RandomState(0)
        sg = net.to_sigma((2 * rng.randint(0, 2, size=(B, H, W)) - 1).astype(np.int8))
        O_re, O_im = net.grad_per_sample(sg, imag=True, engine=FK_ENGINE_TC)
        P = net.num_params
        nkb = (P + 63) // 64
        rld = 2 * B + 8
        Xp = torch.full((nkb, rld, 64), 9.0, dtype=torch.bfloat16, device=sg.device)
        wsb = lib.fk_jacobian_rows_tc_workspace_bytes(net.handle, B)
        ws = torch.empty(wsb, dtype=torch.uint8, device=sg.device)
        _lib.check(lib.fk_jacobian_rows_tc(net.handle, sg.data_ptr(), B, Xp.data_ptr(), rld, 0, B, ws.data_ptr(), wsb, _lib.stream_ptr()))
        rows = rows_from_panels(Xp, nkb * 64).float()
        want = torch.cat([O_re, O_im])
        got = rows[:2 * B, :P]
        err = (got - want).abs().max().item() / want.abs().max().item()
        rel = ((got - want).norm() / want.norm()).item()
        pad_ok = bool((rows[:2 * B, P:] == 0).all().item()) and bool((rows[2 * B:] == 9.0).all().item())
        print('%dx%d depth %d wn %d: bf16 rows vs fp32 TC rows: max %.2e (of max |O|), Frobenius %.2e, padding ok %s'
              % (H, W, depth, wn, err, rel, pad_ok), flush=True)
        assert err < 1e-2 and rel < 6e-3 and pad_ok
        # full pipeline against the torch route on the same machine
        e = torch.complex(torch.randn(B, dtype=torch.float64), torch.randn(B, dtype=torch.float64)).to(sg.device)
        sr_dev = StochasticReconfiguration(model, sample_space=True)
        sr_ref = StochasticReconfiguration(model, sample_space=True, device_pipeline=False, gram_dtype='fp32')
        d_dev = sr_dev.compute_update(sg, e)
        d_ref = sr_ref.compute_update(sg, e)
        derr = ((d_dev - d_ref).norm() / d_ref.norm()).item()
        print('   delta (device pipeline, bf16 rows) vs torch route (fp32 rows, fp32 Gram): %.2e   timings %s'
              % (derr, {k: round(v, 2) for k, v in sr_dev.last_timings_ms.items()}), flush=True)
        out['delta_err_%dx%d_d%d_wn%d' % (H, W, depth, wn)] = derr
        assert derr < 3e-2
    if '--big' in sys.argv:
        inp = Input(shape=(10, 10), dtype='int8')
        m = ConvNetAutoregressive2D(inp, depth=20, num_of_channels=32, seed=0)
        model = Model(inputs=inp, outputs=m.predictions)
        model.engine = FK_ENGINE_TC
        net = m.device_net()
        B = 8192
        sg = net.sample(B, seed=3, engine=FK_ENGINE_TC)
        e = torch.complex(torch.randn(B, dtype=torch.float64), torch.randn(B, dtype=torch.float64)).to(sg.device)
        sr = StochasticReconfiguration(model, sample_space=True)
        for it in range(3):
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            d = sr.compute_update(sg, e)
            t1.record()
            torch.cuda.synchronize()
            print('headline SR update: %.1f ms  %s  |delta| %.4e  peak mem %.1f GB' % (
                t0.elapsed_time(t1), {k: round(v, 1) for k, v in sr.last_timings_ms.items()}, d.norm().item(),
                torch.cuda.max_memory_allocated() / 1e9), flush=True)
        out['sr_update_ms'] = t0.elapsed_time(t1)
        out['sr_timings'] = sr.last_timings_ms
        if '--ref' in sys.argv:
            sr_ref = StochasticReconfiguration(model, sample_space=True, device_pipeline=False)
            d_ref = sr_ref.compute_update(sg, e)
            print('vs torch route (fp32 rows -> bf16, cuBLAS Gram): %.2e' % ((d - d_ref).norm() / d_ref.norm()).item(), flush=True)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
