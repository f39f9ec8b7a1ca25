"""GPU tool: tc-exact engine (fk_tc_exact.cu) vs the fp32 engine and the fp64 oracle; work-list local energy; timing."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from flowket_b200 import Input, Model, FK_ENGINE_FP32, FK_ENGINE_TC, FK_ENGINE_TC_EXACT
from flowket_b200.machines import ConvNetAutoregressive2D
from flowket_b200.operators import Heisenberg, Ising
from oracle import nets


def main():
    out = {}
    for (H, W, depth, wn) in [(4, 4, 3, False), (6, 6, 4, True), (10, 10, 20, True), (10, 10, 20, False), (5, 7, 6, True)]:
        inp = Input(shape=(H, W), dtype='int8')
        m = ConvNetAutoregressive2D(inp, depth=depth, num_of_channels=32, weights_normalization=wn, seed=1)
        model = Model(inputs=inp, outputs=m.predictions)
        # perturb biases / weights a little so that nothing is at its zero-initialised value
        flat = m.flat_params_device()
        g = torch.Generator(device='cpu').manual_seed(5)
        flat += (0.02 * torch.randn(flat.numel(), generator=g)).to(flat.device)
        m.params_updated()
        net = m.device_net()
        rng = np.random.RandomState(0)
        n = 24
        sigma = (2 * rng.randint(0, 2, size=(n, H, W)) - 1).astype(np.int8)
        sg = net.to_sigma(sigma)
        lp32 = net.log_psi(sg, engine=FK_ENGINE_FP32)
        lp16 = net.log_psi(sg, engine=FK_ENGINE_TC)
        lpx = net.log_psi(sg, engine=FK_ENGINE_TC_EXACT)
        spec = nets.Conv2DSpec(H, W, depth, 32, weights_normalization=wn)
        params = [torch.from_numpy(w.astype(np.float64)) for w in m.get_weights()]
        want = torch.from_numpy(np.asarray(nets.log_psi_numpy(spec, params, sigma.astype(np.float64)))).reshape(-1).to(lp32.device)
        def rel(a):
            return ((a.to(torch.complex128) - want).abs().max() / want.abs().max()).item()
        print('%dx%d depth %d wn %d: fp32 %.2e, tc %.2e, tc-exact %.2e (max |dlogpsi| / max |logpsi| vs fp64 oracle)'
              % (H, W, depth, wn, rel(lp32), rel(lp16), rel(lpx)), flush=True)
        out['logpsi_%dx%d_d%d_wn%d' % (H, W, depth, wn)] = {'fp32': rel(lp32), 'tc': rel(lp16), 'tcx': rel(lpx)}
        # local energy through the work list, both tensor-core engines vs the fp32 engine
        for op in (Heisenberg(hilbert_state_shape=[H, W], pbc=False), Ising(hilbert_state_shape=[H, W], pbc=False, h=3.0)):
            e32, st32, nc32 = net.local_energy(op.device_desc(), sg, engine=FK_ENGINE_FP32)
            e16, st16, nc16 = net.local_energy(op.device_desc(), sg, engine=FK_ENGINE_TC)
            ex, stx, ncx = net.local_energy(op.device_desc(), sg, engine=FK_ENGINE_TC_EXACT)
            s = e32.abs().max()
            print('   %s: E_loc tc %.2e, tc-exact %.2e (per-sample max / max|E|); n_conn %d %d %d; stats diff %.1e'
                  % (type(op).__name__, ((e16 - e32).abs().max() / s).item(), ((ex - e32).abs().max() / s).item(), nc32, nc16, ncx,
                     ((stx - st32).abs() / st32.abs().clamp_min(1e-30)).max().item()), flush=True)
            out['eloc_%s_%dx%d_d%d_wn%d' % (type(op).__name__, H, W, depth, wn)] = {
                'tc': ((e16 - e32).abs().max() / s).item(), 'tcx': ((ex - e32).abs().max() / s).item()}
            assert nc32 == nc16 == ncx
    # timing on the headline machine
    inp = Input(shape=(10, 10), dtype='int8')
    m = ConvNetAutoregressive2D(inp, depth=20, num_of_channels=32, seed=0)
    net = m.device_net()
    rng = np.random.RandomState(1)
    sg = net.to_sigma((2 * rng.randint(0, 2, size=(65536, 10, 10)) - 1).astype(np.int8))
    for eng, name in ((FK_ENGINE_TC, 'tc'), (FK_ENGINE_TC_EXACT, 'tcx')):
        net.log_psi(sg, engine=eng)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        net.log_psi(sg, engine=eng)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1)
        print('%s: 65536 configurations in %.2f ms = %.2f M psi/s' % (name, ms, 65536 / ms / 1e3), flush=True)
        out['psi_per_s_' + name] = 65536 / ms * 1e3
    op = Heisenberg(hilbert_state_shape=[10, 10], pbc=False)
    sb = sg[:8192].contiguous()
    for eng, name in ((FK_ENGINE_TC, 'tc'), (FK_ENGINE_TC_EXACT, 'tcx')):
        net.local_energy(op.device_desc(), sb, engine=eng)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        e, st, nc = net.local_energy(op.device_desc(), sb, engine=eng)
        t1.record()
        torch.cuda.synchronize()
        print('%s: E_loc of 8192 samples (%d psi evaluations) in %.1f ms' % (name, nc, t0.elapsed_time(t1)), flush=True)
        out['eloc_ms_' + name] = t0.elapsed_time(t1)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
