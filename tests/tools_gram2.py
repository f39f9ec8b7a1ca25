"""GPU tool (not a pytest file): correctness and timing of the sample-space SR kernels behind the C ABI --
fk_sr_gram_xxt (hand-written cta_group::2 tcgen05 GEMM), fk_sr_centre_shift, fk_sr_xt_w, fk_sr_solve -- against torch.
Usage: python tests/tools_gram2.py [--big]"""
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from flowket_b200 import _lib


def to_panels(Xrow, K, rld):
    """row-major [R, >= K] bf16 -> panel-major [ceil(K/64)][rld][64] (zero padded columns / junk padding rows)"""
    R = Xrow.shape[0]
    nkb = (K + 63) // 64
    Xp = torch.full((nkb, rld, 64), 3.0, dtype=torch.bfloat16, device=Xrow.device)
    pad = torch.zeros((R, nkb * 64), dtype=torch.bfloat16, device=Xrow.device)
    pad[:, :K] = Xrow[:, :K]
    Xp[:, :R, :] = pad.view(R, nkb, 64).permute(1, 0, 2)
    return Xp


def gram_xxt(lib, Xp, R, K, scale=1.0):
    rld = Xp.shape[1]
    G = torch.empty((R, R), dtype=torch.float32, device=Xp.device)
    wsb = lib.fk_sr_gram_xxt_workspace_bytes(R)
    ws = torch.empty(wsb, dtype=torch.uint8, device=Xp.device)
    _lib.check(lib.fk_sr_gram_xxt(Xp.data_ptr(), R, K, rld, 1, 0, scale, G.data_ptr(), R, ws.data_ptr(), wsb, _lib.stream_ptr()))
    return G


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        fn()
        t1.record()
        torch.cuda.synchronize()
        best = min(best, t0.elapsed_time(t1))
    return best


def main():
    lib = _lib.require_cuda()
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    out = {}
    for (R, K, ld) in [(512, 1000, 1000), (256, 64, 64), (700, 4100, 4104), (2048, 70000, 70016), (1300, 33000, 33008)]:
        X = torch.zeros((R, ld), dtype=torch.bfloat16, device=dev)
        X[:, :K] = (torch.randn((R, K), device=dev) * (1 + torch.rand((R, 1), device=dev))).to(torch.bfloat16)
        Xp = to_panels(X, K, ld % 977 + R)      # rld > R: padding rows must be ignored
        G = gram_xxt(lib, Xp, R, K, 0.5)
        ref = 0.5 * (X[:, :K].double() @ X[:, :K].double().T)
        err = ((G.double() - ref).abs().max() / ref.abs().max()).item()
        sym = (G - G.T).abs().max().item()
        print('gram_xxt R=%d K=%d ld=%d: max err / max |G| = %.2e, asymmetry %.1e' % (R, K, ld, err, sym), flush=True)
        out['gram_%d_%d' % (R, K)] = err
        assert err < 1e-4 and sym == 0.0, (err, sym)
        # centring + shift
        B = R // 2
        if R % 2 == 0:
            wsb = lib.fk_sr_centre_shift_workspace_bytes(R)
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            S = torch.empty((R, R), dtype=torch.float64, device=dev)
            _lib.check(lib.fk_sr_centre_shift(G.data_ptr(), R, R, 1, 0.05, S.data_ptr(), ws.data_ptr(), wsb, _lib.stream_ptr()))
            C = torch.eye(B, dtype=torch.float64, device=dev) - 1.0 / B
            Cf = torch.block_diag(C, C)
            Sref = Cf @ G.double() @ Cf / B + 0.05 * torch.eye(R, dtype=torch.float64, device=dev)
            e2 = ((S - Sref).abs().max() / Sref.abs().max()).item()
            print('  centre_shift: %.2e' % e2, flush=True)
            assert e2 < 1e-5, e2
            # solve
            h = ctypes.c_void_p()
            _lib.check(lib.fk_sr_solver_create(ctypes.byref(h)))
            wsb = lib.fk_sr_solve_workspace_bytes(h, R)
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            rhs = torch.randn(R, dtype=torch.float64, device=dev)
            x = rhs.clone()
            info = torch.full((1,), -7, dtype=torch.int32, device=dev)
            Sf = S.clone()
            _lib.check(lib.fk_sr_solve(h, Sf.data_ptr(), x.data_ptr(), R, info.data_ptr(), ws.data_ptr(), wsb, _lib.stream_ptr()))
            xref = torch.linalg.solve(S, rhs)
            e3 = ((x - xref).abs().max() / xref.abs().max()).item()
            r_mine = ((S @ x - rhs).abs().max() / rhs.abs().max()).item()
            r_ref = ((S @ xref - rhs).abs().max() / rhs.abs().max()).item()
            print('  solve: %.2e info %d; residuals: fk_sr_solve %.1e, torch.linalg.solve %.1e' % (e3, info.item(), r_mine, r_ref), flush=True)
            assert r_mine < 1e-10 and info.item() == 0
            _lib.check(lib.fk_sr_solver_destroy(h))
        # X^T w
        w = torch.randn(R, device=dev)
        o = torch.empty(K, dtype=torch.float32, device=dev)
        _lib.check(lib.fk_sr_xt_w(Xp.data_ptr(), R, K, Xp.shape[1], 1, 0, w.data_ptr(), o.data_ptr(), _lib.stream_ptr()))
        oref = X[:, :K].double().T @ w.double()
        e4 = ((o.double() - oref).abs().max() / oref.abs().max()).item()
        print('  xt_w: %.2e' % e4, flush=True)
        assert e4 < 1e-5, e4

    # row blocks (the sharded step): 3 blocks of 256 rows [Re ; Im] each, in separate buffers of one allocation
    nbk, br, K = 3, 256, 5000
    R = nbk * br
    X = (torch.randn((R, K), device=dev) + 0.3).to(torch.bfloat16)
    nkb = (K + 63) // 64
    rld = br + 128
    buf = torch.full((nbk, nkb + 2, rld, 64), 5.0, dtype=torch.bfloat16, device=dev)      # (+ 2 junk panels between the blocks)
    for q in range(nbk):
        buf[q, :nkb] = to_panels(X[q * br:(q + 1) * br], K, rld)
    stride = buf.stride(0) * 2
    G = torch.empty((R, R), dtype=torch.float32, device=dev)
    wsb = lib.fk_sr_gram_xxt_workspace_bytes(R)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    _lib.check(lib.fk_sr_gram_xxt(buf.data_ptr(), R, K, rld, nbk, stride, 1.0, G.data_ptr(), R, ws.data_ptr(), wsb, _lib.stream_ptr()))
    ref = X.double() @ X.double().T
    err = ((G.double() - ref).abs().max() / ref.abs().max()).item()
    print('gram_xxt with 3 row blocks: %.2e' % err, flush=True)
    assert err < 1e-4
    wsb = lib.fk_sr_centre_shift_workspace_bytes(R)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    S = torch.empty((R, R), dtype=torch.float64, device=dev)
    _lib.check(lib.fk_sr_centre_shift(G.data_ptr(), R, R, nbk, 0.05, S.data_ptr(), ws.data_ptr(), wsb, _lib.stream_ptr()))
    half = ((torch.arange(R, device=dev) % br) >= br // 2)
    Cf = torch.eye(R, dtype=torch.float64, device=dev)
    for hsel in (half, ~half):
        idx = hsel.nonzero().reshape(-1)
        Cf[idx[:, None], idx[None, :]] -= 1.0 / idx.numel()
    Sref = Cf @ G.double() @ Cf / (R // 2) + 0.05 * torch.eye(R, dtype=torch.float64, device=dev)
    e2 = ((S - Sref).abs().max() / Sref.abs().max()).item()
    w = torch.randn(R, device=dev)
    o = torch.empty(K, dtype=torch.float32, device=dev)
    _lib.check(lib.fk_sr_xt_w(buf.data_ptr(), R, K, rld, nbk, stride, w.data_ptr(), o.data_ptr(), _lib.stream_ptr()))
    oref = X.double().T @ w.double()
    e4 = ((o.double() - oref).abs().max() / oref.abs().max()).item()
    print('  centre_shift (blocks): %.2e, xt_w (blocks): %.2e' % (e2, e4), flush=True)
    assert e2 < 1e-12 and e4 < 1e-5

    if '--big' in sys.argv:
        R, K = 16384, 854016
        ld = (K + 127) // 128 * 128
        X = torch.empty((R, ld), dtype=torch.bfloat16, device=dev)
        for r in range(0, R, 1024):
            X[r:r + 1024] = (torch.randn((1024, ld), device=dev) * 0.3 + 0.1).to(torch.bfloat16)
        X[:, K:] = 0
        Xp = torch.empty((ld // 64, R, 64), dtype=torch.bfloat16, device=dev)
        for r in range(0, R, 1024):
            Xp[:, r:r + 1024, :] = X[r:r + 1024].view(1024, ld // 64, 64).permute(1, 0, 2)
        G = gram_xxt(lib, Xp, R, K)
        torch.cuda.synchronize()
        # exact fp64 reference on two 256 x 256 blocks (diagonal and off-diagonal)
        for (i0, j0) in [(0, 0), (512, 9216), (16128, 16128), (1024, 16000)]:
            ref = X[i0:i0 + 256, :K].double() @ X[j0:j0 + 256, :K].double().T
            blk = G[i0:i0 + 256, j0:j0 + 256].double()
            err = ((blk - ref).abs().max() / ref.abs().max()).item()
            bias = ((blk - ref).sum() / ref.abs().sum()).item()
            mir = (G[j0:j0 + 256, i0:i0 + 256].T.double() - blk).abs().max().item()
            print('big block (%d,%d): max rel err %.2e, signed bias %.2e, mirror diff %.1e' % (i0, j0, err, bias, mir), flush=True)
            out['big_err_%d_%d' % (i0, j0)] = err
            out['big_bias_%d_%d' % (i0, j0)] = bias
        T = torch.mm(X[:2048], X[:4096].T, out_dtype=torch.float32)
        e = ((T - G[:2048, :4096]).abs().max() / T.abs().max()).item()
        ref = X[:256, :K].double() @ X[:256, :K].double().T
        cb = ((T[:256, :256].double() - ref).sum() / ref.abs().sum()).item()
        print('vs torch.mm (cuBLAS bf16->fp32) block: %.2e; cuBLAS signed bias on block (0,0): %.2e' % (e, cb), flush=True)
        out['cublas_bias'] = cb
        del T
        wsb = lib.fk_sr_gram_xxt_workspace_bytes(R)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)

        def mine():
            _lib.check(lib.fk_sr_gram_xxt(Xp.data_ptr(), R, K, R, 1, 0, 1.0, G.data_ptr(), R, ws.data_ptr(), wsb, _lib.stream_ptr()))
        t_mine = timed(mine)
        nt = R // 256
        flops_tri = 2.0 * (nt * (nt + 1) / 2) * 256 * 256 * K
        print('fk_sr_gram_xxt: %.1f ms, %.0f TFLOP/s on the computed upper block triangle (%.0f TFLOP/s full-matrix equivalent)'
              % (t_mine, flops_tri / t_mine / 1e9, 2.0 * R * R * K / t_mine / 1e9), flush=True)
        out['gram_ms'] = t_mine
        out['gram_tflops_tri'] = flops_tri / t_mine / 1e9
        from flowket_b200.optimizers.stochastic_reconfiguration import StochasticReconfiguration
        t_cublas = timed(lambda: StochasticReconfiguration._symmetric_gram(X), reps=2)
        nb = R // 2048
        flops_cb = sum(2.0 * 2048 * (R - i * 2048) * K for i in range(nb))
        print('cuBLAS block-triangle (torch.mm, 2048-row blocks): %.1f ms, %.0f TFLOP/s' % (t_cublas, flops_cb / t_cublas / 1e9), flush=True)
        out['cublas_ms'] = t_cublas
        out['cublas_tflops'] = flops_cb / t_cublas / 1e9
        w = torch.randn(R, device=dev)
        o = torch.empty(K, dtype=torch.float32, device=dev)
        t_xtw = timed(lambda: _lib.check(lib.fk_sr_xt_w(Xp.data_ptr(), R, K, R, 1, 0, w.data_ptr(), o.data_ptr(), _lib.stream_ptr())))
        print('fk_sr_xt_w: %.2f ms, %.0f GB/s' % (t_xtw, R * K * 2 / t_xtw / 1e6), flush=True)
        out['xtw_ms'] = t_xtw
        out['xtw_gbs'] = R * K * 2 / t_xtw / 1e6
        S = torch.empty((R, R), dtype=torch.float64, device=dev)
        wsb2 = lib.fk_sr_centre_shift_workspace_bytes(R)
        ws2 = torch.empty(wsb2, dtype=torch.uint8, device=dev)
        t_cs = timed(lambda: _lib.check(lib.fk_sr_centre_shift(G.data_ptr(), R, R, 1, 0.05, S.data_ptr(), ws2.data_ptr(), wsb2, _lib.stream_ptr())))
        print('fk_sr_centre_shift: %.2f ms' % t_cs, flush=True)
        out['centre_ms'] = t_cs
        h = ctypes.c_void_p()
        _lib.check(lib.fk_sr_solver_create(ctypes.byref(h)))
        wsb3 = lib.fk_sr_solve_workspace_bytes(h, R)
        ws3 = torch.empty(wsb3, dtype=torch.uint8, device=dev)
        rhs = torch.randn(R, dtype=torch.float64, device=dev)
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        S0 = S.clone()

        def solve():
            S.copy_(S0)
            _lib.check(lib.fk_sr_solve(h, S.data_ptr(), rhs.data_ptr(), R, info.data_ptr(), ws3.data_ptr(), wsb3, _lib.stream_ptr()))
        t_solve = timed(solve)
        print('fk_sr_solve (fp64 potrf + potrs, incl. a 2 GB copy): %.1f ms, info %d' % (t_solve, info.item()), flush=True)
        out['solve_ms'] = t_solve
    print(json.dumps(out))


if __name__ == '__main__':
    main()
