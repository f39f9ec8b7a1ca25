"""GPU tier, round 2: the contract-accuracy tensor-core engine (tc-exact), the work-list local energy, the device
sample-space SR pipeline (bf16 Jacobian rows -> hand-written tcgen05 Gram -> centring -> fp64 solve -> X^T w) and the
exact-enumeration kernels, each against the fp64 oracle / a float64 restatement of the same formula.
Every stated tolerance is <= 3x (contract bounds excepted) the error measured on a B200; the measured value is printed."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import nets, operators as oops, local_energy as oeloc
from tests.helpers import make_pair, random_sigma

pytestmark = pytest.mark.gpu


def _perturbed(shape, depth, wn, seed=1):
    """product model + oracle (spec, params) of a 32-channel ConvNetAutoregressive2D with every tensor perturbed"""
    return make_pair('conv2d', shape, depth, 32, seed=seed, weights_normalization=wn)


# ---------------------------------------------------------------------------------------------------------------------
# tc-exact: fp16 hi+lo split operands (22 bits), two MMAs per k-step, fp32 accumulate.  Contract: 1e-5 relative on log psi
# (north_star); measured <= 1.3e-6 at depth 20 (tests/tools_tcx.py, gpurun_out/r02_tcx_a.log)
# ---------------------------------------------------------------------------------------------------------------------
TCX_CASES = [((4, 4), 3, False), ((6, 6), 4, True), ((5, 7), 6, True), ((10, 10), 5, True), ((10, 10), 20, True),
             ((10, 10), 20, False), ((3, 10), 2, True), ((9, 2), 3, False)]


@pytest.mark.parametrize('shape,depth,wn', TCX_CASES)
def test_tc_exact_log_psi_matches_oracle(shape, depth, wn):
    from flowket_b200 import FK_ENGINE_TC_EXACT
    model, _, spec, params = _perturbed(shape, depth, wn)
    sigma = random_sigma(37, shape, seed=3)
    want = nets.log_psi_numpy(spec, params, sigma)[:, 0]
    model.engine = FK_ENGINE_TC_EXACT
    got = model.predict(sigma)[:, 0]
    rel = np.abs(got - want).max() / np.abs(want).max()
    print('tc-exact log psi', shape, depth, wn, 'rel', rel)
    assert rel < 1e-5
    # Im log psi is a multiple of pi/2-type phase sum: compare directly as well
    assert np.abs(got.imag - want.imag).max() < 1e-4


@pytest.mark.parametrize('opkind,opkw,shape,depth', [
    ('heisenberg', dict(pbc=False), (4, 4), 3), ('ising', dict(pbc=False, h=3.0), (6, 6), 4),
    ('heisenberg', dict(pbc=True), (4, 6), 3), ('heisenberg', dict(pbc=False), (10, 10), 20),
    ('j1j2', dict(pbc=False, j2=0.5), (4, 4), 2)])
def test_tc_exact_local_energy_matches_oracle(opkind, opkw, shape, depth):
    """E_loc through the work list (flips generated in the forward kernel, ratio + segmented sum in its epilogue) on the
    tc-exact engine vs the fp64 oracle: contract 1e-4 per sample, measured <= 8e-6"""
    from flowket_b200 import FK_ENGINE_TC_EXACT, FK_ENGINE_FP32, FK_ENGINE_TC
    from flowket_b200.observables.monte_carlo import Observable
    from tests.test_gpu_parity import _product_operator
    model, _, spec, params = _perturbed(shape, depth, True, seed=2)
    n = 6 if depth >= 20 else 40        # the fp64 oracle evaluates ~90 connected configurations per sample on the host
    oop = oops.OracleOperator(opkind, shape, **opkw)
    if opkind == 'heisenberg':      # total S_z = 0 sector, like the sampler would give
        rng = np.random.RandomState(5)
        N = int(np.prod(shape))
        sigma = np.stack([rng.permutation(np.r_[np.ones(N // 2), -np.ones(N - N // 2)]) for _ in range(n)]).reshape((n,) + shape)
    else:
        sigma = random_sigma(n, shape, seed=5).astype(np.float64)
    want = oeloc.local_values(oop, lambda c: nets.log_psi_numpy(spec, params, c), sigma)
    op = _product_operator(opkind, shape, opkw)
    net = model.machine.device_net()
    sg = net.to_sigma(sigma.astype(np.int8))
    res = {}
    for name, eng in (('fp32', FK_ENGINE_FP32), ('tc', FK_ENGINE_TC), ('tcx', FK_ENGINE_TC_EXACT)):
        e, stats, n_conn = net.local_energy(op.device_desc(), sg, engine=eng)
        res[name] = (e.cpu().numpy(), stats.cpu().numpy(), n_conn)
    scale = np.abs(want).max()
    err = {k: np.abs(v[0] - want).max() / scale for k, v in res.items()}
    print('E_loc vs oracle', opkind, shape, depth, {k: '%.2e' % v for k, v in err.items()})
    assert err['tcx'] < 1e-4 and err['fp32'] < 1e-4
    assert err['tc'] < 2e-2
    # the three engines walk the same connections; the fused statistics (sum E, sum |E|^2 ...) agree with the samples
    assert res['fp32'][2] == res['tc'][2] == res['tcx'][2]
    assert np.allclose(res['tcx'][1], res['fp32'][1], rtol=1e-4, atol=1e-4 * scale * scale * n)
    # Observable route (the call VariationalMonteCarlo makes)
    model.engine = FK_ENGINE_TC_EXACT
    via_api = Observable(op).local_values(model, sigma.astype(np.int8))
    assert np.abs(via_api - want).max() / scale < 1e-4


@pytest.mark.parametrize('opkind,opkw,shape,depth,B', [
    ('heisenberg', dict(pbc=False), (10, 10), 3, 300), ('heisenberg', dict(pbc=True), (4, 6), 2, 257),
    ('ising', dict(pbc=False, h=1.0), (5, 7), 3, 128), ('j1j2', dict(pbc=False, j2=0.5), (6, 6), 2, 200),
    ('heisenberg', dict(pbc=False), (4, 4), 2, 8192 + 77), ('ising', dict(pbc=True, h=3.0), (3, 10), 2, 64),
    ('heisenberg', dict(pbc=False), (9, 2), 2, 100)])
@pytest.mark.parametrize('engine_name', ['tc_exact', 'tc_fp16'])
def test_prefix_reuse_equals_the_full_evaluation(opkind, opkw, shape, depth, B, engine_name, monkeypatch):
    """prefix reuse (rows above the first flipped site come from the sample's cached activations, two row-trimmed
    configurations per tile) against the same engine evaluating every connected configuration in full: identical arithmetic on
    the recomputed rows, so the two agree to the fp32 round-off of the log psi sums, which are taken in a different order
    (measured <= 4e-6 of max |E_loc|; both are within 1e-5 of the fp64 oracle, test above); the 4x4
    Heisenberg case crosses the 8192-sample chunk boundary of the activation cache"""
    from flowket_b200 import FK_ENGINE_TC_EXACT, FK_ENGINE_TC
    from tests.test_gpu_parity import _product_operator
    engine = FK_ENGINE_TC_EXACT if engine_name == 'tc_exact' else FK_ENGINE_TC
    model, _, spec, params = _perturbed(shape, depth, True, seed=3)
    net = model.machine.device_net()
    op = _product_operator(opkind, shape, opkw)
    sg = net.sample(B, seed=5)          # fp32 sampler: samples of the machine itself (total S_z is whatever it draws)
    res = {}
    for flag in ('1', '0'):
        monkeypatch.setenv('FK_PREFIX_REUSE', flag)
        e, stats, n_conn = net.local_energy(op.device_desc(), sg, engine=engine)
        res[flag] = (e.cpu().numpy(), stats.cpu().numpy(), n_conn)
    scale = np.abs(res['0'][0]).max()
    err = np.abs(res['1'][0] - res['0'][0]).max() / scale
    print('prefix reuse vs full evaluation', engine_name, opkind, shape, depth, B, 'err', err)
    assert res['1'][2] == res['0'][2]
    assert err < 2e-5
    assert np.allclose(res["1"][1], res["0"][1], rtol=1e-5, atol=1e-5 * scale * scale * B)


def test_tc_exact_rejects_unsupported_machines():
    """two-tile lattices (12 x 12) are outside the tc-exact envelope: the call must fail loudly, not fall back"""
    from flowket_b200 import FK_ENGINE_TC_EXACT
    from flowket_b200._lib import FlowketB200Error
    model, _, _, _ = make_pair('conv2d', (12, 12), 2, 32, seed=0)
    model.engine = FK_ENGINE_TC_EXACT
    with pytest.raises(FlowketB200Error):
        model.predict(random_sigma(4, (12, 12), seed=0))


# ---------------------------------------------------------------------------------------------------------------------
# sample-space SR behind the C ABI
# ---------------------------------------------------------------------------------------------------------------------
def _to_panels(Xrow, K, rld):
    """row-major [R, >= K] bf16 -> panel-major [ceil(K/64)][rld][64] (zero padded columns, junk in the padding rows)"""
    R = Xrow.shape[0]
    nkb = (K + 63) // 64
    Xp = torch.full((nkb, rld, 64), 3.0, dtype=torch.bfloat16, device=Xrow.device)
    pad = torch.zeros((R, nkb * 64), dtype=torch.bfloat16, device=Xrow.device)
    pad[:, :K] = Xrow[:, :K]
    Xp[:, :R, :] = pad.view(R, nkb, 64).permute(1, 0, 2)
    return Xp


def _gram(lib, Xp, R, K, scale, nblocks=1, stride=0):
    from flowket_b200 import _lib
    G = torch.empty((R, R), dtype=torch.float32, device=Xp.device)
    wsb = lib.fk_sr_gram_xxt_workspace_bytes(R)
    ws = torch.empty(wsb, dtype=torch.uint8, device=Xp.device)
    rld = Xp.shape[-2]
    _lib.check(lib.fk_sr_gram_xxt(Xp.data_ptr(), R, K, rld, nblocks, stride, scale, G.data_ptr(), R, ws.data_ptr(), wsb,
                                  _lib.stream_ptr()))
    return G


@pytest.mark.parametrize('R,K', [(256, 64), (512, 1000), (700, 4100), (1300, 33000), (2048, 70000), (130, 77)])
def test_sr_gram_xxt_centre_solve_xtw(R, K):
    """fk_sr_gram_xxt (hand-written cta_group::2 tcgen05 GEMM, bf16 operands / fp32 accumulate) vs the float64 product of
    the same bf16 numbers: measured <= 5.6e-5 of max |G| at K = 33 000 (truncating fp32 accumulation); exactly symmetric.
    Then fk_sr_centre_shift (fp64, measured 4e-15), fk_sr_solve (residual 4e-14) and fk_sr_xt_w (2e-7) on the result."""
    from flowket_b200 import _lib
    lib = _lib.require_cuda()
    dev = torch.device('cuda')
    torch.manual_seed(R * 7 + K)
    X = (torch.randn((R, K), device=dev) * (1 + torch.rand((R, 1), device=dev))).to(torch.bfloat16)
    Xp = _to_panels(X, K, R + (K % 5) * 8)            # rld >= R: padding rows must be ignored
    G = _gram(lib, Xp, R, K, 0.5)
    ref = 0.5 * (X.double() @ X.double().T)
    err = ((G.double() - ref).abs().max() / ref.abs().max()).item()
    print('gram_xxt', R, K, 'err', err)
    assert err < 1.7e-4 and bool((G == G.T).all())
    if R % 2 == 0:
        B = R // 2
        wsb = lib.fk_sr_centre_shift_workspace_bytes(R)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        S = torch.empty((R, R), dtype=torch.float64, device=dev)
        _lib.check(lib.fk_sr_centre_shift(G.data_ptr(), R, R, 1, 0.05, S.data_ptr(), ws.data_ptr(), wsb, _lib.stream_ptr()))
        C = torch.eye(B, dtype=torch.float64, device=dev) - 1.0 / B
        Cf = torch.block_diag(C, C)
        Sref = Cf @ G.double() @ Cf / B + 0.05 * torch.eye(R, dtype=torch.float64, device=dev)
        assert ((S - Sref).abs().max() / Sref.abs().max()).item() < 1e-13
        h = ctypes.c_void_p()
        _lib.check(lib.fk_sr_solver_create(ctypes.byref(h)))
        wsb = lib.fk_sr_solve_workspace_bytes(h, R)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        rhs = torch.randn(R, dtype=torch.float64, device=dev)
        x = rhs.clone()
        info = torch.full((1,), -7, dtype=torch.int32, device=dev)
        Sf = S.clone()
        _lib.check(lib.fk_sr_solve(h, Sf.data_ptr(), x.data_ptr(), R, info.data_ptr(), ws.data_ptr(), wsb, _lib.stream_ptr()))
        _lib.check(lib.fk_sr_solver_destroy(h))
        assert info.item() == 0
        assert ((S @ x - rhs).abs().max() / rhs.abs().max()).item() < 1e-11
    w = torch.randn(R, device=dev)
    o = torch.empty(K, dtype=torch.float32, device=dev)
    _lib.check(lib.fk_sr_xt_w(Xp.data_ptr(), R, K, Xp.shape[1], 1, 0, w.data_ptr(), o.data_ptr(), _lib.stream_ptr()))
    oref = X.double().T @ w.double()
    assert ((o.double() - oref).abs().max() / oref.abs().max()).item() < 1e-6


@pytest.mark.parametrize('n,cond', [(64, 1e2), (777, 1e4), (2048, 1.2e4), (1500, 2e5)])
def test_sr_solve_mixed_matches_fp64_solve(n, cond):
    """fk_sr_solve_mixed: fp32 Cholesky + fp64 iterative refinement against the fp64 solve on SPD matrices with the SR
    system's spectrum (eigenvalues diag_shift .. diag_shift * cond); the residual history it returns must contract"""
    from flowket_b200 import _lib
    lib = _lib.require_cuda()
    dev = torch.device('cuda')
    g = torch.Generator(device='cpu').manual_seed(n)
    Q, _ = torch.linalg.qr(torch.randn((n, n), generator=g, dtype=torch.float64))
    eig = 0.05 * torch.logspace(0, float(np.log10(cond)), n, dtype=torch.float64)
    S = ((Q * eig) @ Q.T)
    S = (0.5 * (S + S.T)).to(dev).contiguous()
    rhs = torch.randn(n, generator=g, dtype=torch.float64).to(dev)
    want = torch.linalg.solve(S, rhs)
    h = ctypes.c_void_p()
    _lib.check(lib.fk_sr_solver_create(ctypes.byref(h)))
    wsb = lib.fk_sr_solve_mixed_workspace_bytes(h, n)
    assert wsb > 0
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    refinements = 4
    x = rhs.clone()
    keep = S.clone()
    info = torch.full((1,), -7, dtype=torch.int32, device=dev)
    resid = torch.zeros(refinements + 2, dtype=torch.float64, device=dev)
    _lib.check(lib.fk_sr_solve_mixed(h, S.data_ptr(), x.data_ptr(), n, refinements, info.data_ptr(), resid.data_ptr(),
                                     ws.data_ptr(), wsb, _lib.stream_ptr()))
    _lib.check(lib.fk_sr_solver_destroy(h))
    assert info.item() == 0 and torch.equal(S, keep)                      # S is left intact
    hist = torch.sqrt(resid / resid[0]).cpu().numpy()
    print('mixed solve n', n, 'cond', cond, 'relative residuals', ['%.1e' % v for v in hist])
    assert hist[0] == 1.0 and hist[-1] < 1e-11 and hist[1] < 1e-2
    assert ((x - want).abs().max() / want.abs().max()).item() < 1e-9
    assert ((S @ x - rhs).norm() / rhs.norm()).item() < 1e-11           # the history's last entry is this residual


def test_sr_gram_xxt_row_blocks():
    """the sharded step: the rows arrive as per-rank blocks [Re ; Im] in separate buffers of one allocation; Gram, centring
    (per half of every block) and X^T w take (nblocks, block_stride)"""
    from flowket_b200 import _lib
    lib = _lib.require_cuda()
    dev = torch.device('cuda')
    torch.manual_seed(4)
    nbk, br, K = 3, 256, 5000
    R = nbk * br
    X = (torch.randn((R, K), device=dev) + 0.3).to(torch.bfloat16)
    nkb = (K + 63) // 64
    rld = br + 128
    buf = torch.full((nbk, nkb + 2, rld, 64), 5.0, dtype=torch.bfloat16, device=dev)      # 2 junk panels between the blocks
    for q in range(nbk):
        buf[q, :nkb] = _to_panels(X[q * br:(q + 1) * br], K, rld)
    stride = buf.stride(0) * 2
    G = _gram(lib, buf, R, K, 1.0, nbk, stride)
    ref = X.double() @ X.double().T
    assert ((G.double() - ref).abs().max() / ref.abs().max()).item() < 6e-5       # measured 1.7e-5
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
    assert ((S - Sref).abs().max() / Sref.abs().max()).item() < 1e-12
    w = torch.randn(R, device=dev)
    o = torch.empty(K, dtype=torch.float32, device=dev)
    _lib.check(lib.fk_sr_xt_w(buf.data_ptr(), R, K, rld, nbk, stride, w.data_ptr(), o.data_ptr(), _lib.stream_ptr()))
    oref = X.double().T @ w.double()
    assert ((o.double() - oref).abs().max() / oref.abs().max()).item() < 1e-6


@pytest.mark.parametrize('shape,depth,wn,B', [((6, 6), 3, True, 96), ((4, 5), 4, False, 64), ((10, 10), 4, True, 128)])
def test_jacobian_rows_bf16_and_device_sr_pipeline(shape, depth, wn, B):
    """fk_jacobian_rows_tc: per-sample Jacobian rows in bf16, panel-major, weight-norm transform fused, vs the fp32 engine's
    rows (those are pinned to the oracle in test_gpu_parity): measured 1.2e-3 of max |O| / 1.3e-3 Frobenius against the
    fp16 tensor-core rows, + the fp16 engine's own 1.3e-2; then the whole device pipeline (StochasticReconfiguration with the
    defaults) vs the fp64 solve of the oracle's SR system on oracle Jacobians."""
    from flowket_b200 import _lib, FK_ENGINE_TC, FK_ENGINE_FP32
    from flowket_b200.optimizers import StochasticReconfiguration
    lib = _lib.require_cuda()
    model, _, spec, params = _perturbed(shape, depth, wn, seed=2)
    model.engine = FK_ENGINE_TC
    net = model.machine.device_net()
    sigma = random_sigma(B, shape, seed=0)
    sg = net.to_sigma(sigma)
    O_re, O_im = net.grad_per_sample(sg, imag=True, engine=FK_ENGINE_FP32)
    P = net.num_params
    nkb = (P + 63) // 64
    rld = 2 * B + 8
    Xp = torch.full((nkb, rld, 64), 9.0, dtype=torch.bfloat16, device=sg.device)
    wsb = lib.fk_jacobian_rows_tc_workspace_bytes(net.handle, B)
    ws = torch.empty(wsb, dtype=torch.uint8, device=sg.device)
    _lib.check(lib.fk_jacobian_rows_tc(net.handle, sg.data_ptr(), B, Xp.data_ptr(), rld, 0, B, ws.data_ptr(), wsb,
                                       _lib.stream_ptr()))
    rows = Xp.permute(1, 0, 2).reshape(rld, nkb * 64).float()
    want = torch.cat([O_re, O_im])
    got = rows[:2 * B, :P]
    fro = ((got - want).norm() / want.norm()).item()
    print('bf16 Jacobian rows', shape, depth, wn, 'Frobenius vs fp32 engine', fro)
    assert fro < 2e-2
    assert bool((rows[:2 * B, P:] == 0).all()) and bool((rows[2 * B:] == 9.0).all())        # zero K padding, untouched row padding
    # device pipeline vs the oracle's system
    rng = np.random.default_rng(1)
    eloc = rng.normal(size=B) * 2 - 10 + 1j * rng.normal(size=B)
    o_re = nets.per_sample_gradients(spec, params, sigma, 'real').numpy()
    o_im = nets.per_sample_gradients(spec, params, sigma, 'imag').numpy()
    # the oracle's update in its sample-space form (fp64 numpy; the P x P form of osr.real_sr_system would need P^2 doubles --
    # 10^5 parameters here -- and is pinned against this identity on a small machine in test_gpu_vmc.py and
    # tests/test_host_logic.py): delta = X^T (X X^T / B + lambda I)^-1 e' / B, X = [Re Obar ; Im Obar]
    Xo = np.concatenate([o_re - o_re.mean(0), o_im - o_im.mean(0)])
    ec = eloc - eloc.mean()
    want_delta = Xo.T @ np.linalg.solve(Xo @ Xo.T / B + 0.05 * np.eye(2 * B), np.concatenate([ec.real, ec.imag]) / B)
    sr = StochasticReconfiguration(model, diag_shift=0.05, sample_space=True)
    got_delta = sr.compute_update(sg, eloc).cpu().numpy()
    assert {'jacobian', 'gram', 'cholesky'} <= set(sr.last_timings_ms), 'the device pipeline did not run'
    derr = np.linalg.norm(got_delta - want_delta) / np.linalg.norm(want_delta)
    cos = float(got_delta @ want_delta / np.linalg.norm(got_delta) / np.linalg.norm(want_delta))
    # the same TC Jacobian rows through the torch route (fp32 rows, fp32 Gram, fp64 solve): isolates what the bf16 storage,
    # the hand-written Gram and the mixed-precision solve add on top of the fp16 Jacobian engine
    ref = StochasticReconfiguration(model, diag_shift=0.05, sample_space=True, device_pipeline=False, gram_dtype='fp32')
    ref_delta = ref.compute_update(sg, eloc).cpu().numpy()
    perr = np.linalg.norm(got_delta - ref_delta) / np.linalg.norm(ref_delta)
    print('device SR delta vs oracle', shape, depth, wn, derr, 'cos', cos, '; vs the torch route on the same fp16-engine rows', perr)
    # measured on a B200: 5.6e-2 / 9.9e-2 / 6.5e-2 vs the oracle (the fp16 Jacobian engine's 1e-2 row error amplified by the
    # solve at diag_shift 0.05; cosine 0.998 / 0.995 / 0.998), 1.6e-2 .. 2.4e-2 vs the torch route on the same rows (bf16 storage)
    assert derr < 0.2 and cos > 0.99
    assert perr < 5e-2
    # fp64 Cholesky instead of the mixed-precision solve: same update to the refinement tolerance
    sr64 = StochasticReconfiguration(model, diag_shift=0.05, sample_space=True, solver='fp64')
    d64 = sr64.compute_update(sg, eloc).cpu().numpy()
    assert np.linalg.norm(d64 - got_delta) / np.linalg.norm(d64) < 1e-5
    # a refinement that does not reach its tolerance is repeated in fp64 automatically (forced here with a tolerance of zero)
    sr._pipeline.refinement_tol = 0.0
    d_fb = sr.compute_update(sg, eloc).cpu().numpy()
    assert sr._pipeline.fp64_fallbacks == 1
    assert np.linalg.norm(d_fb - d64) / np.linalg.norm(d64) < 1e-6


# ---------------------------------------------------------------------------------------------------------------------
# exact enumeration kernels (fk_exact.cu)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('opkind,opkw,shape', [
    ('heisenberg', dict(pbc=False), (3, 4)), ('heisenberg', dict(pbc=True), (12,)), ('ising', dict(pbc=False, h=3.0), (4, 4)),
    ('ising', dict(pbc=True, h=0.7), (10,)), ('j1j2', dict(pbc=False, j2=0.5), (4, 3)), ('j1j2', dict(pbc=True, j2=0.3), (4, 4))])
def test_exact_tables_match_find_conn(opkind, opkw, shape):
    """fk_exact_states / fk_exact_index round trip and fk_exact_conn_table == indices of the oracle's find_conn
    configurations, matrix elements bit-equal (integer / index work: exact)"""
    from flowket_b200 import _lib
    from tests.test_gpu_parity import _product_operator
    lib = _lib.require_cuda()
    N = int(np.prod(shape))
    S = 2 ** N
    first, n = (S // 4, S // 2) if S > 1024 else (0, S)
    sigma = torch.empty((n, N), dtype=torch.int8, device='cuda')
    _lib.check(lib.fk_exact_states(first, n, N, sigma.data_ptr(), _lib.stream_ptr()))
    idx = np.arange(first, first + n)
    want_states = (2 * ((idx[:, None] >> np.arange(N)) & 1) - 1).astype(np.int8)
    assert np.array_equal(sigma.cpu().numpy(), want_states)
    back = torch.empty(n, dtype=torch.int64, device='cuda')
    _lib.check(lib.fk_exact_index(sigma.data_ptr(), n, N, back.data_ptr(), _lib.stream_ptr()))
    assert np.array_equal(back.cpu().numpy(), idx)
    op = _product_operator(opkind, shape, opkw)
    desc = op.device_desc()
    C = int(desc.max_conn)
    index = torch.empty((C, n), dtype=torch.int64, device='cuda')
    mel = torch.empty((C, n), dtype=torch.float64, device='cuda')
    _lib.check(lib.fk_exact_conn_table(ctypes.byref(desc), first, n, index.data_ptr(), mel.data_ptr(), _lib.stream_ptr()))
    oop = oops.OracleOperator(opkind, shape, **opkw)
    conn, omel, use = oop.find_conn(want_states.reshape((n,) + shape).astype(np.float64))
    bits = (conn.reshape(conn.shape[0], n, N) == 1).astype(np.int64)
    want_index = (bits << np.arange(N)).sum(axis=-1)
    got_index, got_mel = index.cpu().numpy(), mel.cpu().numpy()
    assert got_mel.shape[0] >= omel.shape[0]
    assert np.array_equal(got_mel[:omel.shape[0]], np.real(omel))
    assert not got_mel[omel.shape[0]:].any()
    live = np.real(omel) != 0
    assert np.array_equal(got_index[:omel.shape[0]][live], want_index[live])
    assert np.array_equal(got_index[0], idx)                       # slot 0 is the state itself


def test_exact_local_energy_kernel_matches_float64_expression():
    from flowket_b200 import _lib
    from flowket_b200.operators import Heisenberg
    lib = _lib.require_cuda()
    shape = (3, 4)
    N, S = 12, 4096
    op = Heisenberg(hilbert_state_shape=list(shape), pbc=False)
    desc = op.device_desc()
    C = int(desc.max_conn)
    first, n = 1024, 2048
    index = torch.empty((C, n), dtype=torch.int64, device='cuda')
    mel = torch.empty((C, n), dtype=torch.float64, device='cuda')
    _lib.check(lib.fk_exact_conn_table(ctypes.byref(desc), first, n, index.data_ptr(), mel.data_ptr(), _lib.stream_ptr()))
    g = torch.Generator().manual_seed(3)
    table = torch.complex(torch.randn(S, generator=g, dtype=torch.float64) * 2 - 5,
                          torch.rand(S, generator=g, dtype=torch.float64) * 6.0).cuda()
    log_norm = float(torch.logsumexp(2 * table.real, 0))
    weighted = torch.empty(n, dtype=torch.complex128, device='cuda')
    naive = torch.empty(n, dtype=torch.complex128, device='cuda')
    _lib.check(lib.fk_exact_local_energy(table.data_ptr(), index.data_ptr(), mel.data_ptr(), C, n, log_norm,
                                         weighted.data_ptr(), naive.data_ptr(), _lib.stream_ptr()))
    gathered = table[index]
    own = gathered[0]
    want_w = (mel * torch.exp(gathered.conj() + own - log_norm)).sum(dim=0)
    want_n = (mel * torch.exp(gathered - own)).sum(dim=0)
    assert ((weighted - want_w).abs().max() / want_w.abs().max()).item() < 1e-13
    assert ((naive - want_n).abs().max() / want_n.abs().max()).item() < 1e-13
    only_w = torch.empty(n, dtype=torch.complex128, device='cuda')
    _lib.check(lib.fk_exact_local_energy(table.data_ptr(), index.data_ptr(), mel.data_ptr(), C, n, log_norm,
                                         only_w.data_ptr(), None, _lib.stream_ptr()))
    assert torch.equal(only_w, weighted)


def test_exact_variational_tables_stay_on_the_device():
    """the product route keeps psi, probabilities, connection tables and coefficients as CUDA tensors and
    agrees with the host route of the same class (oracle operator + host find_conn protocol) to fp64 round-off"""
    from flowket_b200.optimization import ExactVariational
    from flowket_b200.operators import Heisenberg
    shape = (3, 4)
    model, _, spec, params = make_pair('conv2d', shape, 2, 16, seed=4)
    ev = ExactVariational(model, Heisenberg(hilbert_state_shape=list(shape), pbc=False), 2 ** 10)
    ev.machine_updated()
    obs = ev.energy_observable
    for t in (ev._t['wave_function'], ev._t['probs'], ev._t['energy_grad_coefficients'], obs._t['index'], obs._t['mel'],
              obs._t['energies']):
        assert t.is_cuda
    host = ExactVariational(model, oops.OracleOperator('heisenberg', shape, pbc=False), 2 ** 10)
    host.machine_updated()
    assert abs(obs.current_energy - host.energy_observable.current_energy) < 1e-10 * abs(obs.current_energy)
    assert np.allclose(ev.energy_grad_coefficients, host.energy_grad_coefficients, rtol=1e-9, atol=1e-14)
    assert obs.current_local_energy_variance == pytest.approx(host.energy_observable.current_local_energy_variance, rel=1e-9)
