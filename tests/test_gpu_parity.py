"""GPU parity tests: CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs.
Tolerances: integer / index work bit-exact; fp32 engine 1e-5 relative (north_star); stated per test."""
import numpy as np
import pytest
import torch

from oracle import exact as oexact
from oracle import local_energy as oeloc
from oracle import nets, operators as oops
from oracle import sampler as osampler
from tests.helpers import make_pair, random_sigma

pytestmark = pytest.mark.gpu

NET_CASES = [
    ('conv2d', (4, 4), 4, 16, {}),
    ('conv2d', (5, 6), 3, 32, {}),
    ('conv2d', (4, 3), 2, 8, {'weights_normalization': False}),
    ('conv2d', (10, 10), 5, 32, {'weights_normalization': False}),
    ('conv1d', (16,), 7, 16, {}),
    ('conv1d', (20,), 8, 64, {'weights_normalization': False, 'max_dilation_rate': 4}),
    ('conv1d', (12,), 5, 32, {'max_dilation_rate': 4, 'add_skip_connections': True}),
    ('cconv1d', (12,), 4, 16, {'max_dilation_rate': 4}),
    ('cconv1d', (20,), 5, 32, {'max_dilation_rate': 4}),
    # channel counts / lattice sizes whose products are not multiples of 4 floats: every carved workspace buffer must
    # still start 16-byte aligned (round-1 fault: cudaErrorMisalignedAddress in the float4 head kernels at C = 6)
    ('conv1d', (11,), 4, 6, {'weights_normalization': False}),
    ('conv1d', (13,), 5, 10, {}),
    ('conv2d', (3, 5), 3, 6, {'weights_normalization': False}),
    ('conv2d', (5, 3), 2, 12, {}),
    ('cconv1d', (9,), 3, 6, {}),
]


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize('kind,shape,depth,channels,kw', NET_CASES)
def test_log_psi_matches_oracle(kind, shape, depth, channels, kw):
    model, cond_model, spec, params = make_pair(kind, shape, depth, channels, seed=3, **kw)
    sigma = random_sigma(67, shape, seed=1)       # ragged size: exercises partial tiles
    got = model.predict(sigma)
    want = nets.log_psi_numpy(spec, params, sigma)
    assert got.shape == (67, 1) and got.dtype == np.complex64
    assert _rel(got, want) < 1e-5
    got_c = cond_model.predict(sigma)
    want_c = nets.conditional_log_probs(spec, params, sigma).detach().numpy()
    assert got_c.shape == (67,) + tuple(shape) + (2,)
    assert np.abs(got_c - want_c).max() < 1e-5 * max(1.0, np.abs(want_c).max())


def test_log_psi_empty_and_single():
    model, _, spec, params = make_pair('conv2d', (4, 4), 3, 8, seed=0)
    assert model.predict(np.zeros((0, 4, 4), np.int8)).shape == (0, 1)
    s = random_sigma(1, (4, 4))
    assert _rel(model.predict(s), nets.log_psi_numpy(spec, params, s)) < 1e-5


def test_normalisation_over_all_states():
    """tests/test_autoregressive.py:18-30 of the reference: sum_sigma |psi|^2 = 1 (rel 1e-5)."""
    for kind, shape, depth, ch in [('conv1d', (16,), 7, 16), ('conv2d', (4, 4), 4, 16)]:
        model, _, _, _ = make_pair(kind, shape, depth, ch, seed=5)
        states = oexact.all_states(16).reshape((-1,) + shape)
        lp = model.predict(states)[:, 0]
        assert np.exp(2.0 * lp.real.astype(np.float64)).sum() == pytest.approx(1.0, rel=1e-5)


FIND_CONN_CASES = {
    'heis_2d_obc': ('Heisenberg', dict(hilbert_state_shape=[4, 5], pbc=False)),
    'heis_2d_pbc': ('Heisenberg', dict(hilbert_state_shape=[4, 4], pbc=True)),
    'heis_2d_obc_norot': ('Heisenberg', dict(hilbert_state_shape=[3, 4], pbc=False, unitary_rotation=False)),
    'heis_1d_pbc': ('Heisenberg', dict(hilbert_state_shape=[7], pbc=True)),
    'heis_1d_obc': ('Heisenberg', dict(hilbert_state_shape=[6], pbc=False)),
    'heis_2d_10x10_obc': ('Heisenberg', dict(hilbert_state_shape=[10, 10], pbc=False)),
    'ising_2d_obc': ('Ising', dict(hilbert_state_shape=[4, 4], pbc=False, h=3.0)),
    'ising_2d_pbc': ('Ising', dict(hilbert_state_shape=[3, 5], pbc=True, h=0.7, j=1.3)),
    'ising_1d_obc': ('Ising', dict(hilbert_state_shape=[9], pbc=False, h=3.0)),
    'ising_1d_pbc': ('Ising', dict(hilbert_state_shape=[8], pbc=True, h=2.0)),
}


@pytest.mark.parametrize('name', sorted(FIND_CONN_CASES))
def test_find_conn_bit_exact_vs_reference_golden(golden, name):
    import flowket_b200.operators as ops
    cls, kw = FIND_CONN_CASES[name]
    op = getattr(ops, cls)(**kw)
    conn, mel, use = op.find_conn(golden[name + '/sigma'])
    assert op.max_number_of_local_connections == int(golden[name + '/max_conn'])
    assert conn.dtype == np.float64 and use.dtype == bool
    assert np.array_equal(conn.astype(np.int8), golden[name + '/conn'])
    assert np.array_equal(use, golden[name + '/use'])
    if name == 'ising_2d_pbc':   # non-integer j: fp32 accumulation order of the diagonal is not pinned
        assert np.allclose(mel, golden[name + '/mel'], rtol=1e-6)
    else:
        assert np.array_equal(mel, golden[name + '/mel'])


def test_find_conn_j1j2_matches_oracle():
    import flowket_b200.operators as ops
    for shape, pbc in [((4, 4), False), ((6, 6), False), ((4, 4), True)]:
        op = ops.J1J2(shape, j2=0.5, pbc=pbc)
        sigma = random_sigma(9, shape, seed=2)
        conn, mel, use = op.find_conn(sigma)
        oconn, omel, ouse = oops.j1j2_find_conn(sigma, shape, j2=0.5, pbc=pbc)
        assert mel.dtype == np.complex128
        assert np.array_equal(conn, oconn) and np.array_equal(use, ouse) and np.array_equal(mel, omel)


ELOC_CASES = [
    ('conv2d', (4, 4), 3, 16, {}, 'heisenberg', dict(pbc=False)),
    ('conv2d', (4, 4), 3, 16, {}, 'ising', dict(pbc=False, h=3.0)),
    ('conv2d', (4, 4), 3, 16, {}, 'j1j2', dict(pbc=False, j2=0.5)),
    ('conv2d', (6, 5), 4, 32, {}, 'heisenberg', dict(pbc=True)),
    ('conv1d', (20,), 6, 32, {'max_dilation_rate': 4}, 'heisenberg', dict(pbc=True)),
    ('cconv1d', (12,), 4, 16, {'max_dilation_rate': 4}, 'heisenberg', dict(pbc=True)),
]


def _product_operator(kind, shape, kw):
    import flowket_b200.operators as ops
    if kind == 'heisenberg':
        return ops.Heisenberg(hilbert_state_shape=list(shape), **kw)
    if kind == 'ising':
        return ops.Ising(hilbert_state_shape=list(shape), **kw)
    return ops.J1J2(tuple(shape), **kw)


@pytest.mark.parametrize('kind,shape,depth,channels,kw,opkind,opkw', ELOC_CASES)
def test_local_energy_matches_oracle(kind, shape, depth, channels, kw, opkind, opkw):
    from flowket_b200.observables.monte_carlo import Observable
    model, _, spec, params = make_pair(kind, shape, depth, channels, seed=11, **kw)
    sigma = random_sigma(37, shape, seed=4)
    obs = Observable(_product_operator(opkind, shape, opkw))
    got = obs.local_values(model, sigma)
    oop = oops.OracleOperator(opkind, shape, **opkw)
    psi32 = [p.to(torch.float32) for p in params]
    want = oeloc.local_values(oop, lambda c: nets.log_psi_numpy(spec, params, c), sigma.astype(np.float64))
    assert got.dtype == np.complex128 and got.shape == (37,)
    assert _rel(got, want) < 1e-5
    # statistics vector used by the multi-GPU allreduce
    st = obs.last_stats.cpu().numpy()
    assert st[3] == 37 and st[0] == pytest.approx(got.real.sum(), rel=1e-12)
    assert st[2] == pytest.approx((got.real ** 2).sum(), rel=1e-12)
    # the generic protocol route (callable psi) gives the same numbers through fk_find_conn
    got2 = obs.local_values(lambda c: model.predict(c), sigma)
    assert _rel(got2, want) < 1e-5


def test_local_energy_constant_on_exact_ground_state():
    """psi = ED ground state => E_loc == E0 for every sample (uses the generic route + device find_conn)."""
    from flowket_b200.observables.monte_carlo import Observable
    from flowket_b200.exact.utils import vector_to_machine
    shape = (3, 4)
    oop = oops.OracleOperator('heisenberg', shape, pbc=False)
    e0, vec = oexact.ground_state(oop, shape)
    vec = np.abs(vec)
    keep = vec > 1e-8
    logv = np.where(keep, np.log(np.maximum(vec, 1e-300)), -80.0).astype(np.complex128)
    sigma = oexact.all_states(12)[keep][:64].reshape((-1,) + shape)
    obs = Observable(_product_operator('heisenberg', shape, dict(pbc=False)))
    lv = obs.local_values(vector_to_machine(logv), sigma)
    assert np.allclose(lv, e0, atol=1e-6)


@pytest.mark.parametrize('shape,depth,channels,B', [((4, 5), 3, 32, 50), ((6, 6), 4, 32, 24), ((10, 10), 3, 32, 40)])
def test_fast_sampler_bit_exact_given_uniforms(shape, depth, channels, B):
    """Explicit-uniform rule (deepar/samplers/autoregressive.py:37-44): identical draws -> identical spins.
    A site may legitimately differ only when |p0 - u| is below fp32 summation-order noise; none is allowed to
    differ by more than that margin, and the fixed seeds below produce no such tie."""
    from flowket_b200.samplers import FastAutoregressiveSampler, AutoregressiveSampler
    model, cond_model, spec, params = make_pair('conv2d', shape, depth, channels, seed=21)
    u = np.random.RandomState(5).random_sample((B,) + shape)
    want, p0 = osampler.sample_with_uniforms(spec, params, u)
    fast = FastAutoregressiveSampler(cond_model, B)
    got = fast.next_device(uniforms=u, return_p0=True).cpu().numpy()
    got_p0 = fast.last_p0.cpu().numpy().reshape((B,) + shape)
    naive = AutoregressiveSampler(cond_model, B)
    got_naive = naive.next_device(uniforms=u).cpu().numpy()
    assert got.dtype == np.int8 and got.shape == (B,) + shape
    for g in (got, got_naive):
        bad = np.argwhere(g != want)
        for idx in bad:   # first differing site of a sample must be a numerical tie
            first = tuple(bad[bad[:, 0] == idx[0]][0])
            assert abs(p0[first] - u[first]) < 1e-5, (first, p0[first], u[first])
        assert len(bad) == 0
    same = (got == want).all(axis=tuple(range(1, got.ndim)))
    assert np.abs(got_p0[same] - p0[same]).max() < 2e-6


def test_sampler_1d_and_complex_given_uniforms():
    from flowket_b200.samplers import FastAutoregressiveSampler
    for kind, shape, depth, ch, kw in [('conv1d', (14,), 5, 16, {'max_dilation_rate': 4}),
                                       ('cconv1d', (10,), 3, 8, {})]:
        model, cond_model, spec, params = make_pair(kind, shape, depth, ch, seed=2, **kw)
        u = np.random.RandomState(9).random_sample((33,) + shape)
        want, _ = osampler.sample_with_uniforms(spec, params, u)
        got = FastAutoregressiveSampler(cond_model, 33).next_device(uniforms=u).cpu().numpy()
        assert np.array_equal(got, want)


@pytest.mark.parametrize('kind,shape,depth,ch,kw', [
    ('conv1d', (20,), 8, 64, {'max_dilation_rate': 4, 'weights_normalization': False}),     # BASELINE configs[1]
    ('conv1d', (16,), 7, 32, {}),
    ('cconv1d', (36,), 5, 16, {}),                                                          # configs[3]: 6x6 flattened
])
def test_incremental_1d_sampler_equals_the_n_forward_sampler(kind, shape, depth, ch, kw):
    """fk_sample on the 1-D machines is the cached incremental kernel (one evaluation per (layer, position));
    AutoregressiveSampler is the N-forward schedule (autoregressive.py:29-48).  Same uniforms -> same spins and the
    same conditional probabilities (both follow conv_kernel's summation order for these shapes up to fp32 noise)."""
    from flowket_b200.samplers import FastAutoregressiveSampler, AutoregressiveSampler
    model, cond_model, spec, params = make_pair(kind, shape, depth, ch, seed=4, **kw)
    B = 77
    u = np.random.RandomState(3).random_sample((B,) + shape)
    fast = FastAutoregressiveSampler(cond_model, B)
    got = fast.next_device(uniforms=u, return_p0=True).cpu().numpy()
    p0_fast = fast.last_p0.cpu().numpy().reshape((B,) + shape)
    naive = AutoregressiveSampler(cond_model, B)
    ref = naive.next_device(uniforms=u, return_p0=True).cpu().numpy()
    p0_naive = naive.last_p0.cpu().numpy().reshape((B,) + shape)
    bad = np.argwhere(got != ref)
    for idx in bad:   # a differing site must be a numerical tie at its first occurrence in the sample
        first = tuple(bad[bad[:, 0] == idx[0]][0])
        assert abs(p0_naive[first] - u[first]) < 1e-5, (first, p0_naive[first], u[first])
    same = (got == ref).all(axis=1)
    assert same.mean() > 0.95
    assert np.abs(p0_fast[same] - p0_naive[same]).max() < 5e-6
    want, _ = osampler.sample_with_uniforms(spec, params, u)
    assert (got == want).all(axis=1).mean() > 0.95


def test_sampler_philox_is_shard_invariant_and_distribution():
    """Philox counters are keyed by the global sample index: two half-batches == one full batch; and the
    histogram of the samples matches |psi|^2 (L1 test of tests/test_samplers.py:49-82, z <= sqrt(n))."""
    from flowket_b200.samplers import FastAutoregressiveSampler
    shape = (4, 3)
    model, cond_model, spec, params = make_pair('conv2d', shape, 2, 32, seed=8)
    full = FastAutoregressiveSampler(cond_model, 64, seed=77).next_device().cpu().numpy()
    lo = FastAutoregressiveSampler(cond_model, 32, seed=77, sample_offset=0).next_device().cpu().numpy()
    hi = FastAutoregressiveSampler(cond_model, 32, seed=77, sample_offset=32).next_device().cpu().numpy()
    assert np.array_equal(full, np.concatenate([lo, hi]))
    n = 2 ** 16
    s = FastAutoregressiveSampler(cond_model, n, seed=3).next_device().cpu().numpy().reshape(n, -1)
    idx = oexact.states_to_index(s)
    states = oexact.all_states(12).reshape((-1,) + shape)
    probs = np.exp(2.0 * nets.log_psi_numpy(spec, params, states)[:, 0].real)
    counts = np.bincount(idx, minlength=4096)
    # two-sample-free variant of the closeness statistic: chi-like sum against exact probabilities
    expected = n * probs
    z = (((counts - expected) ** 2 - counts) / np.maximum(expected, 1e-12))[expected > 1e-3].sum()
    assert z <= 3.0 * np.sqrt(n)


GRAD_CASES = [
    ('conv2d', (4, 4), 3, 16, {}),
    ('conv2d', (5, 4), 4, 32, {'weights_normalization': False}),
    ('conv1d', (12,), 5, 16, {'max_dilation_rate': 4, 'add_skip_connections': True}),
    ('conv1d', (10,), 4, 32, {'weights_normalization': False}),
    ('cconv1d', (10,), 3, 8, {'max_dilation_rate': 2}),
    # odd channel counts / odd site counts / odd batch: workspace alignment (see NET_CASES)
    ('conv1d', (11,), 4, 6, {'weights_normalization': False}),
    ('conv1d', (13,), 3, 10, {}),
    ('conv2d', (3, 5), 3, 6, {'weights_normalization': False}),
    ('conv2d', (5, 3), 2, 12, {}),
    ('cconv1d', (9,), 3, 6, {}),
]


@pytest.mark.parametrize('kind,shape,depth,channels,kw', GRAD_CASES)
def test_gradients_match_oracle(kind, shape, depth, channels, kw):
    model, _, spec, params = make_pair(kind, shape, depth, channels, seed=13, **kw)
    net = model.machine.device_net()
    B = 19
    sigma = random_sigma(B, shape, seed=6)
    rng = np.random.RandomState(3)
    y = (rng.normal(size=B) + 1j * rng.normal(size=B)).astype(np.complex64)
    got = net.grad_weighted(net.to_sigma(sigma), torch.from_numpy(y)).cpu().numpy()
    want = nets.weighted_gradient(spec, params, sigma, y.astype(np.complex128)).numpy()
    scale = np.abs(want).max()
    assert np.abs(got - want).max() < 1e-5 * scale * 10   # fp32 accumulation over B*sites terms
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-5
    O_re, O_im = net.grad_per_sample(net.to_sigma(sigma[:5]), imag=True)
    want_re = nets.per_sample_gradients(spec, params, sigma[:5], 'real').numpy()
    want_im = nets.per_sample_gradients(spec, params, sigma[:5], 'imag').numpy()
    assert np.linalg.norm(O_re.cpu().numpy() - want_re) / np.linalg.norm(want_re) < 1e-5
    assert np.linalg.norm(O_im.cpu().numpy() - want_im) / max(np.linalg.norm(want_im), 1e-30) < 1e-5


def test_samplers_and_local_energy_with_unaligned_channel_counts():
    """naive sampler, incremental 1-D sampler and E_loc on a 6-channel / 11-site machine (alignment of the carved buffers)"""
    from flowket_b200.samplers import FastAutoregressiveSampler, AutoregressiveSampler
    from flowket_b200.observables.monte_carlo import Observable
    model, cond_model, spec, params = make_pair('conv1d', (11,), 4, 6, seed=4, weights_normalization=False)
    B = 13
    u = np.random.RandomState(3).random_sample((B, 11))
    want, _ = osampler.sample_with_uniforms(spec, params, u)
    assert np.array_equal(FastAutoregressiveSampler(cond_model, B).next_device(uniforms=u).cpu().numpy(), want)
    assert np.array_equal(AutoregressiveSampler(cond_model, B).next_device(uniforms=u).cpu().numpy(), want)
    obs = Observable(_product_operator('heisenberg', (11,), dict(pbc=True)))
    got = obs.local_values(model, want)
    ref = oeloc.local_values(oops.OracleOperator('heisenberg', (11,), pbc=True),
                             lambda c: nets.log_psi_numpy(spec, params, c), want.astype(np.float64))
    assert _rel(got, ref) < 1e-5
