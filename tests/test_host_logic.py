"""CPU tier: host-side mirror of the reference interface (no device calls)."""
import numpy as np
import pytest
import torch

from oracle import nets, operators as oops


def test_parameter_layout_matches_oracle_and_survey():
    from flowket_b200 import Input, Model
    from flowket_b200.machines import ConvNetAutoregressive2D, SimpleConvNetAutoregressive1D, \
        ComplexValuesSimpleConvNetAutoregressive1D
    m = ConvNetAutoregressive2D(Input(shape=(10, 10)), depth=20, num_of_channels=32)
    assert m.count_params() == 849156 + 4864
    m2 = ConvNetAutoregressive2D(Input(shape=(10, 10)), depth=20, num_of_channels=32, weights_normalization=False)
    assert m2.count_params() == 849156
    spec = nets.Conv2DSpec(10, 10, 20, 32)
    assert [tuple(s) for _, s, _ in m.weight_specs()] == [tuple(p.shape) for p in nets.init_params(spec)]
    # Keras layer names of the pretrained files: weight_normalization[_k]/{kernel,bias,g}:0, k = 0..89, conv2d_90
    m10 = ConvNetAutoregressive2D(Input(shape=(12, 12)), depth=10, num_of_channels=32)
    names = [n for n, _, _ in m10.weight_specs()]
    assert names[0] == 'weight_normalization/kernel:0' and 'weight_normalization_89/g:0' in names
    assert names[-2:] == ['conv2d_90/kernel:0', 'conv2d_90/bias:0']
    m1 = SimpleConvNetAutoregressive1D(Input(shape=(20,)), depth=8, num_of_channels=64, max_dilation_rate=4,
                                       weights_normalization=False)
    assert m1.count_params() == nets.num_params(nets.Conv1DSpec(20, 8, 64, max_dilation_rate=4, weights_normalization=False))
    mc = ComplexValuesSimpleConvNetAutoregressive1D(Input(shape=(36,)), depth=8, num_of_channels=32, max_dilation_rate=4)
    assert mc.count_params() == 2 * 18818                           # SURVEY appendix B, cfg 4 (complex count)
    model = Model(inputs=m.keras_input_layer, outputs=m.predictions)
    assert model.input_shape == (None, 10, 10) and model.output_shape == (None, 1)


def test_initial_weights_weight_norm_identity():
    """CopyNormaInitializer: with g0 = log|v| (exp norm) the effective kernel equals v at init."""
    from flowket_b200 import Input
    from flowket_b200.machines import ConvNetAutoregressive2D
    m = ConvNetAutoregressive2D(Input(shape=(4, 4)), depth=3, num_of_channels=8, seed=3)
    w = m.initial_weights(3)
    specs = m.weight_specs()
    v, g = w[0], w[2]
    assert specs[2][0].endswith('/g:0')
    eff = v / np.sqrt((v.reshape(-1, v.shape[-1]) ** 2).sum(0)) * np.exp(g)
    assert np.allclose(eff, v, rtol=1e-5)
    m.set_weights(w)
    back = m.get_weights()
    assert all(np.array_equal(a, b) for a, b in zip(w, back))
    with pytest.raises(ValueError):
        m.set_weights(w[:-1])


def test_operator_term_tables_match_oracle_on_host():
    """the device term tables (built on the host) reproduce the oracle's find_conn when interpreted in numpy"""
    from flowket_b200 import _lib
    import flowket_b200.operators as ops
    rng = np.random.default_rng(0)
    cases = [(ops.Heisenberg(hilbert_state_shape=[4, 5], pbc=False), 'heisenberg', (4, 5), dict(pbc=False)),
             (ops.Heisenberg(hilbert_state_shape=[4, 4], pbc=True), 'heisenberg', (4, 4), dict(pbc=True)),
             (ops.Heisenberg(hilbert_state_shape=[7], pbc=True), 'heisenberg', (7,), dict(pbc=True)),
             (ops.Ising(hilbert_state_shape=[3, 5], pbc=True, h=0.5), 'ising', (3, 5), dict(pbc=True, h=0.5)),
             (ops.Ising(hilbert_state_shape=[9], pbc=False, h=3.0), 'ising', (9,), dict(pbc=False, h=3.0)),
             (ops.J1J2((4, 4), j2=0.5), 'j1j2', (4, 4), dict(j2=0.5)),
             # degenerate lattices (pinned against the reference in tests/golden/reference_numpy_half_edge.npz)
             (ops.Heisenberg(hilbert_state_shape=[2, 2], pbc=True), 'heisenberg', (2, 2), dict(pbc=True)),
             (ops.Heisenberg(hilbert_state_shape=[1, 6], pbc=False), 'heisenberg', (1, 6), dict(pbc=False)),
             (ops.Heisenberg(hilbert_state_shape=[6, 1], pbc=True), 'heisenberg', (6, 1), dict(pbc=True)),
             (ops.Heisenberg(hilbert_state_shape=[2], pbc=True), 'heisenberg', (2,), dict(pbc=True)),
             (ops.Heisenberg(hilbert_state_shape=[2, 5], pbc=True), 'heisenberg', (2, 5), dict(pbc=True)),
             (ops.Ising(hilbert_state_shape=[2, 3], pbc=True, h=0.5), 'ising', (2, 3), dict(pbc=True, h=0.5)),
             (ops.Ising(hilbert_state_shape=[1, 4], pbc=True, h=1.0), 'ising', (1, 4), dict(pbc=True, h=1.0)),
             (ops.Ising(hilbert_state_shape=[2], pbc=True, h=1.0), 'ising', (2,), dict(pbc=True, h=1.0)),
             (ops.J1J2((2, 3), j2=0.3), 'j1j2', (2, 3), dict(j2=0.3)),
             (ops.J1J2((3, 3), j2=0.5, pbc=True), 'j1j2', (3, 3), dict(j2=0.5, pbc=True))]
    for op, kind, shape, kw in cases:
        terms, kind_id, compact, _ = op.terms()
        sigma = rng.choice([-1, 1], size=(5,) + shape)
        conn, mel, use = oops.OracleOperator(kind, shape, **kw).find_conn(sigma)
        assert op.max_number_of_local_connections == conn.shape[0]
        flat = sigma.reshape(5, -1)
        for b in range(5):
            diag, nxt = 0.0, 1
            for (a, c, k, slot, dc, oc) in terms:
                sb = flat[b, c] if c >= 0 else 0
                if k != _lib.FK_TERM_FLIP:
                    diag += dc * flat[b, a] * sb
                if k == _lib.FK_TERM_DIAG:
                    continue
                used = (flat[b, a] != flat[b, c]) if k == _lib.FK_TERM_EXCHANGE else True
                if compact:
                    if not used:
                        continue
                    slot = nxt
                    nxt += 1
                new = flat[b].copy()
                if k == _lib.FK_TERM_EXCHANGE:
                    new[a], new[c] = flat[b, c], flat[b, a]
                else:
                    new[a] = -new[a]
                assert np.array_equal(conn[slot, b].reshape(-1), new)
                assert bool(use[slot, b]) == bool(used)
                assert mel[slot, b] == (oc if used else 0.0)
            assert mel[0, b] == pytest.approx(diag)


def test_mini_batch_generator_and_sampler_bookkeeping():
    from flowket_b200.optimization import MiniBatchGenerator
    from flowket_b200.samplers import Sampler

    class Gen(MiniBatchGenerator):
        calls = 0

        def next_batch(self):
            Gen.calls += 1
            return np.arange(10)[:, None] + 100 * Gen.calls, np.arange(10)

    g = Gen(10, 4)
    assert g.update_params_frequency == 3
    x, y = next(g)
    assert x[:, 0].tolist() == [100, 101, 102, 103]
    next(g)
    x, _ = next(g)           # 8 + 4 > 10 -> a new batch is drawn (mini_batch_generator.py:27-32)
    assert Gen.calls == 2 and x[0, 0] == 200
    assert Gen(4, 16).mini_batch_size == 4

    class S(Sampler):
        def __next__(self):
            return None
    s = S((3,), 8, mini_batch_size=32)
    assert s.mini_batch_size == 8 and s.batch_size == 8


def test_exact_utils_conventions(golden):
    from flowket_b200.exact import utils
    b = utils.decimal_array_to_binary_array(np.arange(32), 5, False)
    assert np.array_equal(b.astype(np.int8), golden['bits/binary'])
    assert np.array_equal(utils.binary_array_to_decimal_array(b), golden['bits/decimal'])
    assert utils.binary_to_decimal(utils.decimal_to_binary(19, 6)) == 19
    vec = np.arange(32) * (1 + 1j)
    assert np.array_equal(utils.vector_to_machine(vec)(b[[3, 7]])[:, 0], vec[[3, 7]])


def test_observable_generic_route_matches_reference_golden(golden):
    """Observable's host formulas (the route used for arbitrary psi callables) against the reference's outputs;
    find_conn is taken from the oracle here because there is no GPU in this tier."""
    from flowket_b200.observables.monte_carlo import Observable
    from flowket_b200.exact.utils import vector_to_machine
    obs = Observable(operator=None)
    psi = vector_to_machine(golden['handmade/log_psi_vector'])
    conn = golden['handmade/local_connections'].astype(np.float64)
    mel, use = golden['handmade/hamiltonian_values'], golden['handmade/all_use_conn']
    unb = obs.local_values_optimized_for_unbalanced_local_connections(psi, conn, mel, use)
    bal = obs.local_values_optimized_for_balanced_local_connections(psi, conn, mel)
    assert np.allclose(unb, golden['handmade/unbalanced'], rtol=1e-12)
    assert np.allclose(bal, golden['handmade/balanced'], rtol=1e-12)


def test_sr_algebra_against_pinv():
    """tests/test_stochastic_reconfiguration.py:32-72 of the reference (direct and CG, tol 1e-6, lambda 0.01):
    solve vs SVD pseudo-inverse of the hand-built S, on a random centred complex Jacobian (host torch path)."""
    from flowket_b200.optimizers import conjugate_gradient
    from oracle import sr as osr
    rng = np.random.default_rng(1)
    B, P, lam = 128, 16, 0.01
    O = rng.normal(size=(B, P)) + 1j * rng.normal(size=(B, P))
    Ob = osr.centre(O)
    rhs = rng.normal(size=P) + 1j * rng.normal(size=P)
    want = np.linalg.pinv(osr.s_matrix(Ob, lam)) @ rhs
    assert np.linalg.norm(osr.solve_direct(Ob, rhs, lam) - want) / np.linalg.norm(want) < 1e-5
    x, it, _ = osr.solve_iterative(Ob, rhs, lam, tol=1e-6, max_iter=None)
    assert np.linalg.norm(x - want) / np.linalg.norm(want) < 1e-5
    Ot = torch.from_numpy(Ob)
    xt, _, _ = conjugate_gradient(lambda v: Ot.conj().T @ (Ot @ v) / B + lam * v, torch.from_numpy(rhs), 1e-6, None)
    assert np.linalg.norm(xt.numpy() - want) / np.linalg.norm(want) < 1e-5


def test_sample_space_identity_equals_the_parameter_space_system():
    """the push-through identity behind the device SR pipeline (optimizers/sample_space_sr.py): with X = [Re Obar ; Im Obar]
    delta = X^T (X X^T / B + lambda I)^-1 e' / B solves the oracle's P x P real-parameter system exactly, for P > 2B and P < 2B"""
    from oracle import sr as osr
    rng = np.random.default_rng(5)
    for B, P in ((24, 200), (64, 40)):
        o_re, o_im = rng.normal(size=(B, P)), rng.normal(size=(B, P))
        eloc = rng.normal(size=B) * 3 - 7 + 1j * rng.normal(size=B)
        S, F = osr.real_sr_system(o_re, o_im, eloc, 0.05)
        want = np.linalg.solve(S, F)
        X = np.concatenate([o_re - o_re.mean(0), o_im - o_im.mean(0)])
        ec = eloc - eloc.mean()
        got = X.T @ np.linalg.solve(X @ X.T / B + 0.05 * np.eye(2 * B), np.concatenate([ec.real, ec.imag]) / B)
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-10


def test_lncosh_known_answers():
    """tests/test_tensorflow_complex_numbers_ops.py:6-33 of the reference (oracle restatement, complex128)."""
    for z in [2, 3j, 1 + 7j, 10 - 3j, -6]:
        zt = torch.tensor([z], dtype=torch.complex128)
        assert abs((nets.lncosh(zt) - torch.log(torch.cosh(zt))).item()) < 1e-8


def test_ensemble_ops_match_the_reference_formulas():
    """flowket/machines/ensemble.py:14-25 on a random [n, K] block of complex log-amplitudes"""
    import torch
    from flowket_b200.machines.ensemble import probabilistic_ensemble_op, average_ensemble_op
    rng = np.random.default_rng(0)
    x = rng.normal(size=(7, 8)) * 3 + 1j * rng.uniform(-3.1, 3.1, size=(7, 8))
    got = probabilistic_ensemble_op(torch.from_numpy(x)).numpy()[:, 0]
    re = 0.5 * np.log(np.exp(2 * x.real).sum(axis=1)) - 0.5 * np.log(8)
    im = np.angle(np.exp(1j * x.imag).mean(axis=1))
    assert np.allclose(got.real, re, atol=1e-12) and np.allclose(got.imag, im, atol=1e-12)
    got = average_ensemble_op(torch.from_numpy(x)).numpy()[:, 0]
    want = np.log(np.exp(x).mean(axis=1))
    assert np.allclose(got.real, want.real, atol=1e-12) and np.allclose(np.exp(1j * got.imag), np.exp(1j * want.imag), atol=1e-12)


def test_obc_ensemble_enumerates_the_dihedral_group():
    import torch
    from flowket_b200 import Input
    from flowket_b200.machines.ensemble import make_2d_obc_invariants, make_up_down_invariant, make_pbc_invariants
    inp = Input(shape=(4, 4))
    ens = make_2d_obc_invariants(inp, predictions_model=None)
    x = torch.arange(16, dtype=torch.int8).reshape(1, 4, 4)
    images = {tuple(t(x).reshape(-1).tolist()) for t in ens.transforms}
    a = x[0].numpy()
    want = {tuple(np.rot90(f, k).reshape(-1).tolist()) for k in range(4) for f in (a, a[:, ::-1])}
    assert len(ens.transforms) == 8 and images == want
    assert make_up_down_invariant(inp, ens).ensemble_size == 16
    assert len(make_pbc_invariants(inp, None, apply_also_obc_invariants=False).transforms) == 16


def test_flattened_operator_reshapes_only():
    """FlattenedOperator: a lattice operator for 1-D machines over raster-flattened sites (configs[3])"""
    from flowket_b200.operators import FlattenedOperator
    rng = np.random.default_rng(0)
    lattice = oops.OracleOperator('j1j2', (4, 3), j2=0.5, pbc=False)
    lattice.use_state = lambda s: s.shape == (4, 3)
    lattice.random_states = lambda k: rng.choice([-1, 1], size=(k, 4, 3))
    flat = FlattenedOperator(lattice)
    assert flat.hilbert_state_shape == (12,) and flat.max_number_of_local_connections == lattice.max_number_of_local_connections
    sigma = rng.choice([-1, 1], size=(5, 4, 3)).astype(np.float64)
    conn, mel, use = lattice.find_conn(sigma)
    fconn, fmel, fuse = flat.find_conn(sigma.reshape(5, 12))
    assert fconn.shape == conn.shape[:2] + (12,) and np.array_equal(fconn.reshape(conn.shape), conn)
    assert np.array_equal(fmel, mel) and np.array_equal(fuse, use)
    assert flat.use_state(sigma[0].reshape(12)) and flat.random_states(7).shape == (7, 12)


def test_distributed_exact_variational_degenerates_to_the_single_process_class():
    from flowket_b200.exact.utils import vector_to_machine
    from flowket_b200.optimization import ExactVariational, DistributedExactVariational
    import types
    rng = np.random.default_rng(2)
    vec = rng.normal(scale=0.5, size=32) + 1j * rng.uniform(-3, 3, size=32)
    f = vector_to_machine(vec)
    model = types.SimpleNamespace(input_shape=(None, 5), predict=lambda x, batch_size=None: f(np.asarray(x)))
    op = oops.OracleOperator('ising', (5,), h=0.7, pbc=True)
    a, b = ExactVariational(model, op, 8), DistributedExactVariational(model, op, 8)
    a.machine_updated()
    b.machine_updated()
    assert b.world_size == 1 and (b.slice_lo, b.slice_hi) == (0, 32)
    assert abs(a.energy_observable.current_energy - b.energy_observable.current_energy) < 1e-13
    assert np.allclose(a.energy_grad_coefficients, b.energy_grad_coefficients, atol=1e-14)
    with pytest.raises(Exception, match='must divide'):
        DistributedExactVariational(model, op, 5)


@pytest.mark.parametrize('name,opkind,shape,opkw', [
    ('heis_2x3_obc', 'heisenberg', (2, 3), dict(pbc=False)),
    ('ising_3x3_obc', 'ising', (3, 3), dict(pbc=False, h=3.0)),
    ('ising_1d8_pbc', 'ising', (8,), dict(pbc=True, h=0.7)),
])
def test_exact_variational_matches_the_reference_class(name, opkind, shape, opkw):
    """golden (oracle/make_golden.py exact): the reference's own ExactVariational / ExactObservable run on a log-amplitude
    table; this repository's class (and the oracle's) must reproduce probabilities, per-state energies, <H>, the variance and
    the gradient coefficients."""
    import os
    import types
    from flowket_b200.exact.utils import vector_to_machine
    from flowket_b200.optimization import ExactVariational
    from oracle import exact as oexact
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_exact_variational.npz'))
    vec = g[name + '/log_psi_vector']
    f = vector_to_machine(vec)
    model = types.SimpleNamespace(input_shape=(None,) + shape, predict=lambda x, batch_size=None: f(np.asarray(x)))
    op = oops.OracleOperator(opkind, shape, **opkw)
    ev = ExactVariational(model, op, int(g[name + '/batch_size']))
    ev.machine_updated()
    assert ev.num_of_batch_until_full_cycle == int(g[name + '/num_of_batch_until_full_cycle'])
    assert np.allclose(ev.probs, g[name + '/probs'], rtol=1e-12, atol=1e-15)
    assert np.allclose(ev.energy_observable.energies, g[name + '/energies'], rtol=1e-11, atol=1e-13)
    assert np.allclose(ev.energy_grad_coefficients, g[name + '/energy_grad_coefficients'], rtol=1e-10, atol=1e-13)
    assert ev.energy_observable.current_energy == pytest.approx(complex(g[name + '/current_energy']), rel=1e-12)
    assert ev.energy_observable.current_local_energy_variance == pytest.approx(
        float(g[name + '/current_local_energy_variance']), rel=1e-10)
    oev = oexact.ExactVariationalOracle(lambda s: f(np.asarray(s))[:, 0], op, shape, int(g[name + '/batch_size']))
    oev.machine_updated()
    assert oev.current_energy == pytest.approx(complex(g[name + '/current_energy']), rel=1e-12)
    assert np.allclose(oev.probs, g[name + '/probs'], rtol=1e-12, atol=1e-15)


def test_variational_monte_carlo_matches_the_reference_class():
    """golden (oracle/make_golden.py vmc): the reference's own VariationalMonteCarlo + MiniBatchGenerator driven by a scripted
    sampler and a log-amplitude table -- mini-batch windows, dropped tail, loss coefficients, energy and variance per batch.
    This repository's class runs the host observable route here (no device sampler); on the GPU the same quantities come
    from the device route (tests/test_gpu_vmc.py)."""
    import os
    import types
    from flowket_b200.exact.utils import vector_to_machine
    from flowket_b200.optimization import VariationalMonteCarlo
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_variational_monte_carlo.npz'))
    f = vector_to_machine(g['log_psi_vector'])
    batches = g['batches'].astype(np.float64)

    class ScriptedSampler(object):
        batch_size = 10

        def __init__(self):
            self.i = -1

        def __next__(self):
            self.i += 1
            return batches[self.i % len(batches)]

    model = types.SimpleNamespace(predict=lambda x, batch_size=None: f(np.asarray(x)))
    sampler = ScriptedSampler()
    vmc = VariationalMonteCarlo(model, oops.OracleOperator('heisenberg', (3, 4), pbc=False), sampler, mini_batch_size=4)
    assert vmc.update_params_frequency == int(g['update_params_frequency'])
    for step in range(len(g['energies'])):
        x, y = next(vmc)
        assert np.array_equal(np.asarray(x).astype(np.int8), g['mini_batches_x'][step]), step
        assert np.allclose(y, g['mini_batches_y'][step], rtol=1e-6, atol=1e-9), step      # ratios are complex64 in both
        assert vmc.current_energy == pytest.approx(complex(g['energies'][step]), rel=1e-6)
        assert vmc.current_local_energy_variance == pytest.approx(float(g['variances'][step]), rel=1e-5)
    assert np.allclose(vmc.current_local_energy, g['last_local_energy'], rtol=1e-6)
    assert sampler.i + 1 == int(g['batches_drawn'])


@pytest.mark.parametrize('name,spec', [
    ('conv2d_4x3', nets.Conv2DSpec(4, 3, 2, 8)),
    ('conv1d_10', nets.Conv1DSpec(10, 4, 8, max_dilation_rate=2)),
    ('cconv1d_8', nets.ComplexConv1DSpec(8, 3, 4, max_dilation_rate=2)),
])
def test_sampling_rule_matches_the_reference_sampler(name, spec):
    """golden (oracle/make_golden.py sampler): spins drawn by the reference's own AutoregressiveSampler.__next__ around the
    oracle network.  The oracle's sampler -- the parity contract the CUDA samplers are held to bit-exactly in
    tests/test_gpu_parity.py -- must reproduce them from the same uniforms, both site-by-site and incrementally."""
    import os
    from oracle import sampler as osampler
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_autoregressive_sampler.npz'))
    params = nets.unflatten_params(spec, torch.from_numpy(g[name + '/params']))
    sigma, p0 = osampler.sample_with_uniforms(spec, params, g[name + '/uniforms'])
    assert np.array_equal(sigma, g[name + '/sigma'])
    assert set(np.unique(sigma)) <= {-1, 1}
    if spec.kind == 'conv2d':
        inc = osampler.IncrementalSampler2D(spec, params)
        sigma_inc = inc.sample(g[name + '/uniforms'])
        sigma_inc = sigma_inc[0] if isinstance(sigma_inc, tuple) else sigma_inc
        # incremental == full up to near-ties of exp(log p0) against u (different fp32 summation order)
        assert (np.asarray(sigma_inc) != g[name + '/sigma']).mean() < 0.01


def test_lncosh_and_ensemble_ops_match_the_reference_functions():
    """golden (oracle/make_golden.py complex_ops): the reference's own lncosh / complex_log / angle and the two ensemble ops
    evaluated on numpy arrays through a one-to-one numpy stand-in for the elementwise TF math functions they compose."""
    import os
    from flowket_b200.machines.ensemble import probabilistic_ensemble_op, average_ensemble_op
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_complex_ops.npz'))
    z = torch.from_numpy(g['z'])
    got = nets.lncosh(z).numpy()
    assert np.allclose(got.real, g['lncosh'].real, rtol=1e-12, atol=1e-12)
    assert np.allclose(np.exp(1j * got.imag), np.exp(1j * g['lncosh'].imag), atol=1e-12)       # phase modulo 2 pi
    x = torch.from_numpy(g['ensemble_input'])
    got = probabilistic_ensemble_op(x).numpy()
    assert np.allclose(got.real, g['probabilistic_ensemble'].real, rtol=1e-12, atol=1e-12)
    assert np.allclose(np.exp(1j * got.imag), np.exp(1j * g['probabilistic_ensemble'].imag), atol=1e-12)
    got = average_ensemble_op(x).numpy()
    assert np.allclose(got.real, g['average_ensemble'].real, rtol=1e-11, atol=1e-12)
    assert np.allclose(np.exp(1j * got.imag), np.exp(1j * g['average_ensemble'].imag), atol=1e-11)


def test_split_solve_shares():
    """sample_space_sr.split_solve_shares: the solver rank's share balances T_factor + s0 T_E against (1 - s0) T_E / (N - 1);
    the shares always cover the batch exactly, whatever the ratio"""
    from flowket_b200.optimizers.sample_space_sr import split_solve_shares
    assert split_solve_shares(1, 100, 0.3) == [100]
    for world in (2, 3, 4, 8):
        for batch in (64, 1000, 8192):
            for rho in (0.0, 0.05, 0.087, 0.21, 1.0, 50.0):
                for solver_rank in (0, world - 1):
                    c = split_solve_shares(world, batch, rho, solver_rank)
                    assert len(c) == world and sum(c) == batch and min(c) >= 0
                    others = [v for r, v in enumerate(c) if r != solver_rank]
                    assert max(others) - min(others) <= 1
                    # balance: T_E = 1, T_c = rho
                    t_solver = rho + c[solver_rank] / batch
                    t_other = max(others) / batch
                    if c[solver_rank] > 0:
                        assert abs(t_solver - t_other) <= 2.0 / batch + 1e-12
                    else:
                        assert t_solver >= t_other - 2.0 / batch
    assert split_solve_shares(8, 8192, 0.0) == [1024] * 8
    # a solver rank whose local energies run at a different per-sample speed (kappa = its time per sample / the others')
    for world in (2, 4, 8):
        for kappa in (0.8, 1.0, 1.3):
            c = split_solve_shares(world, 8192, 0.05, 0, kappa)
            assert sum(c) == 8192
            assert abs((0.05 + kappa * c[0] / 8192.0) - c[1] / 8192.0) <= 3.0 / 8192


def test_generator_accepts_local_energies_evaluated_elsewhere():
    """VariationalMonteCarlo.next_samples / set_local_energy (split solve of the sharded SR step): the statistics are those of
    next_batch on the same values"""
    import torch
    from flowket_b200.optimization import VariationalMonteCarlo

    class Sampler(object):
        batch_size = 6

        def __iter__(self):
            return self

        def __next__(self):
            return np.ones((6, 4), np.int8)

    class Obs(object):
        pass

    from flowket_b200.observables.monte_carlo import BaseObservable

    class Fixed(BaseObservable):
        def estimate(self, wave_function, configurations):
            lv = np.arange(6) * (1.0 + 0.5j)
            return np.mean(lv), np.var(np.real(lv)), lv

    class Net(object):
        def predict(self, x, batch_size=None):
            return np.zeros(len(x), np.complex64)

    vmc = VariationalMonteCarlo(Net(), Fixed(), Sampler())
    vmc.next_batch()
    want = (vmc.current_energy, vmc.current_local_energy_variance, vmc.current_local_energy.copy())
    x = vmc.next_samples()
    assert x.shape == (6, 4) and vmc.current_local_energy is None
    vmc.set_local_energy(torch.as_tensor(want[2]))
    assert vmc.current_energy == want[0] and vmc.current_local_energy_variance == want[1]
    assert np.array_equal(vmc.current_local_energy, want[2])
