"""GPU tier: the orchestration objects (VariationalMonteCarlo, ExactVariational, Trainer, SR) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import exact as oexact
from oracle import local_energy as oeloc
from oracle import nets, operators as oops, sr as osr
from tests.helpers import make_pair, random_sigma

pytestmark = pytest.mark.gpu


def test_variational_monte_carlo_attributes_and_coefficients():
    """same attribute names / return values as optimization/variational_monte_carlo.py:15-50"""
    from flowket_b200.operators import Heisenberg
    from flowket_b200.samplers import FastAutoregressiveSampler
    from flowket_b200.optimization import VariationalMonteCarlo
    shape = (4, 4)
    model, cond, spec, params = make_pair('conv2d', shape, 3, 32, seed=1)
    sampler = FastAutoregressiveSampler(cond, 48)
    vmc = VariationalMonteCarlo(model, Heisenberg(hilbert_state_shape=list(shape), pbc=False), sampler, mini_batch_size=16)
    assert vmc.update_params_frequency == 3
    x, y = next(vmc)
    assert x.shape == (16, 4, 4) and y.shape == (16,) and x.dtype == np.int8
    for attr in ('current_batch', 'current_energy', 'current_local_energy_variance', 'current_local_energy',
                 'wave_function', 'sampler', 'batch_size', 'mini_batch_size', 'start_time', 'sampling_end_time',
                 'local_energy_end_time'):
        assert hasattr(vmc, attr), attr
    sigma = vmc.current_batch
    want = oeloc.local_values(oops.OracleOperator('heisenberg', shape, pbc=False),
                              lambda c: nets.log_psi_numpy(spec, params, c), sigma.astype(np.float64))
    assert np.abs(vmc.current_local_energy - want).max() / np.abs(want).max() < 1e-5
    assert vmc.current_energy == pytest.approx(np.mean(want), rel=1e-5)
    assert vmc.current_local_energy_variance == pytest.approx(np.var(want.real), rel=1e-4)
    coeff = oeloc.loss_coefficients(want, want.mean(), 48)
    assert np.abs(y - coeff[:16]).max() < 1e-5 * np.abs(coeff).max() + 1e-9
    # the wave_function callable of the reference protocol still works (Observable generic route)
    e2 = vmc.energy_observable.local_values(vmc.wave_function, sigma)
    assert np.abs(e2 - want).max() / np.abs(want).max() < 1e-5


@pytest.mark.parametrize('opkind,opkw,shape,kind,depth,ch', [
    ('ising', dict(pbc=False, h=3.0), (4, 4), 'conv2d', 3, 16),
    ('heisenberg', dict(pbc=True), (12,), 'conv1d', 5, 16),
    ('j1j2', dict(pbc=False, j2=0.5), (4, 3), 'conv2d', 2, 8),
])
def test_exact_variational_matches_oracle(opkind, opkw, shape, kind, depth, ch):
    from flowket_b200.optimization import ExactVariational
    from tests.test_gpu_parity import _product_operator
    model, _, spec, params = make_pair(kind, shape, depth, ch, seed=4)
    ev = ExactVariational(model, _product_operator(opkind, shape, opkw), batch_size=2 ** 10)
    ev.machine_updated()
    oop = oops.OracleOperator(opkind, shape, **opkw)
    oev = oexact.ExactVariationalOracle(lambda s: nets.log_psi_numpy(spec, params, s)[:, 0], oop, shape, 2 ** 10)
    oev.machine_updated()
    assert ev.num_of_states == 2 ** int(np.prod(shape))
    assert ev.energy_observable.current_energy == pytest.approx(oev.current_energy, rel=2e-5)
    assert np.abs(ev.probs - oev.probs).max() < 1e-5 * oev.probs.max()
    assert ev.energy_observable.current_local_energy_variance == pytest.approx(oev.current_local_energy_variance, rel=1e-3)
    assert np.linalg.norm(ev.energy_grad_coefficients - oev.energy_grad_coefficients) < \
        2e-5 * np.linalg.norm(oev.energy_grad_coefficients)
    with pytest.raises(Exception):
        ExactVariational(model, _product_operator(opkind, shape, opkw), batch_size=1000)


def test_exact_equals_monte_carlo_energy():
    """tests/test_variational.py:49-66 of the reference: MC energy == exact energy within the MC error bar."""
    from flowket_b200.operators import Heisenberg
    from flowket_b200.samplers import FastAutoregressiveSampler
    from flowket_b200.optimization import ExactVariational, VariationalMonteCarlo
    shape = (3, 4)
    model, cond, _, _ = make_pair('conv2d', shape, 3, 32, seed=6)
    op = Heisenberg(hilbert_state_shape=list(shape), pbc=False)
    ev = ExactVariational(model, op, 2 ** 12)
    ev.machine_updated()
    exact_e = ev.energy_observable.current_energy.real
    var = ev.energy_observable.current_local_energy_variance
    B, iters = 4096, 4
    vmc = VariationalMonteCarlo(model, op, FastAutoregressiveSampler(cond, B, seed=11))
    energies = []
    for _ in range(iters):
        vmc.next_batch()
        energies.append(vmc.current_energy.real)
    sigma_mc = np.sqrt(var / (B * iters))
    assert abs(np.mean(energies) - exact_e) < 4 * sigma_mc


def test_complex_sr_matches_oracle():
    from flowket_b200.optimizers import ComplexValuesStochasticReconfiguration
    shape = (10,)
    model, _, spec, params = make_pair('cconv1d', shape, 3, 8, seed=3)
    B = 64
    sigma = random_sigma(B, shape, seed=8)
    rng = np.random.default_rng(0)
    eloc = rng.normal(size=B) + 1j * rng.normal(size=B)
    y_true = np.conj(eloc - eloc.mean()) / B
    # oracle: complex Jacobian assembled pairwise from (real, imag) variables
    O_re = nets.per_sample_gradients(spec, params, sigma, 'real').numpy()
    offs = np.cumsum([0] + [int(p.numel()) for p in params])
    re_idx = np.concatenate([np.arange(offs[i], offs[i + 1]) for i in range(0, len(params), 4)] +
                            [np.arange(offs[i + 2], offs[i + 3]) for i in range(0, len(params), 4)])
    im_idx = np.concatenate([np.arange(offs[i + 1], offs[i + 2]) for i in range(0, len(params), 4)] +
                            [np.arange(offs[i + 3], offs[i + 4]) for i in range(0, len(params), 4)])
    for iterative in (False, True):
        sr = ComplexValuesStochasticReconfiguration(model, diag_shift=0.05, iterative_solver=iterative,
                                                    conjugate_gradient_tol=1e-7, iterative_solver_max_iterations=2000)
        got = sr.compute_update(sigma, y_true).cpu().numpy()
        re, im = [t.cpu().numpy() for t in sr._complex_index('cpu')]
        O = O_re[:, re] + 1j * O_re[:, im]
        Ob = osr.centre(O)
        want = osr.solve_direct(Ob, osr.energy_grad(Ob, y_true), 0.05)
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < 2e-4, iterative
    # and the update moves the parameters as W <- W - lr * delta
    before = model.machine.flat_params_numpy()
    delta = sr.step(sigma, y_true).cpu().numpy()
    after = model.machine.flat_params_numpy()
    W0 = before[re] - 1j * before[im]
    W1 = after[re] - 1j * after[im]
    assert np.allclose(W1, W0 - sr.lr * delta, atol=1e-6)


def test_real_sr_matches_oracle():
    from flowket_b200.optimizers import StochasticReconfiguration
    shape = (4, 4)
    model, _, spec, params = make_pair('conv2d', shape, 2, 8, seed=5, weights_normalization=False)
    B = 96
    sigma = random_sigma(B, shape, seed=9)
    rng = np.random.default_rng(1)
    eloc = rng.normal(size=B) * 2 - 10 + 1j * rng.normal(size=B)
    O_re = nets.per_sample_gradients(spec, params, sigma, 'real').numpy()
    O_im = nets.per_sample_gradients(spec, params, sigma, 'imag').numpy()
    S, F = osr.real_sr_system(O_re, O_im, eloc, 0.05)
    want = np.linalg.solve(S, F)
    for iterative in (False, True):
        sr = StochasticReconfiguration(model, diag_shift=0.05, iterative_solver=iterative, conjugate_gradient_tol=1e-7,
                                       iterative_solver_max_iterations=5000, sample_space=False)
        got = sr.compute_update(sigma, eloc).cpu().numpy()
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < 5e-4, iterative
    # sample-space form (2B x 2B Gram over the parameter axis): the same delta by the push-through identity
    for gram_dtype, tol in (('fp32', 5e-4), ('bf16', 2e-2)):
        sr = StochasticReconfiguration(model, diag_shift=0.05, sample_space=True, gram_dtype=gram_dtype, jacobian_chunk=40)
        got = sr.compute_update(sigma, eloc).cpu().numpy()
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < tol, gram_dtype
        # the sharded code path (global centring, re-shard, partial Gram, replicated solve) on a world of one rank
        srd = StochasticReconfiguration(model, diag_shift=0.05, sample_space=True, gram_dtype=gram_dtype, jacobian_chunk=40,
                                        distributed=True)
        got = srd.compute_update(sigma, eloc).cpu().numpy()
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < tol, gram_dtype
        assert set(srd.last_timings_ms) == {'jacobian', 'exchange', 'gram', 'cholesky', 'update'}


def test_sr_gram_kernel():
    from flowket_b200._device import sr_gram
    rng = np.random.default_rng(2)
    A = torch.from_numpy(rng.normal(size=(300, 97)).astype(np.float32)).cuda()
    G1 = sr_gram(A, transpose_a=True).cpu().numpy()
    G2 = sr_gram(A, transpose_a=False).cpu().numpy()
    An = A.cpu().numpy().astype(np.float64)
    assert np.abs(G1 - An.T @ An).max() < 1e-4 * np.abs(An.T @ An).max()
    assert np.abs(G2 - An @ An.T).max() < 1e-4 * np.abs(An @ An.T).max()


@pytest.mark.parametrize('rows,cols,transpose_a', [(300, 97, True), (300, 97, False), (1024, 700, True), (64, 1000, False),
                                                   (2048, 129, True)])
def test_sr_gram_tensor_core_kernel(rows, cols, transpose_a):
    """fk_sr_gram_tc: hi/lo fp16 operands -> 3e-5 of the largest entry (what is left is the truncating fp32 accumulation of
    the tensor core, ~1e-6 per few hundred samples); single pass -> 2e-3"""
    from flowket_b200._device import sr_gram
    from flowket_b200 import FK_ENGINE_TC
    rng = np.random.default_rng(3)
    A = torch.from_numpy((rng.normal(size=(rows, cols)) * rng.uniform(0.01, 3.0, size=(1, cols) if transpose_a else (rows, 1))
                          ).astype(np.float32)).cuda()
    An = A.cpu().numpy().astype(np.float64)
    want = An.T @ An if transpose_a else An @ An.T
    got = sr_gram(A, transpose_a=transpose_a, engine=FK_ENGINE_TC, precise=True).cpu().numpy()
    assert got.shape == want.shape and np.isfinite(got).all()
    assert np.abs(got - want).max() < 3e-5 * np.abs(want).max()
    assert np.array_equal(got, got.T)
    fast = sr_gram(A, transpose_a=transpose_a, engine=FK_ENGINE_TC, precise=False).cpu().numpy()
    assert np.abs(fast - want).max() < 2e-3 * np.abs(want).max()


def test_training_lowers_the_energy_towards_exact_diagonalisation():
    """Ising 4x4 OBC h=3 (cfg 1 anchor: ED -50.18662388277671, examples/basic_autoregressive_2d.py:39):
    exact-gradient Adam steps must move the variational energy monotonically-ish towards the ED value and stay
    above it (variational principle)."""
    from flowket_b200.operators import Ising
    from flowket_b200.optimization import ExactVariational
    from flowket_b200.optimizers import Adam, Trainer
    from flowket_b200 import Input, Model
    from flowket_b200.machines import ConvNetAutoregressive2D
    inp = Input(shape=(4, 4), dtype='int8')
    machine = ConvNetAutoregressive2D(inp, depth=3, num_of_channels=16, weights_normalization=False, seed=0)
    model = Model(inp, machine.predictions)
    ev = ExactVariational(model, Ising(hilbert_state_shape=[4, 4], pbc=False, h=3.0), batch_size=2 ** 14)
    gen = ev.to_generator()
    opt = Adam(lr=3e-3, beta_1=0.9, beta_2=0.999)

    class ExactGen(object):
        update_params_frequency = ev.num_of_batch_until_full_cycle

        def __next__(self_inner):
            return next(gen)
    trainer = Trainer(model, ExactGen(), opt)
    # Keras averages the loss over the mini-batch; the exact coefficients already carry p(s): undo the 1/mb
    energies = []
    for step in range(120):
        grad = None
        for _ in range(ev.num_of_batch_until_full_cycle):
            x, y = next(gen)
            g = trainer.gradient(x, y) * x.shape[0]
            grad = g if grad is None else grad + g
        opt.step(machine.flat_params_device(), grad)
        machine.params_updated()
        energies.append(ev.energy_observable.current_energy.real)
    e_ed = -50.18662388277671
    print('exact-gradient training: E0 = %.6f -> E = %.6f (ED %.6f)' % (energies[0], energies[-1], e_ed))
    assert energies[-1] < energies[0]
    assert all(e > e_ed - 1e-6 for e in energies)             # variational principle
    assert abs(energies[-1] - e_ed) / abs(e_ed) < 5e-4        # ground-state energy of the 4x4 lattice vs ED


def test_symmetrised_ensemble_matches_manual_combination_and_is_invariant():
    """machines/ensemble.py: psi_sym from 16 forwards; E_loc through the device-generic route == host protocol route"""
    from flowket_b200 import Input
    from flowket_b200.machines import make_2d_obc_invariants, make_up_down_invariant
    from flowket_b200.operators import Ising
    from flowket_b200.observables.monte_carlo import Observable
    shape = (4, 4)
    model, _, spec, params = make_pair('conv2d', shape, 3, 8, seed=3)
    inp = Input(shape=shape)
    ens = make_up_down_invariant(inp, make_2d_obc_invariants(inp, model))
    x = random_sigma(20, shape, seed=4)
    got = ens.predict(x)[:, 0]
    # manual: inner D4 ensemble for +x and -x, then the outer two-member ensemble (nested exactly like the reference)
    def comb(vals):
        vals = np.asarray(vals).T
        re = 0.5 * np.log(np.exp(2 * vals.real).sum(axis=1)) - 0.5 * np.log(vals.shape[1])
        return re + 1j * np.angle(np.exp(1j * vals.imag).mean(axis=1))
    inner = []
    for s in (1, -1):
        imgs = [np.rot90(s * x, k, axes=(1, 2)) for k in range(4)]
        imgs = imgs + [im[:, :, ::-1] for im in imgs]
        inner.append(comb([nets.log_psi_numpy(spec, params, im.copy())[:, 0] for im in imgs]))
    want = comb(inner)
    assert np.allclose(got.real, want.real, atol=2e-4)
    assert np.allclose(np.exp(1j * got.imag), np.exp(1j * want.imag), atol=2e-4)
    # invariance under the group
    for g in (lambda a: np.rot90(a, 1, axes=(1, 2)).copy(), lambda a: a[:, :, ::-1].copy(), lambda a: -a):
        again = ens.predict(g(x))[:, 0]
        assert np.allclose(again.real, got.real, atol=1e-4)
    # local energies: device-generic route vs the reference's host protocol with the same callable
    obs = Observable(Ising(hilbert_state_shape=list(shape), pbc=False, h=3.0))
    e_dev = obs.local_values(ens.predict, x)
    conn, mel, use = obs.operator.find_conn(x)
    e_host = obs.local_values_optimized_for_balanced_local_connections(ens.predict, conn, mel)
    assert np.allclose(e_dev, e_host, rtol=1e-4, atol=1e-4)


def test_vmc_with_a_symmetrised_ensemble_evaluates_the_ensemble():
    """VariationalMonteCarlo(invariant_model, operator, sampler) (examples/j1j2_2d_monte_carlo.py:95 of the reference): the
    local energies must be those of the symmetrised wave function, not of the base network."""
    from flowket_b200 import Input
    from flowket_b200.machines import make_2d_obc_invariants, make_up_down_invariant
    from flowket_b200.operators import Ising
    from flowket_b200.observables.monte_carlo import Observable
    from flowket_b200.optimization import VariationalMonteCarlo
    from flowket_b200.samplers import FastAutoregressiveSampler
    shape = (4, 4)
    model, cond, spec, params = make_pair('conv2d', shape, 3, 8, seed=3)
    inp = Input(shape=shape)
    ens = make_up_down_invariant(inp, make_2d_obc_invariants(inp, model))
    op = Ising(hilbert_state_shape=list(shape), pbc=False, h=3.0)
    vmc = VariationalMonteCarlo(ens, op, FastAutoregressiveSampler(cond, 48, seed=5))
    batch, _ = vmc.next_batch()
    obs = Observable(op)
    want = obs.local_values(ens.predict, batch)
    base = obs.local_values(model, batch)
    assert np.allclose(vmc.current_local_energy, want, rtol=1e-5, atol=1e-5)
    assert np.abs(vmc.current_local_energy - base).max() > 1e-3        # and they do differ from the base network's
    assert vmc.current_energy == pytest.approx(want.mean())


def test_checkpoint_resume_continues_the_same_trajectory(tmp_path):
    """Trainer.save_checkpoint / load_checkpoint (weights, Adam slots, sampler counter): 3 + 3 steps == 6 steps up to the
    summation-order noise of the atomics in the gradient kernels (a lost Adam slot or sampler counter shifts the
    parameters by ~lr = 1e-2)"""
    from flowket_b200 import Input, Model
    from flowket_b200.machines import ConvNetAutoregressive2D
    from flowket_b200.operators import Heisenberg
    from flowket_b200.samplers import FastAutoregressiveSampler
    from flowket_b200.optimization import VariationalMonteCarlo
    from flowket_b200.optimizers import Adam, Trainer

    def build():
        inp = Input(shape=(4, 4), dtype='int8')
        m = ConvNetAutoregressive2D(inp, depth=3, num_of_channels=8, seed=11)
        model = Model(inp, m.predictions)
        cond = Model(inp, m.conditional_log_probs)
        vmc = VariationalMonteCarlo(model, Heisenberg(hilbert_state_shape=[4, 4], pbc=False),
                                    FastAutoregressiveSampler(cond, 64, seed=5))
        return m, Trainer(model, vmc, Adam(lr=1e-2, beta_1=0.9, beta_2=0.9))

    m_ref, t_ref = build()
    t_ref.fit(6)
    m_a, t_a = build()
    t_a.fit(3, checkpoint_path=str(tmp_path / 'ckpt'))
    m_b, t_b = build()
    t_b.load_checkpoint(str(tmp_path / 'ckpt'))
    t_b.fit(3)
    assert (m_b.flat_params_device() - m_ref.flat_params_device()).abs().max().item() < 2e-4
    assert abs(t_b.history[-1] - t_ref.history[-1]) < 1e-3


def test_metropolis_hastings_against_exact_distribution():
    """SURVEY 8f-4 cross-check (reference: tests/test_samplers.py:49-82): MH chains driven by the CUDA forward and the
    device find_conn reproduce |psi|^2 of the same machine, which the exact autoregressive sampler also reproduces."""
    from flowket_b200.exact.utils import binary_array_to_decimal_array
    from flowket_b200.operators import Heisenberg
    from flowket_b200.optimization import ExactVariational
    from flowket_b200.samplers import MetropolisHastingsLocal, MetropolisHastingsHamiltonian, ExactSampler, \
        FastAutoregressiveSampler
    shape = (3, 2)
    model, cond, spec, params = make_pair('conv2d', shape, 2, 8, seed=2)
    op = Heisenberg(hilbert_state_shape=list(shape), pbc=False)
    ev = ExactVariational(model, op, batch_size=64)
    ev.machine_updated()

    def l1(batch, probs):
        idx = binary_array_to_decimal_array(np.asarray(batch).reshape(len(batch), -1))
        return np.abs(np.bincount(idx, minlength=64) / float(len(batch)) - probs).sum()

    local = MetropolisHastingsLocal(model, 256 * 60, num_of_chains=256, unused_sampels=5, discard_ratio=3, seed=1)
    assert l1(next(local), ev.probs) < 0.1
    assert l1(next(ExactSampler(ev, 1 << 14, seed=0)), ev.probs) < 0.1
    assert l1(next(FastAutoregressiveSampler(cond, 1 << 14)), ev.probs) < 0.1
    # Hamiltonian moves stay in the S_z = 0 sector: compare with the conditional distribution
    ham = MetropolisHastingsHamiltonian(model, 256 * 60, op, num_of_chains=256, unused_sampels=5, discard_ratio=3, seed=1)
    batch = next(ham)
    assert np.all(batch.reshape(len(batch), -1).sum(axis=1) == 0)
    in_sector = ev.states.reshape(64, -1).sum(axis=1) == 0
    sector_probs = np.where(in_sector, ev.probs, 0.0)
    assert l1(batch, sector_probs / sector_probs.sum()) < 0.1
    assert 0.0 < ham.acceptance_ratio <= 1.0


def test_fit_with_callbacks_and_evaluate(tmp_path):
    """Trainer.fit drives the reference's callback protocol (experiments/train.py:117-130), evaluation.evaluate the
    evaluation protocol (evaluation/evaluate.py:16-29); logged energies are the generator's device results."""
    from flowket_b200 import Input, Model
    from flowket_b200.callbacks import default_wave_function_stats_callbacks_factory, CheckpointByTime
    from flowket_b200.callbacks.monte_carlo import BadEigenStateStopping
    from flowket_b200.evaluation import evaluate
    from flowket_b200.machines import ConvNetAutoregressive2D
    from flowket_b200.operators import Heisenberg
    from flowket_b200.samplers import FastAutoregressiveSampler
    from flowket_b200.optimization import VariationalMonteCarlo
    from flowket_b200.optimizers import Adam, Trainer
    inp = Input(shape=(4, 4), dtype='int8')
    m = ConvNetAutoregressive2D(inp, depth=3, num_of_channels=8, seed=11)
    model, cond = Model(inp, m.predictions), Model(inp, m.conditional_log_probs)
    op = Heisenberg(hilbert_state_shape=[4, 4], pbc=False)
    vmc = VariationalMonteCarlo(model, op, FastAutoregressiveSampler(cond, 128, seed=5))
    val = VariationalMonteCarlo(model, op, FastAutoregressiveSampler(cond, 256, seed=6))
    callbacks = default_wave_function_stats_callbacks_factory(vmc, validation_generator=val,
                                                              true_ground_state_energy=-36.7546)
    ckpt = CheckpointByTime(str(tmp_path / 'ck'), save_frequency_in_minutes=1e9)
    trainer = Trainer(model, vmc, Adam(lr=1e-2, beta_1=0.9, beta_2=0.9))
    trainer.fit(6, callbacks=callbacks + [ckpt, BadEigenStateStopping(-36.7546, min_epoch=100)], steps_per_epoch=2)
    assert len(trainer.logs) == 3 and ckpt.saves == 1 and (tmp_path / 'ck.npz').exists()
    last = trainer.logs[-1]
    for key in ('energy/energy', 'energy/local_energy_variance', 'energy/relative_error', 'observables/sigma_z',
                'observables/abs_sigma_z', 'times/sampling', 'times/local_energy', 'times/gradients', 'times/total',
                'val_energy/energy', 'val_observables/abs_sigma_z'):
        assert key in last, key
    assert last['energy/energy'] == pytest.approx(np.real(vmc.current_energy))
    assert last['val_energy/energy'] == pytest.approx(np.real(val.current_energy))
    assert last['observables/sigma_z'] == pytest.approx(vmc.current_batch.reshape(128, -1).sum(axis=1).mean() / 16.0)
    res = evaluate(val, 4, callbacks[1:], verbose=False)
    assert abs(res['energy/energy'] - last['val_energy/energy']) < 2.0
    assert res['energy/local_energy_variance'] > 0
    # a callback may stop the loop
    stopper = BadEigenStateStopping(-36.7546, variance_tol=1e9, relative_error_to_stop=-1e9, min_epoch=0)
    before = len(trainer.history)
    trainer.fit(5, callbacks=callbacks + [stopper])
    assert len(trainer.history) == before + 1 and stopper.stopped_epoch == 0


def test_compile_and_fit_generator_exact_gradient_script_flow():
    """The calls of the reference's examples/basic_autoregressive_exact_gradient.py (BASELINE configs[0]) on a 10-site
    chain: compile, accumulate-gradient optimizer over one enumeration cycle, fit_generator with the exact callbacks."""
    from flowket_b200 import Input, Model
    from flowket_b200.callbacks.exact import default_wave_function_callbacks_factory
    from flowket_b200.machines import SimpleConvNetAutoregressive1D
    from flowket_b200.operators import Ising
    from flowket_b200.optimization import ExactVariational, loss_for_energy_minimization
    from flowket_b200.optimizers import Adam, convert_to_accumulate_gradient_optimizer
    inputs = Input(shape=[10], dtype='int8')
    convnet = SimpleConvNetAutoregressive1D(inputs, depth=4, num_of_channels=16, weights_normalization=False, seed=0)
    model = Model(inputs=inputs, outputs=convnet.predictions)
    operator = Ising(h=3.0, hilbert_state_shape=[10], pbc=False)
    ev = ExactVariational(model, operator, 2 ** 8)
    assert ev.num_of_batch_until_full_cycle == 4
    optimizer = Adam(lr=0.01, beta_1=0.9, beta_2=0.999)
    convert_to_accumulate_gradient_optimizer(optimizer, update_params_frequency=ev.num_of_batch_until_full_cycle,
                                             accumulate_sum_or_mean=True)
    model.compile(optimizer=optimizer, loss=loss_for_energy_minimization)
    lines = []
    model.summary(print_fn=lines.append)
    assert 'Total params' in lines[-1]
    e_ed = oexact.ground_state(oops.OracleOperator('ising', (10,), h=3.0, pbc=False), (10,))[0]
    callbacks = default_wave_function_callbacks_factory(ev, true_ground_state_energy=e_ed)
    logs = model.fit_generator(ev.to_generator(), steps_per_epoch=4 * 20, epochs=3, callbacks=callbacks,
                               max_queue_size=0, workers=0)
    assert len(logs) == 3 and optimizer.accumulated_iterations == 240 and optimizer.t == 60
    energies = [l['energy/energy'] for l in logs]
    assert energies[-1] < energies[0] - 1e-3                      # training lowers the exact energy
    assert all(e > e_ed - 1e-6 for e in energies)                 # variational principle
    assert logs[-1]['energy/relative_error'] == pytest.approx((e_ed - energies[-1]) / e_ed)
    assert 'observables/sigma_z' in logs[-1] and 'times/total' in logs[-1]
