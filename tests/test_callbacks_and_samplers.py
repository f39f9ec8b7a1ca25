"""CPU tier for SURVEY.md section 8f-4: stats callbacks, evaluate, the enumerated samplers and the
Metropolis-Hastings samplers (host logic; the wave function here is a table, on the GPU box it is the CUDA forward --
see tests/test_gpu_vmc.py::test_metropolis_hastings_against_exact_distribution).

The sampler checks follow the reference's own test (tests/test_samplers.py:49-82): draw many samples, histogram them
over all 2^N states, compare with |psi|^2 in L1."""
import time
import types

import numpy as np
import pytest

from flowket_b200.exact.utils import binary_array_to_decimal_array, decimal_array_to_binary_array, vector_to_machine
from oracle import operators as oops


class TableMachine(object):
    """A normalised random log-amplitude table behind the `.predict` / `.input_shape` protocol of a Model."""

    def __init__(self, shape, seed=0, sz0_only=False):
        self.shape = tuple(shape)
        n = int(np.prod(shape))
        rng = np.random.default_rng(seed)
        log_amp = rng.normal(scale=0.7, size=2 ** n) + 1j * rng.uniform(-np.pi, np.pi, size=2 ** n)
        if sz0_only:
            states = decimal_array_to_binary_array(np.arange(2 ** n), n)
            log_amp[states.sum(axis=1) != 0] = -60.0
        log_amp -= 0.5 * np.log(np.exp(2 * log_amp.real).sum())
        self.vector = log_amp
        self.probs = np.exp(2 * log_amp.real)
        self._machine = vector_to_machine(self.vector)
        self.calls = 0

    @property
    def input_shape(self):
        return (None,) + self.shape

    def predict(self, x, batch_size=None):
        self.calls += 1
        return self._machine(np.asarray(x))


def histogram(samples, n):
    idx = binary_array_to_decimal_array(np.asarray(samples).reshape(len(samples), -1))
    return np.bincount(idx, minlength=2 ** n) / float(len(samples))


def test_exact_and_wave_function_samplers_follow_the_distribution():
    from flowket_b200.samplers import ExactSampler, WaveFunctionSampler
    table = TableMachine((2, 3), seed=1)
    ev = types.SimpleNamespace(input_size=(2, 3), num_of_states=64, number_of_spins=6, probs=table.probs)
    for sampler in (ExactSampler(ev, 20000, seed=3), WaveFunctionSampler(table.vector, (2, 3), 20000, seed=4)):
        batch = next(sampler)
        assert batch.shape == (20000, 2, 3) and set(np.unique(batch)) == {-1.0, 1.0}
        assert np.abs(histogram(batch, 6) - table.probs).sum() < 0.06
    # ExactSampler reads the *current* probs (the machine moves during training)
    sampler = ExactSampler(ev, 500, seed=5)
    ev.probs = np.zeros(64)
    ev.probs[37] = 1.0
    assert np.all(binary_array_to_decimal_array(next(sampler).reshape(500, -1)) == 37)
    assert iter(sampler) is sampler and sampler.mini_batch_size == 500


@pytest.mark.parametrize('kind', ['local', 'uniform', 'exchange', 'hamiltonian', 'global'])
def test_metropolis_hastings_samplers_converge_to_psi_squared(kind):
    from flowket_b200 import samplers as S
    shape = (2, 2) if kind != 'local' else (5,)
    n = int(np.prod(shape))
    sz0 = kind in ('exchange', 'hamiltonian')
    table = TableMachine(shape, seed=7, sz0_only=sz0)
    kwargs = dict(num_of_chains=64, unused_sampels=3, discard_ratio=4, seed=11)
    if kind == 'local':
        sampler = S.MetropolisHastingsLocal(table, 64 * 400, **kwargs)
    elif kind == 'uniform':
        sampler = S.MetropolisHastingsUniform(table, 64 * 400, **kwargs)
    elif kind == 'exchange':
        sampler = S.MetropolisHastingsExchange(table, 64 * 400, **kwargs)
        base = np.array([1, 1, -1, -1])          # exchange moves conserve S_z: start every chain at S_z = 0
        sampler.sample = np.stack([sampler.rng.permutation(base).reshape(shape) for _ in range(64)])
        sampler.candidates = sampler.sample.copy()
    elif kind == 'hamiltonian':
        op = oops.OracleOperator('heisenberg', shape, pbc=False)
        op.random_states = lambda k: np.stack([np.random.default_rng(i).permutation([1, 1, -1, -1]).reshape(shape)
                                               for i in range(k)])
        sampler = S.MetropolisHastingsHamiltonian(table, 64 * 400, op, **kwargs)
    else:
        proposals = S.WaveFunctionSampler(np.full(2 ** n, -0.5 * n * np.log(2.0)), shape, 64, seed=2)   # flat proposal
        kwargs.pop('num_of_chains')
        sampler = S.MetropolisHastingsGlobal(table, 64 * 400, proposals, **kwargs)
        assert sampler.num_of_chains == 64
    batch = next(sampler)
    assert batch.shape == (64 * 400,) + shape
    assert np.abs(histogram(batch, n) - table.probs).sum() < 0.08
    assert 0.0 < sampler.acceptance_ratio <= 1.0
    if sz0:
        assert np.all(batch.reshape(len(batch), -1).sum(axis=1) == 0)
    # chain-major layout: consecutive rows of one chain differ by few spins for the local sampler
    if kind == 'local':
        chains = batch.reshape(64, 400, n)
        assert (np.abs(np.diff(chains, axis=1)).sum(axis=2) / 2).max() <= 4      # <= (unused + 1) flips between keeps


def test_metropolis_hastings_rejects_bad_chain_count_and_non_finite_values():
    from flowket_b200 import samplers as S
    table = TableMachine((4,), seed=0)
    with pytest.raises(Exception, match='must divide'):
        S.MetropolisHastingsLocal(table, 10, num_of_chains=4)
    table.vector[:] = np.nan
    sampler = S.MetropolisHastingsLocal(table, 8, num_of_chains=4, seed=0)
    with pytest.raises(Exception, match='finite'):
        next(sampler)


def test_r_hat_and_effective_sample_size():
    from flowket_b200.samplers import MetropolisHastingsLocal
    from flowket_b200.samplers.metropolis_hastings import sum_correlations
    table = TableMachine((4,), seed=0)
    sampler = MetropolisHastingsLocal(table, 8 * 500, num_of_chains=8, seed=0)
    rng = np.random.default_rng(0)
    iid = rng.normal(size=8 * 500)
    r_hat, variance, corr_sum, ess = sampler.calc_r_hat_value(iid)
    assert abs(r_hat - 1.0) < 0.01 and abs(variance - 1.0) < 0.1
    assert abs(corr_sum) < 0.2 and 0.7 * 4000 < ess < 1.4 * 4000
    # AR(1) chains with rho = 0.8: integrated autocorrelation sum rho / (1 - rho) = 4 -> ESS = n / 9
    ar = np.zeros((8, 500))
    noise = rng.normal(size=(8, 500))
    for t in range(1, 500):
        ar[:, t] = 0.8 * ar[:, t - 1] + noise[:, t]
    _, _, corr_sum_ar, ess_ar = sampler.calc_r_hat_value(ar.reshape(-1))
    assert 2.5 < corr_sum_ar < 5.5 and ess_ar < 0.2 * 4000
    # chains stuck at different levels: r_hat >> 1
    stuck = (np.arange(8)[:, None] + 0.01 * rng.normal(size=(8, 500))).reshape(-1)
    assert sampler.calc_r_hat_value(stuck)[0] > 10
    # one sample per chain: nothing to estimate (metropolis_hastings.py:75-77)
    single = MetropolisHastingsLocal(table, 8, num_of_chains=8, seed=0)
    assert single.calc_r_hat_value(rng.normal(size=8))[0] == 1.0
    assert sum_correlations(np.array([0.5, 0.2, -0.3, 0.4, 0.1])) == pytest.approx(0.7)
    assert sum_correlations(np.array([0.5, 0.2, 0.1])) == pytest.approx(0.8)


class FakeGenerator(object):
    """Publishes what VariationalMonteCarlo publishes (optimization/variational_monte_carlo.py:36-50)."""

    def __init__(self, energies, batch=None):
        self._energies = list(energies)
        self._i = -1
        self.current_batch = np.array([[1, 1, -1, -1], [1, 1, 1, -1]], dtype=np.float64) if batch is None else batch
        self.wave_function = lambda x: np.zeros((len(x), 1), np.complex64)
        self.sampler = None

    def __next__(self):
        self._i += 1
        e = self._energies[self._i % len(self._energies)]
        self.current_energy = complex(e, 0.25)
        self.current_local_energy_variance = 0.5 * abs(e)
        self.current_local_energy = np.full(8, e, np.complex128)
        self.start_time = time.time() - 3.0
        self.sampling_end_time = self.start_time + 1.0
        self.local_energy_end_time = self.start_time + 2.5
        return self.current_batch, np.zeros(2)


def test_monte_carlo_callbacks_fill_the_reference_log_keys():
    from flowket_b200.callbacks import default_wave_function_stats_callbacks_factory
    from flowket_b200.callbacks.monte_carlo import LocalEnergyStats, GeneratorIterator, RuntimeStats, ObservableStats
    gen, val = FakeGenerator([-10.0, -12.0]), FakeGenerator([-11.0])
    callbacks = default_wave_function_stats_callbacks_factory(gen, validation_generator=val,
                                                              true_ground_state_energy=-20.0, validation_period=2)
    assert [type(c) for c in callbacks] == [GeneratorIterator, LocalEnergyStats, ObservableStats, ObservableStats,
                                            RuntimeStats]
    next(gen)
    logs = {}
    for c in callbacks:
        c.on_batch_end(0, logs)
    assert logs['energy/energy'] == -10.0 and logs['energy/local_energy_variance'] == 5.0
    assert logs['energy/relative_error'] == pytest.approx(0.5)
    assert logs['observables/sigma_z'] == pytest.approx(0.25) and logs['observables/abs_sigma_z'] == pytest.approx(0.25)
    assert logs['times/sampling'] == pytest.approx(1.0) and logs['times/local_energy'] == pytest.approx(1.5)
    assert logs['times/gradients'] == pytest.approx(0.5, abs=0.2) and logs['times/total'] == pytest.approx(3.0, abs=0.2)
    assert not any(k.startswith('val_') for k in logs)
    for epoch in (0, 1):
        elogs = {}
        for c in callbacks:
            c.on_epoch_end(epoch, elogs)
        if epoch == 0:      # validation generator advanced by GeneratorIterator, then measured
            assert elogs['val_energy/energy'] == -11.0 and elogs['val_energy/relative_error'] == pytest.approx(0.45)
            assert 'val_observables/sigma_z' in elogs and 'energy/energy' not in elogs
        else:
            assert elogs == {}
    # epoch-mode logging
    per_epoch = LocalEnergyStats(gen, log_in_batch_or_epoch=False)
    logs = {}
    per_epoch.on_batch_end(0, logs)
    assert logs == {}
    per_epoch.on_epoch_end(0, logs)
    assert set(logs) == {'energy/energy', 'energy/local_energy_variance'}


def test_bad_eigen_state_stopping_and_mcmc_stats():
    from flowket_b200.callbacks.monte_carlo import BadEigenStateStopping, MCMCStats
    from flowket_b200.samplers import MetropolisHastingsLocal
    model = types.SimpleNamespace(stop_training=False)
    cb = BadEigenStateStopping(-100.0, variance_tol=1e-2, relative_error_to_stop=0.1, min_epoch=2)
    cb.set_model(model)
    stuck = {'energy/energy': -80.0, 'energy/local_energy_variance': 1e-4}
    cb.on_epoch_end(1, stuck)
    assert not model.stop_training                         # before min_epoch
    cb.on_epoch_end(2, {'energy/energy': -80.0, 'energy/local_energy_variance': 1.0})
    assert not model.stop_training                         # still fluctuating: keep training
    cb.on_epoch_end(3, {'energy/energy': -95.0, 'energy/local_energy_variance': 1e-4})
    assert not model.stop_training                         # close to the bound: fine
    cb.on_epoch_end(4, dict(stuck, **{'val_energy/energy': -99.0, 'val_energy/local_energy_variance': 1e-4}))
    assert not model.stop_training                         # validation stats win
    cb.on_epoch_end(5, stuck)
    assert model.stop_training and cb.stopped_epoch == 5
    with pytest.warns(RuntimeWarning):
        cb.on_epoch_end(6, {})
    gen = FakeGenerator([-1.0])
    gen.sampler = MetropolisHastingsLocal(TableMachine((4,)), 64, num_of_chains=4, seed=0)
    next(gen)
    gen.current_local_energy = np.random.default_rng(0).normal(size=64) + 0j
    logs = {}
    MCMCStats(gen).on_batch_end(0, logs)
    assert set(logs) == {'mcmc/acceptance_ratio', 'mcmc/energy_r_hat', 'mcmc/energy_effective_sample_size',
                         'mcmc/energy_correlations_sum'}


def test_evaluate_means_the_logs_over_steps():
    from flowket_b200.callbacks.monte_carlo import LocalEnergyStats
    from flowket_b200.evaluation import evaluate, mean_logs
    gen = FakeGenerator([-10.0, -12.0, -14.0])
    res = evaluate(gen, 3, [LocalEnergyStats(gen, true_ground_state_energy=-24.0)], verbose=False,
                   keys_to_progress_bar_mapping={'energy/energy': 'energy'})
    assert res['energy/energy'] == pytest.approx(-12.0) and res['energy/relative_error'] == pytest.approx(0.5)
    assert mean_logs([]) == {} and mean_logs([{'a': 1.0, 'b': 2.0}, {'a': 3.0, 'b': 0.0}], keys=['a']) == {'a': 2.0}


def test_exact_callbacks_and_exact_evaluate():
    from flowket_b200.callbacks.exact import default_wave_function_callbacks_factory, MachineUpdated, \
        ExactObservableCallback
    from flowket_b200.evaluation import exact_evaluate
    from flowket_b200.optimization import ExactVariational
    table = TableMachine((2, 2), seed=3)
    model = types.SimpleNamespace(input_shape=table.input_shape, predict=table.predict)
    op = oops.OracleOperator('ising', (2, 2), h=1.0, pbc=False)
    ev = ExactVariational(model, op, 4)
    assert ev.num_of_batch_until_full_cycle == 4
    callbacks = default_wave_function_callbacks_factory(ev, true_ground_state_energy=-5.0)
    logs = exact_evaluate(ev, callbacks)
    # <H> straight from the table: sum_s p(s) E_loc(s)
    states = ev.states
    conn, mel, use = op.find_conn(states)
    psi = table.predict
    ratio = np.exp(np.stack([psi(c)[:, 0] for c in conn]) - psi(states)[:, 0][None, :])
    want = ((mel * use * ratio).sum(axis=0) * table.probs).sum()
    assert logs['energy/energy'] == pytest.approx(want.real, rel=1e-10)
    assert logs['energy/relative_error'] == pytest.approx((-5.0 - want.real) / -5.0)
    m = states.reshape(16, -1).sum(axis=1) / 4.0
    assert logs['observables/sigma_z'] == pytest.approx((m * table.probs).sum())
    assert logs['observables/abs_sigma_z'] == pytest.approx((np.abs(m) * table.probs).sum())
    assert 'times/total' not in logs        # RuntimeStats reports on the last mini-batch of a cycle only
    last = {}
    callbacks[-1].on_batch_end(3, last)
    assert set(last) == {'times/wave_function_update', 'times/local_energy', 'times/gradients', 'times/total'}
    # per-batch callbacks are silent inside a cycle
    inside = {}
    callbacks[0].on_batch_end(1, inside)
    assert inside == {}
    # MachineUpdated re-reads the machine; an extra observable goes through its own ExactObservable
    calls = table.calls
    MachineUpdated(ev).on_batch_end(0)
    assert table.calls == calls + 4
    MachineUpdated(ev, update_in_batch_or_epoch=False).on_batch_end(0)
    assert table.calls == calls + 4
    extra = {}
    ExactObservableCallback(ev, oops.OracleOperator('ising', (2, 2), h=0.0, pbc=False), 'zz').on_batch_end(0, extra)
    zz = -sum(states[:, a, b] * states[:, c, d] for (a, b), (c, d) in
              [((0, 0), (0, 1)), ((1, 0), (1, 1)), ((0, 0), (1, 0)), ((0, 1), (1, 1))])
    assert extra['observables/zz'] == pytest.approx((zz * table.probs).sum())


def test_checkpoint_by_time_callback(tmp_path):
    from flowket_b200.callbacks import CheckpointByTime
    saved = []
    trainer = types.SimpleNamespace(save_checkpoint=lambda p: saved.append(p))
    cb = CheckpointByTime(str(tmp_path / 'ckpt_{epoch}'), save_frequency_in_minutes=1e9)
    cb.set_trainer(trainer)
    cb.on_epoch_begin(3)
    cb.on_batch_end(0, {})
    assert saved == []
    cb.save_frequency_in_minutes = 0.0
    cb.on_batch_end(1, {})
    assert saved == [str(tmp_path / 'ckpt_3')]
    cb.on_train_end()
    assert saved[-1] == str(tmp_path / 'ckpt_4') and cb.saves == 2
    weights = []
    cb2 = CheckpointByTime('w', save_weights_only=True)
    cb2.set_model(types.SimpleNamespace(save_weights=lambda p: weights.append(p)))
    cb2.on_train_end()
    assert weights == ['w']


class _StubMachine(object):
    def __init__(self, n=3):
        import torch
        self.params = torch.zeros(n, dtype=torch.float64)
        self.updates = 0

    def flat_params_device(self):
        return self.params

    def params_updated(self):
        self.updates += 1


def _stub_trainer(optimizer, gradients, distributed=False):
    """Trainer whose device gradient is replaced by a scripted sequence (host logic only)."""
    import torch
    from flowket_b200.optimizers import Trainer
    machine = _StubMachine()
    model = types.SimpleNamespace(machine=machine, stop_training=False)
    seq = iter(gradients)

    def generator():
        while True:
            yield np.zeros((2, 3)), np.zeros(2)

    trainer = Trainer(model, generator(), optimizer, distributed=distributed)
    trainer.gradient = lambda x, y: torch.tensor(next(seq), dtype=torch.float64)
    return trainer, machine


def test_accumulate_gradient_optimizer_sum_mean_and_ema(tmp_path):
    """accumulate_gradient_optimizer.py:54-82: step once per `update_params_frequency` mini-batches on the sum (or mean) of
    their gradients; EMA of the parameters after every step with the bias-corrected read-out (:28-29,66-68)."""
    from flowket_b200.optimizers import SGD, convert_to_accumulate_gradient_optimizer
    grads = [[1.0, 0, 0], [0, 2.0, 0], [0, 0, 4.0], [1.0, 1.0, 1.0], [1.0, 1.0, 1.0], [1.0, 1.0, 1.0]]
    opt = convert_to_accumulate_gradient_optimizer(SGD(lr=0.5), 3, accumulate_sum_or_mean=True, ema_decay=0.5)
    trainer, machine = _stub_trainer(opt, grads)
    moved = [trainer.train_on_batch(None, None) for _ in range(3)]
    assert moved == [False, False, True] and machine.updates == 1
    assert machine.params.tolist() == [-0.5, -1.0, -2.0]
    trainer.train_step()                                   # three more mini-batches, one update
    assert machine.params.tolist() == [-2.0, -2.5, -3.5] and opt.accumulated_iterations == 6
    # ema after two steps: e1 = 0.5 p1, e2 = 0.25 p1 + 0.5 p2; corrected by 1 - 0.5^2
    want = (0.25 * np.array([-0.5, -1.0, -2.0]) + 0.5 * np.array([-2.0, -2.5, -3.5])) / 0.75
    opt.set_weights_ema()
    assert np.allclose(machine.params.numpy(), want)
    # mean mode and a changed frequency
    opt2 = convert_to_accumulate_gradient_optimizer(SGD(lr=1.0), 2, accumulate_sum_or_mean=False)
    trainer2, machine2 = _stub_trainer(opt2, [[2.0, 0, 0], [0, 4.0, 0], [6.0, 6.0, 6.0]])
    trainer2.train_step()
    assert machine2.params.tolist() == [-1.0, -2.0, 0.0]
    opt2.set_update_params_frequency(1)
    trainer2.train_step()
    assert machine2.params.tolist() == [-7.0, -8.0, -6.0]
    with pytest.raises(ValueError):
        convert_to_accumulate_gradient_optimizer(SGD(), 0)
    with pytest.raises(RuntimeError):
        opt2.set_weights_ema(machine2)


def test_fit_generator_keras_semantics_and_scalar_logger(tmp_path):
    """steps_per_epoch counts mini-batches, callbacks see every mini-batch, epochs run initial_epoch..epochs"""
    import json
    from flowket_b200.callbacks import Callback, TensorBoard
    from flowket_b200.optimizers import SGD, convert_to_accumulate_gradient_optimizer
    opt = convert_to_accumulate_gradient_optimizer(SGD(lr=1.0), 2)
    trainer, machine = _stub_trainer(opt, [[1.0, 0, 0]] * 12)
    seen = []

    class Probe(Callback):
        def on_batch_end(self, batch, logs=None):
            seen.append(('batch', batch))
            logs['energy/energy'] = -float(len(seen))

        def on_epoch_end(self, epoch, logs=None):
            seen.append(('epoch', epoch))

    board = TensorBoard(log_dir=str(tmp_path / 'tb'), update_freq=2)
    logs = trainer.fit_generator(steps_per_epoch=4, epochs=3, initial_epoch=1, callbacks=[Probe(), board])
    assert [s for s in seen if s[0] == 'epoch'] == [('epoch', 1), ('epoch', 2)]
    assert [b for k, b in seen if k == 'batch'] == [0, 1, 2, 3, 0, 1, 2, 3]
    assert machine.updates == 4 and machine.params[0].item() == -8.0        # 8 mini-batches, 2 per update
    assert len(logs) == 2 and 'energy/energy' in logs[-1]
    lines = [json.loads(l) for l in open(tmp_path / 'tb' / 'scalars.jsonl')]
    assert [l['step'] for l in lines] == [2, 4, 6, 8] and all('energy/energy' in l for l in lines)


def test_sr_optimizers_own_the_update_and_terminate_on_nan(tmp_path):
    """model.compile(optimizer=<SR optimizer>) + fit_generator: the Trainer hands the batch to the optimizer
    (optimizers/stochastic_reconfiguration/optimizer.py:14-31 is a Keras optimizer in the reference)."""
    from flowket_b200.callbacks import TerminateOnNaN, Callback
    calls = []

    class ComplexSR(object):
        def compute_update(self, sigma, y):
            raise AssertionError('step() is the entry point')

        def complex_jacobian(self, sigma):
            pass

        def step(self, sigma, y):
            calls.append(('complex', np.asarray(y).copy()))

    class RealSR(object):
        def compute_update(self, sigma, e):
            pass

        def step(self, sigma, local_energy):
            calls.append(('real', np.asarray(local_energy).copy()))

    y = np.array([0.5 - 0.25j, -0.5 + 0.25j]) / 2                      # conj(E_loc - E) / B with B = 2
    for opt in (ComplexSR(), RealSR()):
        trainer, machine = _stub_trainer(opt, [])
        assert trainer.train_on_batch(np.zeros((2, 3)), y) is True
    assert calls[0][0] == 'complex' and np.allclose(calls[0][1], y)
    assert calls[1][0] == 'real' and np.allclose(calls[1][1], [0.5 + 0.25j, -0.5 - 0.25j])     # E_loc - E recovered
    trainer, _ = _stub_trainer(RealSR(), [])
    trainer.generator = types.SimpleNamespace(update_params_frequency=4)
    with pytest.raises(ValueError, match='mini_batch_size == batch_size'):
        trainer.train_on_batch(np.zeros((2, 3)), y)
    # TerminateOnNaN watches the energy entries of the logs
    model = types.SimpleNamespace(stop_training=False)
    cb = TerminateOnNaN()
    cb.set_model(model)
    cb.on_batch_end(0, {'energy/energy': -3.0})
    assert not model.stop_training
    cb.on_epoch_end(0, {'energy/energy': float('nan')})
    assert model.stop_training and cb.stopped


def test_save_weights_with_an_h5_name_round_trips(tmp_path):
    from flowket_b200 import Input, Model
    from flowket_b200.machines import ConvNetAutoregressive2D
    inp = Input(shape=(4, 4))
    m = ConvNetAutoregressive2D(inp, depth=2, num_of_channels=4, seed=1)
    model = Model(inp, m.predictions)
    path = str(tmp_path / 'final_run.h5')                              # the reference scripts save to '<name>.h5'
    model.save_weights(path)
    before = [w.copy() for w in model.get_weights()]
    m.set_weights([w * 0 for w in before])
    model.load_weights(path)
    assert all(np.array_equal(a, b) for a, b in zip(before, model.get_weights()))


@pytest.mark.parametrize('name', ['iid', 'ar1', 'stuck', 'short', 'single'])
def test_r_hat_matches_the_reference_implementation(name):
    """golden: MetropolisHastingsSampler.calc_r_hat_value of the reference itself (oracle/make_golden.py mcmc)"""
    import os
    from flowket_b200.samplers import MetropolisHastingsLocal
    from flowket_b200.samplers.metropolis_hastings import sum_correlations
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_mcmc_diagnostics.npz'))
    values, chains = g[name + '/values'], int(g[name + '/chains'])
    sampler = MetropolisHastingsLocal(TableMachine((4,)), len(values), num_of_chains=chains, seed=0)
    got = np.array([float(x) for x in sampler.calc_r_hat_value(values)])
    assert np.allclose(got, g[name + '/result'], rtol=1e-12, atol=1e-12, equal_nan=True), (got, g[name + '/result'])
    assert sum_correlations(g['sum_correlations/input']) == pytest.approx(float(g['sum_correlations/result']), rel=1e-13)


# ---- golden: the reference's own callbacks / evaluate on a scripted generator (oracle/make_golden.py callbacks) ------------
class ScriptedGenerator(object):
    """same script as oracle/make_golden.py::ScriptedGenerator (kept in sync by hand; the golden file is the arbiter)"""

    def __init__(self, energies, batch):
        self.energies, self.i = list(energies), -1
        self.current_batch = np.asarray(batch, dtype=np.float64)
        self.wave_function = lambda x: np.zeros((len(x), 1), np.complex64)
        self.sampler = None

    def __next__(self):
        self.i += 1
        e = self.energies[self.i % len(self.energies)]
        self.current_energy = complex(e, 0.125)
        self.current_local_energy_variance = 0.5 * abs(e)
        self.current_local_energy = np.full(8, e, np.complex128)
        self.start_time, self.sampling_end_time, self.local_energy_end_time = 100.0, 101.0, 102.5
        return self.current_batch, np.zeros(len(self.current_batch))


SCRIPT_BATCH = [[1, 1, -1, -1, 1, -1], [1, 1, 1, -1, 1, 1], [-1, -1, -1, -1, 1, -1]]
TIME_KEYS = ('times/gradients', 'times/total')


def _clean(logs):
    return {k: (float(np.real(v)) if k not in TIME_KEYS else None) for k, v in logs.items()}


def _same_logs(got, want):
    assert set(got) == set(want), (sorted(got), sorted(want))
    for k, v in want.items():
        if v is None:
            assert got[k] is None
        else:
            assert got[k] == pytest.approx(v, rel=1e-12, abs=1e-12), k


@pytest.fixture(scope='module')
def golden_logs():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_callbacks_logs.json')) as f:
        return json.load(f)


def test_monte_carlo_callbacks_reproduce_the_reference_logs(golden_logs):
    from flowket_b200.callbacks import default_wave_function_stats_callbacks_factory
    from flowket_b200.callbacks.monte_carlo import LocalEnergyStats, ObservableStats
    from flowket_b200.observables.monte_carlo import SigmaZ
    gen, val = ScriptedGenerator([-10.0, -12.0, -11.0], SCRIPT_BATCH), ScriptedGenerator([-9.0, -13.0], SCRIPT_BATCH[:2])
    callbacks = default_wave_function_stats_callbacks_factory(gen, validation_generator=val, true_ground_state_energy=-20.0,
                                                              validation_period=2)      # this repository's factory
    trace = iter(golden_logs['monte_carlo_batch_mode'])
    for epoch in range(3):
        for batch in range(2):
            next(gen)
            logs = {}
            for c in callbacks:
                c.on_batch_end(batch, logs)
            kind, e, b, want = next(trace)
            assert (kind, e, b) == ('batch', epoch, batch)
            _same_logs(_clean(logs), want)
        logs = {}
        for c in callbacks:
            c.on_epoch_end(epoch, logs)
        kind, e, _, want = next(trace)
        assert (kind, e) == ('epoch', epoch)
        _same_logs(_clean(logs), want)
    gen = ScriptedGenerator([-10.0, -12.0], SCRIPT_BATCH)
    callbacks = [LocalEnergyStats(gen, log_in_batch_or_epoch=False), ObservableStats(gen, SigmaZ(), 'sigma_z', log_in_batch_or_epoch=False)]
    next(gen)
    b, e = {}, {}
    for c in callbacks:
        c.on_batch_end(0, b)
        c.on_epoch_end(0, e)
    _same_logs(_clean(b), golden_logs['monte_carlo_epoch_mode'][0])
    _same_logs(_clean(e), golden_logs['monte_carlo_epoch_mode'][1])


def test_evaluate_and_bad_eigen_state_stopping_reproduce_the_reference(golden_logs):
    from flowket_b200.callbacks.monte_carlo import LocalEnergyStats, ObservableStats, BadEigenStateStopping
    from flowket_b200.evaluation import evaluate
    from flowket_b200.observables.monte_carlo import AbsSigmaZ
    gen = ScriptedGenerator([-10.0, -12.0, -14.0, -11.0], SCRIPT_BATCH)
    res = evaluate(gen, 4, [LocalEnergyStats(gen, true_ground_state_energy=-24.0), ObservableStats(gen, AbsSigmaZ(), 'abs_sigma_z')],
                   verbose=False)
    _same_logs(_clean(res), golden_logs['evaluate'])
    cb = BadEigenStateStopping(-100.0, variance_tol=1e-2, relative_error_to_stop=0.1, min_epoch=2)
    cb.set_model(types.SimpleNamespace(stop_training=False))
    for epoch, logs, stop, stopped_epoch in golden_logs['bad_eigen_state_stopping']:
        cb.on_epoch_end(epoch, logs)
        assert (bool(cb.model.stop_training), cb.stopped_epoch) == (stop, stopped_epoch), epoch


def test_exact_callbacks_reproduce_the_reference_logs(golden_logs):
    from flowket_b200.callbacks.exact import default_wave_function_callbacks_factory
    from flowket_b200.evaluation import exact_evaluate
    from flowket_b200.optimization import ExactVariational
    vec = np.array([complex(a, b) for a, b in golden_logs['exact_log_psi_vector']])
    f = vector_to_machine(vec)
    model = types.SimpleNamespace(input_shape=(None, 2, 3), predict=lambda x, batch_size=None: f(np.asarray(x)))
    ev = ExactVariational(model, oops.OracleOperator('ising', (2, 3), pbc=False, h=1.5), 16)
    callbacks = default_wave_function_callbacks_factory(ev, true_ground_state_energy=-12.0)
    _same_logs(_clean(exact_evaluate(ev, callbacks)), golden_logs['exact_evaluate'])
    for batch, want_keys in enumerate(golden_logs['exact_batch_gating']):
        logs = {}
        for c in callbacks:
            c.on_batch_end(batch, logs)
        assert sorted(logs) == want_keys, batch
