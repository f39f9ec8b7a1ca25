"""GPU tier, runs last: checks added at the end of round 1 after the GPU budget was spent, i.e. NOT yet run on a B200 (every
earlier GPU test file was).  Ordered from the most to the least certain, because the driver runs with -x:
  1. the CUDA engines against numbers produced by the reference's OWN machine classes, loss gradient, sampler loop and
     symmetrisation ensembles (tests/golden/reference_{machines,autoregressive_sampler,ensembles}.npz -- oracle/make_golden.py
     runs the reference's code on oracle/tf_standin.py; the CPU tier checks the oracle against the same files);
  2. the device find_conn against the reference's output over ALL states of degenerate lattices
     (tests/golden/reference_numpy_half_edge.npz);
  3. the BASELINE configs[3] composition through compile()/fit_generator, and local energies on degenerate lattices."""
import numpy as np
import pytest
import torch

from oracle import local_energy as oeloc
from oracle import nets, operators as oops
from tests.helpers import make_pair, random_sigma

pytestmark = pytest.mark.gpu

EDGE = {
    'heis_2x2_pbc': ('Heisenberg', dict(hilbert_state_shape=[2, 2], pbc=True)),
    'heis_1x6_obc': ('Heisenberg', dict(hilbert_state_shape=[1, 6], pbc=False)),
    'heis_6x1_pbc': ('Heisenberg', dict(hilbert_state_shape=[6, 1], pbc=True)),
    'heis_2_pbc': ('Heisenberg', dict(hilbert_state_shape=[2], pbc=True)),
    'heis_3x2_pbc_norot': ('Heisenberg', dict(hilbert_state_shape=[3, 2], pbc=True, unitary_rotation=False)),
    'heis_2x5_pbc': ('Heisenberg', dict(hilbert_state_shape=[2, 5], pbc=True)),
    'ising_2x3_pbc': ('Ising', dict(hilbert_state_shape=[2, 3], pbc=True, h=0.5, j=1.0)),
    'ising_1x4_pbc': ('Ising', dict(hilbert_state_shape=[1, 4], pbc=True, h=1.0)),
    'ising_2_pbc': ('Ising', dict(hilbert_state_shape=[2], pbc=True, h=1.0)),
    'ising_2x2_obc': ('Ising', dict(hilbert_state_shape=[2, 2], pbc=False, h=2.0, j=0.5)),
}


MACHINE_GOLDEN = {
    # name -> (product machine kind, shape, depth, channels, constructor kwargs)
    'conv2d_4x3_d3_wn': ('conv2d', (4, 3), 3, 8, {}),
    'conv2d_3x4_d2_plain': ('conv2d', (3, 4), 2, 6, {'weights_normalization': False}),
    'conv1d_12_d5_dil4_skip': ('conv1d', (12,), 5, 8, {'max_dilation_rate': 4, 'add_skip_connections': True}),
    'conv1d_10_d4_plain': ('conv1d', (10,), 4, 6, {'weights_normalization': False}),
    'cconv1d_10_d4_dil2': ('cconv1d', (10,), 4, 6, {'max_dilation_rate': 2}),
    'cconv1d_8_d3': ('cconv1d', (8,), 3, 4, {}),
}


@pytest.mark.parametrize('name', sorted(MACHINE_GOLDEN))
def test_device_wave_function_matches_the_reference_machine_classes(name):
    """golden tests/golden/reference_machines.npz: log psi and conditional log-probabilities computed by the reference's OWN
    machine classes (oracle/make_golden.py machines); the CUDA fp32 engine gets the same weights and spins (1e-5)."""
    import os
    import torch
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_machines.npz'))
    kind, shape, depth, channels, kw = MACHINE_GOLDEN[name]
    model, cond_model, spec, _ = make_pair(kind, shape, depth, channels, seed=0, **kw)
    params = nets.unflatten_params(spec, torch.from_numpy(g[name + '/params']))
    model.machine.set_weights([p.numpy().astype(np.float32) for p in params])
    sigma = g[name + '/sigma']
    got = model.predict(sigma)[:, 0]
    want = g[name + '/log_psi']
    assert np.abs(got.real - want.real).max() < 1e-5 * max(1.0, np.abs(want.real).max())
    assert np.abs(np.exp(1j * got.imag) - np.exp(1j * want.imag)).max() < 1e-4           # phases modulo 2 pi, fp32
    got_c = cond_model.predict(sigma)
    want_c = g[name + '/conditional_log_probs']
    assert np.abs(got_c - want_c).max() < 1e-5 * max(1.0, np.abs(want_c).max())


@pytest.mark.parametrize('name', sorted(MACHINE_GOLDEN))
def test_device_gradients_match_autograd_through_the_reference_forward(name):
    """golden: the reference's loss differentiated through the reference machine's own forward (make_golden.py machines);
    the hand-written CUDA backward (fk_grad_weighted, fk_grad_per_sample) gets the same weights, spins and coefficients."""
    import os
    import torch
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_machines.npz'))
    kind, shape, depth, channels, kw = MACHINE_GOLDEN[name]
    model, _, spec, _ = make_pair(kind, shape, depth, channels, seed=0, **kw)
    params = nets.unflatten_params(spec, torch.from_numpy(g[name + '/params']))
    model.machine.set_weights([p.numpy().astype(np.float32) for p in params])
    net = model.machine.device_net()
    sigma = g[name + '/sigma']
    y = g[name + '/y'].astype(np.complex64)
    got = net.grad_weighted(net.to_sigma(sigma), torch.from_numpy(y)).cpu().numpy()
    want = g[name + '/weighted_gradient']
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 2e-5
    O_re, O_im = net.grad_per_sample(net.to_sigma(sigma[3:4]), imag=True)
    want_re, want_im = g[name + '/jacobian_row3_real'], g[name + '/jacobian_row3_imag']
    assert np.linalg.norm(O_re.cpu().numpy()[0] - want_re) / np.linalg.norm(want_re) < 2e-5
    assert np.linalg.norm(O_im.cpu().numpy()[0] - want_im) / max(np.linalg.norm(want_im), 1e-30) < 2e-5


def test_tensor_core_engine_matches_the_reference_machine_class():
    """the 32-channel machine of tests/golden/reference_machines.npz through the tcgen05 engines: log psi within the
    tensor-core tolerance of tests/test_gpu_tc.py (|d| <= 0.05 + 2e-3 |log psi|), weighted gradient within 2e-2 in norm --
    against numbers from the reference's own machine code, not only against this repository's fp32 engine"""
    import os
    import torch
    from flowket_b200 import FK_ENGINE_TC, FK_ENGINE_FP32
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_machines.npz'))
    name = 'conv2d_6x6_d3_c32'
    model, _, spec, _ = make_pair('conv2d', (6, 6), 3, 32, seed=0)
    params = nets.unflatten_params(spec, torch.from_numpy(g[name + '/params']))
    model.machine.set_weights([p.numpy().astype(np.float32) for p in params])
    sigma, want = g[name + '/sigma'], g[name + '/log_psi']
    got32 = model.predict(sigma)[:, 0]
    assert np.abs(got32.real - want.real).max() < 1e-5 * max(1.0, np.abs(want.real).max())
    model.engine = FK_ENGINE_TC
    got = model.predict(sigma)[:, 0]
    assert np.all(np.abs(got.real - want.real) <= 0.05 + 2e-3 * np.abs(want.real))
    assert np.abs(np.exp(1j * got.imag) - np.exp(1j * want.imag)).max() < 0.05
    net = model.machine.device_net()
    y = torch.from_numpy(g[name + '/y'].astype(np.complex64))
    want_g = g[name + '/weighted_gradient']
    g32 = net.grad_weighted(net.to_sigma(sigma), y, engine=FK_ENGINE_FP32).cpu().numpy().astype(np.float64)
    assert np.linalg.norm(g32 - want_g) / np.linalg.norm(want_g) < 2e-5
    gtc = net.grad_weighted(net.to_sigma(sigma), y, engine=FK_ENGINE_TC).cpu().numpy().astype(np.float64)
    assert np.linalg.norm(gtc - want_g) / np.linalg.norm(want_g) < 2e-2


@pytest.mark.parametrize('name,kind,shape,depth,channels,kw', [
    ('conv2d_4x3', 'conv2d', (4, 3), 2, 8, {}),
    ('conv1d_10', 'conv1d', (10,), 4, 8, {'max_dilation_rate': 2}),
    ('cconv1d_8', 'cconv1d', (8,), 3, 4, {'max_dilation_rate': 2}),
])
def test_device_samplers_reproduce_the_reference_samplers_spins(name, kind, shape, depth, channels, kw):
    """golden (oracle/make_golden.py sampler): spins drawn by the reference's own AutoregressiveSampler.__next__ around the
    oracle network; the CUDA samplers get the same weights and the same uniforms.  A spin may differ only at a numerical
    tie |p0 - u| < 1e-5 (fp32 summation order), as in tests/test_gpu_parity.py."""
    import os
    import torch
    from flowket_b200.samplers import FastAutoregressiveSampler, AutoregressiveSampler
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_autoregressive_sampler.npz'))
    model, cond_model, spec, _ = make_pair(kind, shape, depth, channels, seed=0, **kw)
    params = nets.unflatten_params(spec, torch.from_numpy(g[name + '/params']))
    model.machine.set_weights([p.numpy() for p in params])
    u, want = g[name + '/uniforms'], g[name + '/sigma']
    p0 = np.exp(nets.conditional_log_probs(spec, [p.double() for p in params], want).numpy()[..., 0])
    B = len(u)
    for sampler in (FastAutoregressiveSampler(cond_model, B), AutoregressiveSampler(cond_model, B)):
        got = sampler.next_device(uniforms=u).cpu().numpy()
        bad = np.argwhere(got != want)
        for idx in bad:   # the first differing site of a sample must be a numerical tie
            first = tuple(bad[bad[:, 0] == idx[0]][0])
            assert abs(p0[first] - u[first]) < 1e-5, (type(sampler).__name__, first, p0[first], u[first])


@pytest.mark.parametrize('name,kind,shape,depth,channels,kw', [
    ('conv2d_4x3', 'conv2d', (4, 3), 2, 8, {}),
    ('conv2d_3x3_d3', 'conv2d', (3, 3), 3, 4, {}),
    ('conv1d_10', 'conv1d', (10,), 4, 8, {'max_dilation_rate': 2, 'add_skip_connections': True}),
    ('cconv1d_8', 'cconv1d', (8,), 3, 4, {'max_dilation_rate': 2}),
])
def test_device_samplers_reproduce_the_reference_fast_samplers_spins(name, kind, shape, depth, channels, kw):
    """golden tests/golden/reference_fast_sampler.npz: spins drawn by the reference's OWN FastAutoregressiveSampler (dependency
    graph + layer topologies, make_golden.py fast_sampler); the CUDA incremental samplers get the same weights and uniforms.
    A spin may differ only at a numerical tie |p0 - u| < 1e-5."""
    import os
    from flowket_b200.samplers import FastAutoregressiveSampler
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_fast_sampler.npz'))
    model, cond_model, spec, _ = make_pair(kind, shape, depth, channels, seed=0, **kw)
    params = nets.unflatten_params(spec, torch.from_numpy(g[name + '/params']))
    model.machine.set_weights([p.numpy().astype(np.float32) for p in params])
    u, want = g[name + '/uniforms'], g[name + '/sigma']
    p0 = np.exp(nets.conditional_log_probs(spec, params, want).numpy()[..., 0])
    got = FastAutoregressiveSampler(cond_model, len(u)).next_device(uniforms=u).cpu().numpy()
    bad = np.argwhere(got != want)
    for idx in bad:   # the first differing site of a sample must be a numerical tie
        first = tuple(bad[bad[:, 0] == idx[0]][0])
        assert abs(p0[first] - u[first]) < 1e-5, (first, p0[first], u[first])


def test_device_ensembles_match_the_reference_ensembles():
    """golden tests/golden/reference_ensembles.npz: the reference's own symmetrisation ensembles around its own 2-D machine;
    the product's EnsembleModel (device route) gets the same weights and spins."""
    import os
    import torch
    from flowket_b200.machines import make_2d_obc_invariants, make_up_down_invariant, make_pbc_invariants
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_ensembles.npz'))
    model, _, spec, _ = make_pair('conv2d', (4, 4), 2, 8, seed=0)
    params = nets.unflatten_params(spec, torch.from_numpy(g['params']))
    model.machine.set_weights([p.numpy().astype(np.float32) for p in params])
    inp = model.input
    sigma = g['sigma']

    def same(got, want):
        got = np.asarray(got)[:, 0]
        assert np.abs(got.real - want.real).max() < 1e-5 * max(1.0, np.abs(want.real).max())
        assert np.abs(np.exp(1j * got.imag) - np.exp(1j * want.imag)).max() < 1e-4

    same(model.predict(sigma), g['base'])
    same(make_2d_obc_invariants(inp, model).predict(sigma), g['obc'])
    same(make_2d_obc_invariants(inp, model, probabilistic=False).predict(sigma), g['obc_average'])
    same(make_up_down_invariant(inp, make_2d_obc_invariants(inp, model)).predict(sigma), g['up_down_of_obc'])
    same(make_pbc_invariants(inp, model, apply_also_obc_invariants=False).predict(sigma), g['translations'])


def test_device_complex_sr_update_matches_the_reference_pipeline():
    """golden tests/golden/reference_complex_sr_pipeline.npz: the SR update of the complex 1-D machine computed end to end by
    the reference's own code (machine -> per-sample Jacobian -> complex assembly -> S, F -> delta -> new weights); the device
    optimizer (fk_grad_per_sample, fk_sr_gram, Cholesky) gets the same weights, spins and targets."""
    import os
    from flowket_b200.optimizers import ComplexValuesStochasticReconfiguration
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_complex_sr_pipeline.npz'))
    model, _, spec, _ = make_pair('cconv1d', (9,), 3, 4, seed=0, max_dilation_rate=2)
    params = nets.unflatten_params(spec, torch.from_numpy(g['params']))
    model.machine.set_weights([p.numpy().astype(np.float32) for p in params])
    sr = ComplexValuesStochasticReconfiguration(model, lr=float(g['lr']), diag_shift=float(g['diag_shift']), iterative_solver=False)
    O = sr.complex_jacobian(g['sigma']).cpu().numpy()
    assert np.linalg.norm(O - g['jacobian']) / np.linalg.norm(g['jacobian']) < 2e-5
    delta = sr.compute_update(g['sigma'], g['y_true']).cpu().numpy()
    assert np.linalg.norm(delta - g['delta']) / np.linalg.norm(g['delta']) < 5e-4
    sr.apply_complex_gradient(torch.from_numpy(g['delta']).to(model.machine.flat_params_device().device))
    new = model.machine.flat_params_device().cpu().numpy()
    assert np.abs(new - g['new_params']).max() < 1e-6


@pytest.mark.parametrize('name', sorted(EDGE))
def test_device_find_conn_on_degenerate_lattices(golden_edge, name):
    import flowket_b200.operators as ops
    cls, kw = EDGE[name]
    op = getattr(ops, cls)(**kw)
    conn, mel, use = op.find_conn(golden_edge[name + '/sigma'])
    assert op.max_number_of_local_connections == int(golden_edge[name + '/max_conn'])
    assert np.array_equal(conn.astype(np.int8), golden_edge[name + '/conn'])
    assert np.array_equal(use, golden_edge[name + '/use'])
    assert np.allclose(mel, golden_edge[name + '/mel'], rtol=1e-6, atol=1e-6)


def test_sr_optimizer_through_compile_and_flattened_operator():
    """BASELINE configs[3] composition: J1J2 lattice operator seen through raster-flattened configurations by the complex
    1-D machine, ComplexValuesStochasticReconfiguration passed to compile() like any Keras optimizer."""
    from flowket_b200 import Input, Model
    from flowket_b200.callbacks import TerminateOnNaN
    from flowket_b200.callbacks.monte_carlo import default_wave_function_stats_callbacks_factory
    from flowket_b200.machines import ComplexValuesSimpleConvNetAutoregressive1D
    from flowket_b200.operators import J1J2, FlattenedOperator
    from flowket_b200.optimization import VariationalMonteCarlo, loss_for_energy_minimization
    from flowket_b200.optimizers import ComplexValuesStochasticReconfiguration
    from flowket_b200.samplers import FastAutoregressiveSampler
    lattice = J1J2(hilbert_state_shape=[4, 4], j2=0.5, pbc=False)
    operator = FlattenedOperator(lattice)
    sigma = random_sigma(9, (4, 4), seed=3)
    conn, mel, use = lattice.find_conn(sigma)
    fconn, fmel, fuse = operator.find_conn(sigma.reshape(9, 16))
    assert np.array_equal(fconn.reshape(conn.shape), conn) and np.array_equal(fmel, mel) and np.array_equal(fuse, use)
    inputs = Input(shape=(16,), dtype='int8')
    convnet = ComplexValuesSimpleConvNetAutoregressive1D(inputs, depth=3, num_of_channels=8, max_dilation_rate=4, seed=0)
    model = Model(inputs=inputs, outputs=convnet.predictions)
    cond = Model(inputs=inputs, outputs=convnet.conditional_log_probs)
    optimizer = ComplexValuesStochasticReconfiguration(model, lr=0.02, diag_shift=0.05, iterative_solver=False)
    model.compile(optimizer=optimizer, loss=loss_for_energy_minimization)
    vmc = VariationalMonteCarlo(model, operator, FastAutoregressiveSampler(cond, 512, seed=2))
    # the flattened operator gives the same local energies as the oracle on the lattice
    x, _ = vmc.next_batch()
    oop = oops.OracleOperator('j1j2', (4, 4), j2=0.5, pbc=False)
    spec = nets.ComplexConv1DSpec(16, 3, 8, max_dilation_rate=4)
    params = [torch.from_numpy(w.astype(np.float64)) for w in convnet.get_weights()]
    want = oeloc.local_values(oop, lambda c: nets.log_psi_numpy(spec, params, np.asarray(c).reshape(len(c), 16)),
                              x[:32].reshape(32, 4, 4).astype(np.float64))
    assert np.abs(vmc.current_local_energy[:32] - want).max() / np.abs(want).max() < 1e-4
    before = convnet.flat_params_device().clone()
    callbacks = default_wave_function_stats_callbacks_factory(vmc, log_in_batch_or_epoch=False) + [TerminateOnNaN()]
    logs = model.fit_generator(vmc.to_generator(), steps_per_epoch=5, epochs=4, callbacks=callbacks, max_queue_size=0, workers=0)
    assert len(logs) == 4 and all(np.isfinite(l['energy/energy']) for l in logs)
    assert (convnet.flat_params_device() - before).abs().max().item() > 0          # SR moved the parameters
    assert np.mean([l['energy/energy'] for l in logs[2:]]) < logs[0]['energy/energy'] + 5.0   # and not uphill (512-sample noise)


@pytest.mark.parametrize('kind,shape,opkind,opkw', [
    ('conv2d', (2, 2), 'heisenberg', dict(pbc=True)),
    ('conv2d', (2, 5), 'heisenberg', dict(pbc=True)),
    ('conv1d', (2,), 'ising', dict(pbc=True, h=1.0)),
    ('conv2d', (1, 4), 'ising', dict(pbc=True, h=1.0)),
])
def test_local_energy_on_degenerate_lattices(kind, shape, opkind, opkw):
    """every state of the lattice: E_loc from the device pipeline == oracle (fp32 engine, 1e-5)"""
    from flowket_b200.exact.utils import decimal_array_to_binary_array
    from flowket_b200.observables.monte_carlo import Observable
    from tests.test_gpu_parity import _product_operator
    model, _, spec, params = make_pair(kind, shape, 3, 8, seed=21)
    n = int(np.prod(shape))
    sigma = decimal_array_to_binary_array(np.arange(2 ** n), n, False).reshape((2 ** n,) + shape).astype(np.int8)
    got = Observable(_product_operator(opkind, shape, opkw)).local_values_device(model, sigma).cpu().numpy()
    want = oeloc.local_values(oops.OracleOperator(opkind, shape, **opkw), lambda c: nets.log_psi_numpy(spec, params, c),
                              sigma.astype(np.float64))
    assert np.abs(got - want).max() / np.abs(want).max() < 1e-5
