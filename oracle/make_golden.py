"""Generate tests/golden/*.npz by running the REFERENCE's own numpy code (read-only /root/reference).

TEST INFRASTRUCTURE ONLY.  Run once in the build container (`python -m oracle.make_golden`);
the vectors are committed because /root/reference does not exist on the GPU box.

The reference's root package imports TensorFlow (src/flowket/__init__.py:1-3), which is not installed;
registering stub parent packages with the right __path__ lets the TF-free sub-modules load unmodified
(SURVEY.md appendix C).  Nothing is copied from the reference: only its outputs are stored.
"""
import importlib
import os
import sys
import types

import numpy as np

REF = '/root/reference/src/flowket'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def load_reference():
    def stub(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    stub('flowket', REF)
    stub('flowket.observables', REF + '/observables')
    ops = importlib.import_module('flowket.operators')
    mc = importlib.import_module('flowket.observables.monte_carlo')
    ex = importlib.import_module('flowket.exact.utils')
    return ops, mc, ex


def main():
    ops, mc, ex = load_reference()
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20261017)
    out = {}

    # ---- find_conn fixtures --------------------------------------------------------------
    cases = [
        ('heis_2d_obc', lambda: ops.Heisenberg(hilbert_state_shape=[4, 5], pbc=False), (4, 5)),
        ('heis_2d_pbc', lambda: ops.Heisenberg(hilbert_state_shape=[4, 4], pbc=True), (4, 4)),
        ('heis_2d_obc_norot', lambda: ops.Heisenberg(hilbert_state_shape=[3, 4], pbc=False, unitary_rotation=False), (3, 4)),
        ('heis_1d_pbc', lambda: ops.Heisenberg(hilbert_state_shape=[7], pbc=True), (7,)),
        ('heis_1d_obc', lambda: ops.Heisenberg(hilbert_state_shape=[6], pbc=False), (6,)),
        ('heis_2d_10x10_obc', lambda: ops.Heisenberg(hilbert_state_shape=[10, 10], pbc=False), (10, 10)),
        ('ising_2d_obc', lambda: ops.Ising(hilbert_state_shape=[4, 4], pbc=False, h=3.0), (4, 4)),
        ('ising_2d_pbc', lambda: ops.Ising(hilbert_state_shape=[3, 5], pbc=True, h=0.7, j=1.3), (3, 5)),
        ('ising_1d_obc', lambda: ops.Ising(hilbert_state_shape=[9], pbc=False, h=3.0), (9,)),
        ('ising_1d_pbc', lambda: ops.Ising(hilbert_state_shape=[8], pbc=True, h=2.0), (8,)),
    ]
    for name, make, shape in cases:
        B = 6 if name != 'heis_2d_10x10_obc' else 3
        sigma = rng.choice([-1, 1], size=(B,) + shape).astype(np.int8)
        op = make()
        conn, mel, use = op.find_conn(sigma.astype(np.float64) if name.startswith('heis') else sigma.astype(np.int64))
        out[name + '/sigma'] = sigma
        out[name + '/conn'] = np.asarray(conn).astype(np.int8)
        out[name + '/mel'] = np.asarray(mel, np.float64)
        out[name + '/use'] = np.asarray(use, bool)
        out[name + '/max_conn'] = np.int64(op.max_number_of_local_connections)

    # ---- local energy with a synthetic wave function (vector_to_machine) --------------------
    for name, make, shape in [
        ('eloc_heis_2d_obc', lambda: ops.Heisenberg(hilbert_state_shape=[3, 4], pbc=False), (3, 4)),
        ('eloc_heis_1d_pbc', lambda: ops.Heisenberg(hilbert_state_shape=[7], pbc=True), (7,)),
        ('eloc_ising_2d_obc', lambda: ops.Ising(hilbert_state_shape=[3, 4], pbc=False, h=3.0), (3, 4)),
    ]:
        n = int(np.prod(shape))
        vec = (rng.normal(size=2 ** n) * 0.3 + 1j * rng.normal(size=2 ** n) * 0.5).astype(np.complex64)
        psi = ex.vector_to_machine(vec)
        sigma = rng.choice([-1, 1], size=(16,) + shape).astype(np.int8)
        obs = mc.Observable(make())
        lv = obs.local_values(psi, sigma.astype(np.float64))
        out[name + '/sigma'] = sigma
        out[name + '/log_psi_vector'] = vec
        out[name + '/local_values'] = np.asarray(lv, np.complex128)

    # hand-written fixture of tests/test_variational.py:22-34 (balanced == unbalanced) evaluated by the reference
    sample = np.array([[1, 1, 1, -1, -1, -1, -1], [1, 1, 1, -1, 1, -1, -1], [1, -1, 1, 1, -1, -1, -1]])
    local_connections = rng.choice([-1, 1], size=(5, 3, 7))
    local_connections[0] = sample
    hamiltonian_values = np.array([[2.0, 7j + 8, 0.0, 0.0, 3], [0.0, 0.0, 0.0, 0.0, -1.0], [5.0, 3j, 0.0, -2, 9]]).T
    all_use_conn = np.array([[True, True, False, False, True], [True, False, False, False, True],
                             [True, True, False, True, True]]).T
    vec = (rng.normal(size=2 ** 7) * 0.3 + 1j * rng.normal(size=2 ** 7) * 0.5).astype(np.complex64)
    obs = mc.Observable(ops.Heisenberg(hilbert_state_shape=(7,)))
    obs.operator.hilbert_state_shape = (7,)
    unb = obs.local_values_optimized_for_unbalanced_local_connections(ex.vector_to_machine(vec), local_connections,
                                                                      hamiltonian_values, all_use_conn)
    bal = obs.local_values_optimized_for_balanced_local_connections(ex.vector_to_machine(vec), local_connections,
                                                                    hamiltonian_values)
    out['handmade/local_connections'] = local_connections.astype(np.int8)
    out['handmade/hamiltonian_values'] = hamiltonian_values
    out['handmade/all_use_conn'] = all_use_conn
    out['handmade/log_psi_vector'] = vec
    out['handmade/unbalanced'] = np.asarray(unb)
    out['handmade/balanced'] = np.asarray(bal)

    # ---- bit conventions -------------------------------------------------------------------
    dec = np.arange(32)
    out['bits/binary'] = ex.decimal_array_to_binary_array(dec.copy(), 5, False).astype(np.int8)
    out['bits/decimal'] = ex.binary_array_to_decimal_array(out['bits/binary'].astype(np.int64)).astype(np.int64)

    # ---- exact-diagonalisation anchors embedded in the reference's scripts --------------------
    out['ed/ising_4x4_obc_h3'] = np.float64(-50.18662388277671)        # examples/basic_autoregressive_2d.py:39
    out['ed/ising_1d16_obc_h3'] = np.float64(-49.257706531889006)      # examples/basic_autoregressive_exact_gradient.py:38
    out['ed/j1j2_4x4_obc_j2_0.5'] = np.float64(-30.022227800323677)    # examples/j1j2_2d_monte_carlo_4.py:43
    out['ed/heis_1d20_pbc'] = np.float64(-35.6175461195)               # examples/complex_ops_autoregressive_heisenberg_1d.py:57

    np.savez_compressed(os.path.join(OUT, 'reference_numpy_half.npz'), **out)
    print('wrote', os.path.join(OUT, 'reference_numpy_half.npz'), len(out), 'arrays')


EDGE_CASES = [
    # degenerate lattices: two-site periodic dimensions (both bonds of a pair exist), singleton dimensions, tiny chains
    ('heis_2x2_pbc', 'Heisenberg', dict(hilbert_state_shape=[2, 2], pbc=True)),
    ('heis_1x6_obc', 'Heisenberg', dict(hilbert_state_shape=[1, 6], pbc=False)),
    ('heis_6x1_pbc', 'Heisenberg', dict(hilbert_state_shape=[6, 1], pbc=True)),
    ('heis_2_pbc', 'Heisenberg', dict(hilbert_state_shape=[2], pbc=True)),
    ('heis_3x2_pbc_norot', 'Heisenberg', dict(hilbert_state_shape=[3, 2], pbc=True, unitary_rotation=False)),
    ('heis_2x5_pbc', 'Heisenberg', dict(hilbert_state_shape=[2, 5], pbc=True)),
    ('ising_2x3_pbc', 'Ising', dict(hilbert_state_shape=[2, 3], pbc=True, h=0.5, j=1.0)),
    ('ising_1x4_pbc', 'Ising', dict(hilbert_state_shape=[1, 4], pbc=True, h=1.0)),
    ('ising_2_pbc', 'Ising', dict(hilbert_state_shape=[2], pbc=True, h=1.0)),
    ('ising_2x2_obc', 'Ising', dict(hilbert_state_shape=[2, 2], pbc=False, h=2.0, j=0.5)),
]


def edge_cases():
    """tests/golden/reference_numpy_half_edge.npz: find_conn of the reference on degenerate lattices, ALL 2^N states of
    each (they are tiny), so that every boundary branch of operators/heisenberg.py:96-120 and ising.py:19-41 is pinned."""
    ops, mc, ex = load_reference()
    out = {}
    for name, cls, kw in EDGE_CASES:
        shape = tuple(kw['hilbert_state_shape'])
        n = int(np.prod(shape))
        sigma = ex.decimal_array_to_binary_array(np.arange(2 ** n), n, False).reshape((2 ** n,) + shape).astype(np.int8)
        op = getattr(ops, cls)(**kw)
        conn, mel, use = op.find_conn(sigma.astype(np.float64) if cls == 'Heisenberg' else sigma.astype(np.int64))
        out[name + '/sigma'] = sigma
        out[name + '/conn'] = np.array(conn).astype(np.int8).reshape((len(conn), 2 ** n) + shape)
        out[name + '/mel'] = np.array(mel, np.float64)
        out[name + '/use'] = np.array(use, bool)
        out[name + '/max_conn'] = np.int64(op.max_number_of_local_connections)
    path = os.path.join(OUT, 'reference_numpy_half_edge.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, len(out), 'arrays')


def mcmc_diagnostics():
    """tests/golden/reference_mcmc_diagnostics.npz: MetropolisHastingsSampler.calc_r_hat_value of the reference
    (samplers/metropolis_hastings.py:66-91 -- numpy only, loaded through stub parent packages) on fixed inputs."""
    import contextlib
    import io

    def stub(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    stub('flowket', REF)
    stub('flowket.deepar', REF + '/deepar')
    stub('flowket.deepar.samplers', REF + '/deepar/samplers')
    stub('flowket.samplers', REF + '/samplers')
    mh = importlib.import_module('flowket.samplers.metropolis_hastings')
    if not hasattr(np, 'bool'):
        np.bool = bool                       # alias the reference still uses (removed in numpy 1.24)

    class Machine(object):
        input_shape = (None, 4)

        def predict(self, x, batch_size=None):
            return np.zeros((len(x), 1), np.complex128)

    rng = np.random.default_rng(20261017)
    out = {}
    for name, chains, per_chain in [('iid', 8, 200), ('ar1', 4, 300), ('stuck', 6, 50), ('short', 16, 2), ('single', 8, 1)]:
        if name == 'ar1':
            v = np.zeros((chains, per_chain))
            noise = rng.normal(size=(chains, per_chain))
            for t in range(1, per_chain):
                v[:, t] = 0.8 * v[:, t - 1] + noise[:, t]
        elif name == 'stuck':
            v = np.arange(chains)[:, None] + 0.01 * rng.normal(size=(chains, per_chain))
        else:
            v = rng.normal(size=(chains, per_chain))
        with contextlib.redirect_stdout(io.StringIO()):
            sampler = mh.MetropolisHastingsLocal(Machine(), chains * per_chain, num_of_chains=chains)
            res = sampler.calc_r_hat_value(v.reshape(-1).copy())
        out[name + '/values'] = v.reshape(-1)
        out[name + '/chains'] = np.int64(chains)
        out[name + '/result'] = np.array([float(x) for x in res])
    out['sum_correlations/input'] = rng.normal(size=12) * 0.4
    out['sum_correlations/result'] = np.float64(mh.sum_correlations(out['sum_correlations/input']))
    path = os.path.join(OUT, 'reference_mcmc_diagnostics.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, len(out), 'arrays')


def j1j2_graph():
    """tests/golden/reference_j1j2_graph.npz: the coloured edge list and the bond operators the reference hands to netket
    (operators/j1j2.py:7-55), captured by running the reference's own builder against a *recording stand-in* for the
    netket package (netket itself is not installed; its internal connection ordering stays unpinned -- this pins everything
    FlowKet itself decides: which bonds exist, their colours and order, J1 / J2 and the two-site operator)."""
    class Recorder(object):
        def __init__(self, *args, **kwargs):
            self.args, self.kwargs = args, kwargs

    fake = types.ModuleType('netket')
    fake.graph = types.SimpleNamespace(CustomGraph=Recorder)
    fake.hilbert = types.SimpleNamespace(Spin=Recorder)
    fake.operator = types.SimpleNamespace(GraphOperator=Recorder)
    sys.modules['netket'] = fake
    m = types.ModuleType('flowket')
    m.__path__ = [REF]
    sys.modules['flowket'] = m
    j1j2 = importlib.import_module('flowket.operators.j1j2')
    out = {}
    for L1, L2, j2, pbc in [(4, 4, 0.5, False), (6, 6, 0.5, False), (4, 4, 0.5, True), (2, 3, 0.3, False), (3, 3, 0.5, True),
                            (3, 5, 0.25, False)]:
        op = j1j2.j1j2_two_dim_netket_operator((L1, L2), j2=j2, pbc=pbc)
        hilbert = op.args[0]
        graph = hilbert.kwargs['graph']
        key = '%dx%d_%s_j2_%g' % (L1, L2, 'pbc' if pbc else 'obc', j2)
        out[key + '/edges'] = np.array(graph.args[0], dtype=np.int64)                       # [n_edges, 3] = (a, b, colour)
        out[key + '/bondops'] = np.array(op.kwargs['bondops'], dtype=np.complex128)         # [2, 4, 4]
        out[key + '/bondops_colors'] = np.array(op.kwargs['bondops_colors'], dtype=np.int64)
        out[key + '/total_sz_constrained'] = np.bool_('total_sz' in hilbert.kwargs)
    path = os.path.join(OUT, 'reference_j1j2_graph.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, len(out), 'arrays')
    del sys.modules['netket']


def exact_variational():
    """tests/golden/reference_exact_variational.npz: the reference's own ExactVariational / ExactObservable
    (optimization/exact_variational.py:10-161) run on a log-amplitude table.  The module imports tensorflow only for
    `K.function` (the wave-function callable) and the default-graph guard; a minimal stand-in for those two lets the numpy
    logic -- normalisation, probabilities, local energies over all states, gradient coefficients -- run unmodified."""
    import contextlib

    class Graph(object):
        def as_default(self):
            return contextlib.nullcontext()

    backend = types.ModuleType('tensorflow.keras.backend')
    backend.function = lambda inputs, outputs: (lambda xs: [outputs[0](xs[0])])
    keras = types.ModuleType('tensorflow.keras')
    keras.backend = backend
    tf = types.ModuleType('tensorflow')
    tf.keras = keras
    tf.get_default_graph = lambda: Graph()
    sys.modules.update({'tensorflow': tf, 'tensorflow.keras': keras, 'tensorflow.keras.backend': backend})
    ops, mc, ex = load_reference()
    m = types.ModuleType('flowket.optimization')
    m.__path__ = [REF + '/optimization']
    sys.modules['flowket.optimization'] = m
    ev_mod = importlib.import_module('flowket.optimization.exact_variational')
    rng = np.random.default_rng(20261017)
    out = {}
    for name, make, shape, batch in [
        ('heis_2x3_obc', lambda: ops.Heisenberg(hilbert_state_shape=[2, 3], pbc=False), (2, 3), 16),
        ('ising_3x3_obc', lambda: ops.Ising(hilbert_state_shape=[3, 3], pbc=False, h=3.0), (3, 3), 128),
        ('ising_1d8_pbc', lambda: ops.Ising(hilbert_state_shape=[8], pbc=True, h=0.7), (8,), 256),
    ]:
        n = int(np.prod(shape))
        vec = rng.normal(size=2 ** n) * 0.6 + 1j * rng.uniform(-np.pi, np.pi, size=2 ** n)
        table = ex.vector_to_machine(vec)
        model = types.SimpleNamespace(input=None, output=lambda x: table(x), input_shape=(None,) + shape)
        ev = ev_mod.ExactVariational(model, make(), batch)
        ev.machine_updated()
        out[name + '/log_psi_vector'] = vec
        out[name + '/batch_size'] = np.int64(batch)
        out[name + '/probs'] = ev.probs.copy()
        out[name + '/energies'] = ev.energy_observable.energies.copy()
        out[name + '/energy_grad_coefficients'] = ev.energy_grad_coefficients.copy()
        out[name + '/current_energy'] = np.complex128(ev.energy_observable.current_energy)
        out[name + '/current_local_energy_variance'] = np.float64(ev.energy_observable.current_local_energy_variance)
        out[name + '/num_of_batch_until_full_cycle'] = np.int64(ev.num_of_batch_until_full_cycle)
    path = os.path.join(OUT, 'reference_exact_variational.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, len(out), 'arrays')
    for k in ('tensorflow', 'tensorflow.keras', 'tensorflow.keras.backend'):
        del sys.modules[k]


def variational_monte_carlo():
    """tests/golden/reference_variational_monte_carlo.npz: the reference's own VariationalMonteCarlo + MiniBatchGenerator
    (optimization/variational_monte_carlo.py:12-50, mini_batch_generator.py:5-45) driven by a scripted sampler and a
    log-amplitude table; tensorflow is replaced by a stand-in for the session / graph guards it only passes through."""
    import contextlib

    class Graph(object):
        def as_default(self):
            return contextlib.nullcontext()

    backend = types.ModuleType('tensorflow.python.keras.backend')
    backend.get_session = lambda: None
    backend.set_session = lambda session: None
    tf = types.ModuleType('tensorflow')
    tf.get_default_graph = lambda: Graph()
    tf_python = types.ModuleType('tensorflow.python')
    tf_python_keras = types.ModuleType('tensorflow.python.keras')
    tf_python_keras.backend = backend
    tf_python.keras = tf_python_keras
    tf.python = tf_python
    sys.modules.update({'tensorflow': tf, 'tensorflow.python': tf_python, 'tensorflow.python.keras': tf_python_keras,
                        'tensorflow.python.keras.backend': backend})
    ops, mc, ex = load_reference()
    m = types.ModuleType('flowket.optimization')
    m.__path__ = [REF + '/optimization']
    sys.modules['flowket.optimization'] = m
    vmc_mod = importlib.import_module('flowket.optimization.variational_monte_carlo')
    rng = np.random.default_rng(20261018)
    shape = (3, 4)
    n = 12
    vec = rng.normal(size=2 ** n) * 0.4 + 1j * rng.uniform(-np.pi, np.pi, size=2 ** n)
    table = ex.vector_to_machine(vec)
    batches = rng.choice([-1, 1], size=(3, 10) + shape).astype(np.float64)

    class ScriptedSampler(object):
        batch_size = 10

        def __init__(self):
            self.i = -1

        def __next__(self):
            self.i += 1
            return batches[self.i % len(batches)]

    model = types.SimpleNamespace(predict=lambda x, batch_size=None: table(x))
    vmc = vmc_mod.VariationalMonteCarlo(model, ops.Heisenberg(hilbert_state_shape=list(shape), pbc=False), ScriptedSampler(),
                                        mini_batch_size=4)
    out = {'log_psi_vector': vec, 'batches': batches.astype(np.int8),
           'update_params_frequency': np.int64(vmc.update_params_frequency)}
    xs, ys, energies, variances = [], [], [], []
    for step in range(7):                       # 10 samples in windows of 4: two windows per batch, the tail of 2 is dropped
        x, y = next(vmc)
        xs.append(np.asarray(x).astype(np.int8))
        ys.append(np.asarray(y, np.complex128))
        energies.append(np.complex128(vmc.current_energy))
        variances.append(np.float64(vmc.current_local_energy_variance))
    out['mini_batches_x'] = np.stack(xs)
    out['mini_batches_y'] = np.stack(ys)
    out['energies'] = np.array(energies)
    out['variances'] = np.array(variances)
    out['last_local_energy'] = np.asarray(vmc.current_local_energy, np.complex128)
    out['batches_drawn'] = np.int64(vmc.sampler.i + 1)
    path = os.path.join(OUT, 'reference_variational_monte_carlo.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, len(out), 'arrays')
    for k in ('tensorflow', 'tensorflow.python', 'tensorflow.python.keras', 'tensorflow.python.keras.backend'):
        del sys.modules[k]


class ScriptedGenerator(object):
    """what VariationalMonteCarlo publishes after a batch (optimization/variational_monte_carlo.py:36-50), scripted:
    energies follow a fixed list, the batch is fixed, the time stamps are fixed offsets (shared with the tests)"""

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
TIME_KEYS = ('times/gradients', 'times/total')     # depend on the wall clock at the moment of the call


def callbacks_logs():
    """tests/golden/reference_callbacks_logs.json: the `logs` dicts the reference's own stats callbacks and `evaluate` /
    `exact_evaluate` produce for a scripted generator and for an ExactVariational over a log-amplitude table
    (callbacks/monte_carlo/*.py, callbacks/exact/*.py, evaluation/evaluate.py); keras' Callback base class and the
    tensorflow guards are replaced by minimal stand-ins, everything else is the reference's code."""
    import contextlib
    import json

    class Callback(object):
        def __init__(self, **kwargs):
            self.model = None

        def on_batch_end(self, batch, logs=None):
            pass

        def on_epoch_end(self, epoch, logs=None):
            pass

    class Graph(object):
        def as_default(self):
            return contextlib.nullcontext()

    def module(name, **attrs):
        mod = types.ModuleType(name)
        mod.__dict__.update(attrs)
        sys.modules[name] = mod
        return mod
    backend = module('tensorflow.keras.backend', function=lambda inputs, outputs: (lambda xs: [outputs[0](xs[0])]))
    cbs = module('tensorflow.keras.callbacks', Callback=Callback)
    keras = module('tensorflow.keras', backend=backend, callbacks=cbs)
    module('tensorflow', keras=keras, get_default_graph=lambda: Graph())
    ops, mc, ex = load_reference()
    for name in ('optimization', 'callbacks', 'callbacks/monte_carlo', 'callbacks/exact', 'evaluation'):
        pkg = types.ModuleType('flowket.' + name.replace('/', '.'))
        pkg.__path__ = [REF + '/' + name]
        sys.modules[pkg.__name__] = pkg
    imp = importlib.import_module
    les = imp('flowket.callbacks.monte_carlo.local_energy_stats').LocalEnergyStats
    obs = imp('flowket.callbacks.monte_carlo.observable').ObservableStats
    rts = imp('flowket.callbacks.monte_carlo.runtime_stats').RuntimeStats
    git = imp('flowket.callbacks.monte_carlo.generator_iterator').GeneratorIterator
    bad = imp('flowket.callbacks.monte_carlo.bad_eigen_state_stopping').BadEigenStateStopping
    evaluate_mod = imp('flowket.evaluation.evaluate')
    ev_mod = imp('flowket.optimization.exact_variational')

    def clean(logs):
        return {k: (float(np.real(v)) if k not in TIME_KEYS else None) for k, v in logs.items()}
    out = {}
    # --- Monte-Carlo stats, batch mode with a validation generator every 2nd epoch (the factory's list, built by hand
    #     because the factory module imports the TensorFlow event writer)
    gen, val = ScriptedGenerator([-10.0, -12.0, -11.0], SCRIPT_BATCH), ScriptedGenerator([-9.0, -13.0], SCRIPT_BATCH[:2])
    shared = dict(validation_generator=val, log_in_batch_or_epoch=True, validation_period=2)
    callbacks = [git(val, period=2), les(gen, true_ground_state_energy=-20.0, **shared),
                 obs(gen, mc.SigmaZ(), 'sigma_z', **shared), obs(gen, mc.AbsSigmaZ(), 'abs_sigma_z', **shared),
                 rts(gen, log_in_batch_or_epoch=True)]
    trace = []
    for epoch in range(3):
        for batch in range(2):
            next(gen)
            logs = {}
            for c in callbacks:
                c.on_batch_end(batch, logs)
            trace.append(['batch', epoch, batch, clean(logs)])
        logs = {}
        for c in callbacks:
            c.on_epoch_end(epoch, logs)
        trace.append(['epoch', epoch, -1, clean(logs)])
    out['monte_carlo_batch_mode'] = trace
    # --- epoch mode
    gen = ScriptedGenerator([-10.0, -12.0], SCRIPT_BATCH)
    callbacks = [les(gen, log_in_batch_or_epoch=False), obs(gen, mc.SigmaZ(), 'sigma_z', log_in_batch_or_epoch=False)]
    next(gen)
    b, e = {}, {}
    for c in callbacks:
        c.on_batch_end(0, b)
        c.on_epoch_end(0, e)
    out['monte_carlo_epoch_mode'] = [clean(b), clean(e)]
    # --- evaluate(): mean of the logs over steps
    gen = ScriptedGenerator([-10.0, -12.0, -14.0, -11.0], SCRIPT_BATCH)
    res = evaluate_mod.evaluate(gen, 4, [les(gen, true_ground_state_energy=-24.0), obs(gen, mc.AbsSigmaZ(), 'abs_sigma_z')],
                                verbose=False)
    out['evaluate'] = clean(res)
    # --- BadEigenStateStopping decisions
    decisions = []
    cb = bad(-100.0, variance_tol=1e-2, relative_error_to_stop=0.1, min_epoch=2)
    cb.model = types.SimpleNamespace(stop_training=False)
    for epoch, logs in enumerate([{'energy/energy': -80.0, 'energy/local_energy_variance': 1e-4},
                                  {'energy/energy': -80.0, 'energy/local_energy_variance': 1e-4},
                                  {'energy/energy': -80.0, 'energy/local_energy_variance': 1.0},
                                  {'energy/energy': -95.0, 'energy/local_energy_variance': 1e-4},
                                  {'energy/energy': -80.0, 'energy/local_energy_variance': 1e-4,
                                   'val_energy/energy': -99.0, 'val_energy/local_energy_variance': 1e-4},
                                  {'energy/energy': -80.0, 'energy/local_energy_variance': 1e-4}]):
        cb.on_epoch_end(epoch, logs)
        decisions.append([epoch, logs, bool(cb.model.stop_training), cb.stopped_epoch])
    out['bad_eigen_state_stopping'] = decisions
    # --- exact callbacks + exact_evaluate on a table machine
    exl = imp('flowket.callbacks.exact.local_energy').ExactLocalEnergy
    exs = imp('flowket.callbacks.exact.sigma_z').ExactSigmaZ
    exr = imp('flowket.callbacks.exact.runtime_stats').RuntimeStats
    rng = np.random.default_rng(20261019)
    shape = (2, 3)
    vec = rng.normal(size=64) * 0.5 + 1j * rng.uniform(-np.pi, np.pi, size=64)
    table = ex.vector_to_machine(vec)
    model = types.SimpleNamespace(input=None, output=lambda x: table(x), input_shape=(None,) + shape)
    ev = ev_mod.ExactVariational(model, ops.Ising(hilbert_state_shape=list(shape), pbc=False, h=1.5), 16)
    callbacks = [exl(ev, true_ground_state_energy=-12.0), exs(exact_variational=ev), exr(ev)]
    out['exact_log_psi_vector'] = [[float(z.real), float(z.imag)] for z in vec]
    out['exact_evaluate'] = clean(evaluate_mod.exact_evaluate(ev, callbacks))
    gating = []
    for batch in range(9):                      # 4 mini-batches per cycle: who reports on which batch index
        logs = {}
        for c in callbacks:
            c.on_batch_end(batch, logs)
        gating.append(sorted(logs))
    out['exact_batch_gating'] = gating
    path = os.path.join(OUT, 'reference_callbacks_logs.json')
    with open(path, 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print('wrote', path)
    for k in ('tensorflow', 'tensorflow.keras', 'tensorflow.keras.backend', 'tensorflow.keras.callbacks'):
        del sys.modules[k]


def autoregressive_sampler():
    """tests/golden/reference_autoregressive_sampler.npz: spins drawn by the reference's own AutoregressiveSampler.__next__
    (deepar/samplers/autoregressive.py:29-48, +-1 variant) when its `conditional_log_probs_machine.predict` is the oracle's
    fp32 network (oracle/nets.py) -- pins the sampling rule itself (raster order, `exp(log p0) > u`, unsampled sites = 0,
    spins written back as +-1) against the reference's code, given the uniforms numpy's global generator produced."""
    import torch
    from oracle import nets

    def stub(name, path):
        mod = types.ModuleType(name)
        mod.__path__ = [path]
        sys.modules[name] = mod
    stub('flowket', REF)
    stub('flowket.deepar', REF + '/deepar')
    stub('flowket.deepar.samplers', REF + '/deepar/samplers')
    ar = importlib.import_module('flowket.deepar.samplers.autoregressive')
    out = {}
    for name, spec, B, seed in [
        ('conv2d_4x3', nets.Conv2DSpec(4, 3, 2, 8), 64, 11),
        ('conv1d_10', nets.Conv1DSpec(10, 4, 8, max_dilation_rate=2), 48, 12),
        ('cconv1d_8', nets.ComplexConv1DSpec(8, 3, 4, max_dilation_rate=2), 32, 13),
    ]:
        params = nets.init_params(spec, seed=seed, dtype=torch.float32, bias_scale=0.3)

        class Machine(object):
            input_shape = (None,) + tuple(spec.input_shape)

            def predict(self, batch, batch_size=None):
                with torch.no_grad():
                    return nets.conditional_log_probs(spec, params, np.asarray(batch)).numpy()

        np.random.seed(seed)
        sampler = ar.AutoregressiveSampler(Machine(), B, zero_base=False)
        sigma = next(sampler)
        np.random.seed(seed)
        uniforms = np.random.rand(*((B,) + tuple(spec.input_shape)))       # the first draw of __next__ after seeding
        out[name + '/params'] = nets.flatten_params(params).numpy()
        out[name + '/uniforms'] = uniforms
        out[name + '/sigma'] = np.asarray(sigma).astype(np.int8)
    path = os.path.join(OUT, 'reference_autoregressive_sampler.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, len(out), 'arrays')


def complex_ops():
    """tests/golden/reference_complex_ops.npz: the reference's own `lncosh`, `complex_log`, `angle`
    (layers/complex/tensorflow_ops.py:69-85) and `probabilistic_ensemble_op` / `average_ensemble_op`
    (machines/ensemble.py:14-25) evaluated on numpy arrays.  These functions are compositions of a dozen elementwise
    TensorFlow math functions; the stand-in below maps each of those one-to-one onto the numpy function of the same
    meaning (real, imag, abs, exp, log, atan2, complex, cast, zeros_like, reduce_mean, reduce_logsumexp), so what is
    pinned is the reference's *composition* -- the numerically stable lncosh, the circular-mean phase, the -log(K)/2
    normalisation of the probabilistic ensemble."""
    import contextlib
    from scipy.special import logsumexp

    def module(name, **attrs):
        mod = types.ModuleType(name)
        mod.__dict__.update(attrs)
        sys.modules[name] = mod
        return mod
    tmath = types.SimpleNamespace(
        real=np.real, imag=np.imag, abs=np.abs, exp=np.exp, log=np.log, atan2=np.arctan2,
        reduce_mean=lambda x, axis=None, keepdims=False: np.mean(x, axis=axis, keepdims=keepdims),
        reduce_logsumexp=lambda x, axis=None, keepdims=False: logsumexp(x, axis=axis, keepdims=keepdims))
    backend = module('tensorflow.keras.backend')
    models = module('tensorflow.keras.models', Model=object)
    layers = module('tensorflow.keras.layers', Lambda=object, Input=object, Concatenate=object)
    keras = module('tensorflow.keras', backend=backend, models=models, layers=layers)
    module('tensorflow', math=tmath, keras=keras, complex=lambda re, im: np.asarray(re) + 1j * np.asarray(im),
           cast=lambda x, dtype: np.asarray(x).astype(dtype), zeros_like=np.zeros_like,
           name_scope=lambda *a, **k: contextlib.nullcontext('scope'))
    pkg = types.ModuleType('flowket')
    pkg.__path__ = [REF]
    sys.modules['flowket'] = pkg
    module('flowket.layers', Rot90=object, FlipLeftRight=object, Roll=object).__path__ = [REF + '/layers']
    cpx = types.ModuleType('flowket.layers.complex')
    cpx.__path__ = [REF + '/layers/complex']
    sys.modules['flowket.layers.complex'] = cpx
    mach = types.ModuleType('flowket.machines')
    mach.__path__ = [REF + '/machines']
    sys.modules['flowket.machines'] = mach
    tops = importlib.import_module('flowket.layers.complex.tensorflow_ops')
    ens = importlib.import_module('flowket.machines.ensemble')
    rng = np.random.default_rng(20261020)
    z = np.concatenate([np.array([2, 3j, 1 + 7j, 10 - 3j, -6, 0.0, 40 + 2j, -55 - 1j, 1e-3 + 1e-3j], dtype=np.complex128),
                        rng.normal(size=64) * 4 + 1j * rng.normal(size=64) * 4])
    x = rng.normal(size=(9, 16)) * 3 + 1j * rng.uniform(-np.pi, np.pi, size=(9, 16))
    out = {'z': z, 'lncosh': tops.lncosh(z), 'complex_log': tops.complex_log(z[z != 0]), 'angle': tops.angle(z),
           'ensemble_input': x,
           'probabilistic_ensemble': ens.probabilistic_ensemble_op(x, 16),
           'average_ensemble': ens.average_ensemble_op(x)}
    path = os.path.join(OUT, 'reference_complex_ops.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, len(out), 'arrays')
    for k in [k for k in sys.modules if k == 'tensorflow' or k.startswith('tensorflow.')]:
        del sys.modules[k]


def load_reference_machines():
    """import the reference's machine classes with oracle/tf_standin.py in place of tensorflow (see that file's header)"""
    from oracle import tf_standin
    tf_standin.install()

    def package(name, path):
        mod = types.ModuleType(name)
        mod.__path__ = [path]
        sys.modules[name] = mod
        return mod

    def module(name, **attrs):
        mod = types.ModuleType(name)
        mod.__dict__.update(attrs)
        sys.modules[name] = mod
        return mod
    package('flowket', REF)
    package('flowket.machines', REF + '/machines')
    package('flowket.optimization', REF + '/optimization')
    layers = package('flowket.layers', REF + '/layers')
    package('flowket.layers.complex', REF + '/layers/complex')
    package('flowket.deepar', REF + '/deepar')
    deepar_layers = package('flowket.deepar.layers', REF + '/deepar/layers')
    # graph analysis (the fast sampler's topology registry) and the weight initialisers are not part of the forward pass

    class TopologyManager(object):
        def register_layer_topology(self, *args, **kwargs):
            pass
    ga = module('flowket.deepar.graph_analysis', TopologyManager=TopologyManager, OneToOneTopology=object)
    module('flowket.deepar.graph_analysis.convolutional_topology', TopologyManager=TopologyManager, ConvolutionalTopology=object)
    module('flowket.deepar.graph_analysis.dependency_graph', assert_valid_probabilistic_model=lambda model: None)
    ga.__path__ = []

    class ConjugateDecorator(object):
        def __init__(self, initializer):
            pass

        def get_real_part_initializer(self):
            return None

        def get_imag_part_initializer(self):
            return None
    module('flowket.layers.complex.initializers', ConjugateDecorator=ConjugateDecorator, get=lambda identifier: identifier)
    imp = importlib.import_module
    for sub, names in [('autoregressive', ['NormalizeInLogSpace', 'CombineAutoregressiveConditionals']),
                       ('casting', ['CastingLayer', 'ToFloat32', 'ToFloat64']),
                       ('lambda_with_one_to_one_topology', ['LambdaWithOneToOneTopology']),
                       ('masking', ['DownShiftLayer', 'RightShiftLayer']),
                       ('one_hot', ['ToOneHot', 'PlusMinusOneToOneHot']),
                       ('padding', ['ExpandInputDim', 'PeriodicPadding']),
                       ('wrappers', ['WeightNormalization'])]:
        mod = imp('flowket.deepar.layers.' + sub)
        for n in names:
            setattr(deepar_layers, n, getattr(mod, n))
    casting = imp('flowket.layers.complex.casting')
    conv = imp('flowket.layers.complex.conv')
    spins = imp('flowket.layers.spins_invariants')
    for mod, names in [(casting, ['VectorToComplexNumber', 'ToComplex64', 'ToComplex128']), (conv, ['ComplexConv1D']),
                       (spins, ['EqualUpDownSpins'])]:
        for n in names:
            setattr(layers, n, getattr(mod, n))
    return (imp('flowket.machines.conv_net_autoregressive_2D').ConvNetAutoregressive2D,
            imp('flowket.machines.simple_conv_net_autoregressive_1D').SimpleConvNetAutoregressive1D,
            imp('flowket.machines.complex_values_simple_conv_net_autoregressive_1D').ComplexValuesSimpleConvNetAutoregressive1D)


MACHINE_CASES = [
    # name, kind, input shape, constructor kwargs of the REFERENCE class (== oracle spec arguments)
    ('conv2d_4x3_d3_wn', 'conv2d', (4, 3), dict(depth=3, num_of_channels=8)),
    ('conv2d_3x4_d2_plain', 'conv2d', (3, 4), dict(depth=2, num_of_channels=6, weights_normalization=False)),
    ('conv2d_5x5_d4_linear_g', 'conv2d', (5, 5), dict(depth=4, num_of_channels=4, exponential_norm=False)),
    ('conv1d_12_d5_dil4_skip', 'conv1d', (12,), dict(depth=5, num_of_channels=8, max_dilation_rate=4, add_skip_connections=True)),
    ('conv1d_10_d4_plain', 'conv1d', (10,), dict(depth=4, num_of_channels=6, weights_normalization=False)),
    ('conv1d_9_d6_nodil', 'conv1d', (9,), dict(depth=6, num_of_channels=4, use_dilation=False, max_dilation_rate=4)),
    ('cconv1d_10_d4_dil2', 'cconv1d', (10,), dict(depth=4, num_of_channels=6, max_dilation_rate=2)),
    ('cconv1d_8_d3', 'cconv1d', (8,), dict(depth=3, num_of_channels=4)),
    # the width the tensor-core engines are built for (32 channels)
    ('conv2d_6x6_d3_c32', 'conv2d', (6, 6), dict(depth=3, num_of_channels=32)),
]


def machines():
    """tests/golden/reference_machines.npz: log psi and the conditional log-probabilities computed by the reference's OWN
    machine classes for random weights and random spins (8 machines covering weight norm on / off / linear g, dilation
    rules, skip connections, the complex machine).  The oracle network (oracle/nets.py), until now pinned only by property
    tests and by the pretrained 2-D weights, is checked against these numbers on the CPU and the CUDA engines on the GPU."""
    import torch
    from oracle import nets, tf_standin
    classes = dict(zip(('conv2d', 'conv1d', 'cconv1d'), load_reference_machines()))
    rng = np.random.default_rng(20261021)
    out = {}
    for name, kind, shape, kw in MACHINE_CASES:
        if kind == 'conv2d':
            spec = nets.Conv2DSpec(shape[0], shape[1], kw['depth'], kw['num_of_channels'],
                                   weights_normalization=kw.get('weights_normalization', True),
                                   exponential_norm=kw.get('exponential_norm', True))
        elif kind == 'conv1d':
            spec = nets.Conv1DSpec(shape[0], kw['depth'], kw['num_of_channels'], use_dilation=kw.get('use_dilation', True),
                                   add_skip_connections=kw.get('add_skip_connections', False),
                                   max_dilation_rate=kw.get('max_dilation_rate'),
                                   weights_normalization=kw.get('weights_normalization', True))
        else:
            spec = nets.ComplexConv1DSpec(shape[0], kw['depth'], kw['num_of_channels'],
                                          max_dilation_rate=kw.get('max_dilation_rate'))
        # shapes and order of the weights come from the oracle's layout; the reference's add_weight calls must ask for
        # exactly these shapes in exactly this order (asserted inside the stand-in) -- that pins the parameter layout too
        params = [p + 0.3 * torch.randn(p.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(len(out) + i))
                  * (p.dim() == 1) for i, p in enumerate(nets.init_params(spec, seed=7, dtype=torch.float64))]
        sigma = rng.choice([-1, 1], size=(12,) + shape).astype(np.int8)
        tf_standin.inject_weights([p.numpy() for p in params])
        machine = classes[kind](torch.from_numpy(sigma.astype(np.float64)), **kw)
        assert not tf_standin._WEIGHTS, 'the reference created fewer weights than the oracle layout holds'
        out[name + '/params'] = nets.flatten_params(params).numpy()
        out[name + '/sigma'] = sigma
        out[name + '/log_psi'] = machine.predictions.numpy()[:, 0]
        out[name + '/conditional_log_probs'] = machine.conditional_log_probs.numpy()
        out[name + '/weight_names'] = np.array([n for n, _ in tf_standin.created_weights()])
        # gradient of the reference's loss through the reference's forward (autograd over the stand-in's torch ops):
        # loss_for_energy_minimization (optimization/loss.py:4-5) summed over the batch, and one per-sample Jacobian row
        loss_fn = importlib.import_module('flowket.optimization.loss').loss_for_energy_minimization
        y = torch.from_numpy(rng.normal(size=12) + 1j * rng.normal(size=12))
        leaves = tf_standin.inject_weights([p.numpy() for p in params], requires_grad=True)
        machine = classes[kind](torch.from_numpy(sigma.astype(np.float64)), **kw)
        loss = loss_fn(y.reshape(-1, 1), machine.predictions).sum()
        grads = torch.autograd.grad(loss, leaves, retain_graph=True)
        out[name + '/y'] = y.numpy()
        out[name + '/weighted_gradient'] = torch.cat([g.reshape(-1) for g in grads]).numpy()
        row = torch.autograd.grad(machine.predictions[3, 0].real, leaves, retain_graph=True)
        out[name + '/jacobian_row3_real'] = torch.cat([g.reshape(-1) for g in row]).numpy()
        row = torch.autograd.grad(machine.predictions[3, 0].imag, leaves)
        out[name + '/jacobian_row3_imag'] = torch.cat([g.reshape(-1) for g in row]).numpy()
    path = os.path.join(OUT, 'reference_machines.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, len(out), 'arrays')


def ensembles():
    """tests/golden/reference_ensembles.npz: the reference's own symmetrisation ensembles (machines/ensemble.py:28-69 with
    layers/dihedral_4_invariants.py and transition_invariants.py) around the reference's own 2-D machine, run eagerly on the
    stand-in: D4 (probabilistic and averaged), spin flip nested around D4 -- the composition of the published evaluation,
    experiments/run_evaluation.py:17-22 -- and all lattice translations.  A Keras sub-model reused on several inputs is, in
    eager mode, the same function called several times: `base` below re-injects the same weights for every call."""
    import torch
    from oracle import nets, tf_standin
    conv2d_cls = load_reference_machines()[0]
    package = lambda name, path: sys.modules.setdefault(name, types.ModuleType(name))     # noqa: E731
    layers = sys.modules['flowket.layers']
    imp = importlib.import_module
    d4, tr = imp('flowket.layers.dihedral_4_invariants'), imp('flowket.layers.transition_invariants')
    layers.Rot90, layers.FlipLeftRight, layers.Roll = d4.Rot90, d4.FlipLeftRight, tr.Roll
    ens = imp('flowket.machines.ensemble')
    spec = nets.Conv2DSpec(4, 4, 2, 8)
    kw = dict(depth=2, num_of_channels=8)
    params = [p + 0.3 * torch.randn(p.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(100 + i)) * (p.dim() == 1)
              for i, p in enumerate(nets.init_params(spec, seed=9, dtype=torch.float64))]

    def base(x):
        tf_standin.inject_weights([p.numpy() for p in params])
        return conv2d_cls(x, **kw).predictions

    rng = np.random.default_rng(20261022)
    sigma = rng.choice([-1, 1], size=(10, 4, 4)).astype(np.int8)
    x = torch.from_numpy(sigma.astype(np.float64))
    out = {'params': nets.flatten_params(params).numpy(), 'sigma': sigma, 'base': base(x).numpy()[:, 0]}
    out['obc'] = ens.make_2d_obc_invariants(x, base).output.numpy()[:, 0]
    out['obc_average'] = ens.make_2d_obc_invariants(x, base, probabilistic=False).output.numpy()[:, 0]
    obc = lambda inputs: ens.make_2d_obc_invariants(inputs, base).output                       # noqa: E731
    out['up_down_of_obc'] = ens.make_up_down_invariant(x, obc).output.numpy()[:, 0]
    out['translations'] = ens.make_pbc_invariants(x, base, apply_also_obc_invariants=False).output.numpy()[:, 0]
    path = os.path.join(OUT, 'reference_ensembles.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, len(out), 'arrays')


def sr_algebra():
    """tests/golden/reference_sr_algebra.npz: the algebra of the reference's ComplexValuesStochasticReconfiguration
    (optimizers/stochastic_reconfiguration/optimizer.py:33-124, complex_values_optimizer.py:48-58) on given per-sample
    derivatives O [B, P] complex, targets y_true and complex weights: centred Jacobian, right-hand side, direct solve of
    (Obar^H Obar / B + lambda I) delta = F, and the update of the (real, imag) variables.  The class cannot be instantiated
    without Keras (it reads model.layers / .weights in its constructor), so its METHODS are bound to a plain namespace that
    holds what the constructor would have stored; the bodies that run are the reference's."""
    import torch
    from oracle import tf_standin
    load_reference_machines()            # installs the stand-in and the flowket package stubs
    pkg = types.ModuleType('flowket.optimizers')
    pkg.__path__ = [REF + '/optimizers']
    sys.modules['flowket.optimizers'] = pkg
    sub = types.ModuleType('flowket.optimizers.stochastic_reconfiguration')
    sub.__path__ = [REF + '/optimizers/stochastic_reconfiguration']
    sys.modules['flowket.optimizers.stochastic_reconfiguration'] = sub
    le = types.ModuleType('flowket.optimizers.stochastic_reconfiguration.linear_equations')
    le.conjugate_gradient = None          # (the TF while-loop solver is not run here; the direct solve is)
    sys.modules[le.__name__] = le
    sr_cls = importlib.import_module('flowket.optimizers.stochastic_reconfiguration.optimizer').ComplexValuesStochasticReconfiguration
    rng = np.random.default_rng(20261023)
    B, shapes = 40, [(3, 2, 4), (4,), (1, 4, 2), (2,)]               # two complex convs: kernel + bias each
    P = sum(int(np.prod(sh)) for sh in shapes)
    O = torch.from_numpy(rng.normal(size=(B, P)) + 1j * rng.normal(size=(B, P)))
    e_loc = rng.normal(size=B) * 2 - 7 + 1j * rng.normal(size=B)
    y_true = torch.from_numpy(np.conj(e_loc - e_loc.mean()) / B)      # what VariationalMonteCarlo feeds as targets
    real = [tf_standin.Variable(rng.normal(size=sh)) for sh in shapes]
    imag = [tf_standin.Variable(rng.normal(size=sh)) for sh in shapes]
    me = types.SimpleNamespace(
        batch_size=torch.tensor(complex(B)), diag_shift=0.05, lr=0.01, use_cholesky=True, add_s_matrix_stats=False,
        use_energy_loss=False, iterative_solver=False,
        predictions_keras_model=types.SimpleNamespace(output=torch.zeros((B, 1), dtype=torch.complex128), targets=[y_true]),
        get_predictions_jacobian=lambda: O, model_real_weights=real, model_imag_weights=imag)
    for name in ('get_wave_function_jacobian_minus_mean', 'get_energy_grad', '_update_s_matrix_stats',
                 'compute_wave_function_gradient_covariance_inverse_multiplication',
                 'compute_wave_function_gradient_covariance_inverse_multiplication_directly', 'apply_complex_gradient'):
        setattr(me, name, types.MethodType(getattr(sr_cls, name), me))
    o_bar = me.get_wave_function_jacobian_minus_mean()
    energy_grad = me.get_energy_grad(None, o_bar)
    delta = me.compute_wave_function_gradient_covariance_inverse_multiplication(energy_grad, o_bar)
    updates = me.apply_complex_gradient(delta * (-1.0 + 0j))          # as in get_updates (optimizer.py:43)
    out = {'O': O.numpy(), 'y_true': y_true.numpy(), 'diag_shift': np.float64(0.05), 'lr': np.float64(0.01),
           'o_bar': o_bar.numpy(), 'energy_grad': energy_grad.numpy()[:, 0], 'delta': delta.numpy()[:, 0],
           'weights_real': np.concatenate([np.asarray(w).reshape(-1) for w in real]),
           'weights_imag': np.concatenate([np.asarray(w).reshape(-1) for w in imag]),
           'new_weights_real': np.concatenate([np.asarray(new).reshape(-1) for _, new in updates[:len(shapes)]]),
           'new_weights_imag': np.concatenate([np.asarray(new).reshape(-1) for _, new in updates[len(shapes):]])}
    path = os.path.join(OUT, 'reference_sr_algebra.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, len(out), 'arrays')


def conjugate_gradient_solver():
    """tests/golden/reference_conjugate_gradient.npz: the reference's conjugate-gradient solver
    (optimizers/stochastic_reconfiguration/linear_equations.py:34-137, the TF-contrib solver it vendors) on complex
    Hermitian positive-definite systems, at the loose tolerance the SR optimizer uses (1e-3) and a tight one.  The TF
    internals it imports are replaced by their meaning: vdot, norm, expand_dims / squeeze, and `while_loop` as a Python
    loop."""
    import torch
    from oracle import tf_standin
    tf_standin.install()

    def module(name, **attrs):
        mod = types.ModuleType(name)
        mod.__dict__.update(attrs)
        sys.modules[name] = mod
        return mod

    def while_loop(cond, body, loop_vars):
        loop_vars = list(loop_vars)
        while bool(cond(*loop_vars)):
            loop_vars = list(body(*loop_vars))
        return loop_vars
    tf = sys.modules['tensorflow']
    module('tensorflow.contrib')
    module('tensorflow.contrib.solvers')
    module('tensorflow.contrib.solvers.python')
    util = module('tensorflow.contrib.solvers.python.ops.util', dot=lambda x, y: torch.sum(torch.conj(x) * y))
    module('tensorflow.contrib.solvers.python.ops', util=util)
    module('tensorflow.python.framework')
    module('tensorflow.python.framework.constant_op', constant=lambda v, dtype=None: v)
    module('tensorflow.python.framework.dtypes', int32=tf.int32)
    module('tensorflow.python.framework.ops', name_scope=tf.name_scope)
    module('tensorflow.python.ops.array_ops', expand_dims=lambda x, axis: x.unsqueeze(axis), squeeze=lambda x: x.squeeze(),
           zeros=lambda n, dtype=None: torch.zeros(tuple(int(v) for v in n), dtype=dtype))
    module('tensorflow.python.ops.control_flow_ops', while_loop=while_loop)
    module('tensorflow.python.ops.linalg_ops', norm=lambda v: torch.linalg.vector_norm(v))
    module('tensorflow.python.ops.math_ops', logical_and=lambda a, b: bool(a) and bool(b),
           cast=lambda x, dtype: torch.as_tensor(x).to(torch.float32))
    pkg = types.ModuleType('flowket')
    pkg.__path__ = [REF]
    sys.modules['flowket'] = pkg
    for name, path in [('flowket.optimizers', '/optimizers'),
                       ('flowket.optimizers.stochastic_reconfiguration', '/optimizers/stochastic_reconfiguration')]:
        mod = types.ModuleType(name)
        mod.__path__ = [REF + path]
        sys.modules[name] = mod
    le = importlib.import_module('flowket.optimizers.stochastic_reconfiguration.linear_equations')
    Operator = __import__('collections').namedtuple('Operator', 'shape,dtype,apply')
    rng = np.random.default_rng(20261024)
    out = {}
    for name, n, cond_spread, tol, max_iter in [('loose', 60, 3.0, 1e-3, 200), ('tight', 60, 3.0, 1e-6, 500),
                                                ('capped', 80, 5.0, 1e-12, 7)]:
        A = rng.normal(size=(n, 2 * n)) + 1j * rng.normal(size=(n, 2 * n))
        A = A * np.logspace(0, -cond_spread, n)[:, None]
        S = torch.from_numpy(A @ A.conj().T / (2 * n) + 0.05 * np.eye(n))
        rhs = torch.from_numpy(rng.normal(size=n) + 1j * rng.normal(size=n))

        op = Operator(shape=[n, n], dtype=torch.complex128, apply=lambda v, S=S: S @ v)
        x0 = torch.zeros(n, dtype=torch.complex128)    # explicit start vector: the zeros() branch needs dtype.base_dtype
        res = le.conjugate_gradient(op, rhs, x=x0, tol=tol, max_iter=max_iter)
        out[name + '/S'] = S.numpy()
        out[name + '/rhs'] = rhs.numpy()
        out[name + '/tol'] = np.float64(tol)
        out[name + '/max_iter'] = np.int64(max_iter)
        out[name + '/x'] = res.x.numpy()
        out[name + '/iterations'] = np.int64(int(res.i))
        out[name + '/residual_norm'] = np.float64(float(torch.linalg.vector_norm(res.r)))
    path = os.path.join(OUT, 'reference_conjugate_gradient.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, len(out), 'arrays', {k: int(out[k]) for k in out if k.endswith('iterations')})




PRETRAINED = {'2': 'ising_2.h5', '2_5': 'ising_2_5.h5', '3': 'ising_3.h5', '3_5': 'ising_3_5.h5', '4': 'ising_4.h5'}


def export_pretrained_weights():
    """tests/golden/ising_12x12_gamma<G>_keras_weights.npz: the reference's pretrained Keras weights
    experiments/weights/ising_<G>.h5 (Ising 12x12 OBC, Gamma = 2, 2.5, 3, 3.5, 4; depth 10 / 32 channels, weight-normalised),
    read with the repository's pure-Python HDF5 reader and stored in Keras layer-creation order.  Published evaluation
    (experiments/README.md:38-47, symmetrised psi, 2^15 samples): E = -346.9817926, -395.6618438, -457.0420317,
    -524.5172088, -593.5389339."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from flowket_b200 import Input
    from flowket_b200.machines import ConvNetAutoregressive2D
    from flowket_b200.utils.keras_h5 import read_keras_weights
    m = ConvNetAutoregressive2D(Input(shape=(12, 12)), depth=10, num_of_channels=32)
    for tag, fname in PRETRAINED.items():
        w = read_keras_weights('/root/reference/experiments/weights/' + fname, m.weight_specs())
        path = os.path.join(OUT, 'ising_12x12_gamma%s_keras_weights.npz' % tag)
        np.savez_compressed(path, **{'w%04d' % i: a for i, a in enumerate(w)})
        print('wrote', path, sum(a.size for a in w), 'parameters')




def fast_sampler():
    """tests/golden/reference_fast_sampler.npz: the reference's own FastAutoregressiveSampler
    (deepar/samplers/fast_autoregressive.py:13-76) with its DependencyGraph (deepar/graph_analysis/dependency_graph.py:35-94)
    and every layer topology it needs (convolution incl. weight-norm wrappers and complex convolutions, zero padding, down /
    right shifts, one-to-one layers, concatenation, the +-1 sampling topology) around the reference's own machines.
    The stand-in records which layer feeds which while the machine is built (the information Keras keeps in
    `inbound_nodes`), evaluates the per-site operations the sampler schedules eagerly, and defines `tf.multinomial` -- the
    one unseeded primitive -- by the explicit-uniform rule with injected uniforms, handed out in the order in which the
    sampler visits the sites.  Result: the spins the reference's cached incremental sampler draws for given weights and
    uniforms, to be compared with the N-forward rule (oracle) and the CUDA samplers."""
    import networkx  # noqa: F401  (the reference's dependency graph needs it)
    if not hasattr(np, 'product'):
        np.product = np.prod                 # alias the reference still uses (removed in numpy 2)
    import torch
    from oracle import nets, tf_standin
    tf_standin.install()

    def package(name, path):
        mod = types.ModuleType(name)
        mod.__path__ = [path]
        sys.modules[name] = mod
        return mod
    imp = importlib.import_module
    package('flowket', REF)
    package('flowket.machines', REF + '/machines')
    layers = package('flowket.layers', REF + '/layers')
    package('flowket.layers.complex', REF + '/layers/complex')
    package('flowket.deepar', REF + '/deepar')
    package('flowket.deepar.utils', REF + '/deepar/utils').Singleton = imp('flowket.deepar.utils.singleton').Singleton
    deepar_layers = package('flowket.deepar.layers', REF + '/deepar/layers')
    package('flowket.deepar.samplers', REF + '/deepar/samplers')
    ga = package('flowket.deepar.graph_analysis', REF + '/deepar/graph_analysis')
    for sub, names in [('autoregressive', ['NormalizeInLogSpace', 'CombineAutoregressiveConditionals']),
                       ('casting', ['CastingLayer', 'ToFloat32', 'ToFloat64']),
                       ('lambda_with_one_to_one_topology', ['LambdaWithOneToOneTopology']),
                       ('masking', ['DownShiftLayer', 'RightShiftLayer']),
                       ('one_hot', ['ToOneHot', 'PlusMinusOneToOneHot']),
                       ('padding', ['ExpandInputDim', 'PeriodicPadding']),
                       ('wrappers', ['WeightNormalization']),
                       ('layer_normalization', ['LayerNormalization'])]:
        mod = imp('flowket.deepar.layers.' + sub)
        for n in names:
            setattr(deepar_layers, n, getattr(mod, n))
    for sub, names in [('topology_manager', ['TopologyManager']), ('layer_topology', ['LayerTopology']),
                       ('data_structures', ['Dependency', 'GraphNode']),
                       ('one_to_one_topology', ['OneToOneTopology', 'OneToOneTopologyWithIdentity']),
                       ('convolutional_topology', ['ConvolutionalTopology']), ('padding_topology', ['PaddingTopology']),
                       ('masking_topology', ['DownShiftTopology', 'RightShiftTopology']),
                       ('concatenate_topology', ['ConcatenateTopology']),
                       ('sampling_topology', ['PlusMinusOneSamplingTopology', 'CategorialSamplingTopology']),
                       ('dependency_graph', ['DependencyGraph'])]:
        mod = imp('flowket.deepar.graph_analysis.' + sub)
        for n in names:
            setattr(ga, n, getattr(mod, n))
    module = types.ModuleType('flowket.layers.complex.initializers')
    module.ConjugateDecorator = type('ConjugateDecorator', (), {
        '__init__': lambda self, initializer: None, 'get_real_part_initializer': lambda self: None,
        'get_imag_part_initializer': lambda self: None})
    module.get = lambda identifier: identifier
    sys.modules[module.__name__] = module
    casting = imp('flowket.layers.complex.casting')
    conv = imp('flowket.layers.complex.conv')
    spins = imp('flowket.layers.spins_invariants')
    for mod, names in [(casting, ['VectorToComplexNumber', 'ToComplex64', 'ToComplex128']), (conv, ['ComplexConv1D']),
                       (spins, ['EqualUpDownSpins'])]:
        for n in names:
            setattr(layers, n, getattr(mod, n))
    # +-1 spins: flowket/samplers/fast_autoregressive/__init__.py:8
    ga.TopologyManager().register_layer_topology(tf_standin.InputLayer, ga.PlusMinusOneSamplingTopology)
    fast_cls = imp('flowket.deepar.samplers.fast_autoregressive').FastAutoregressiveSampler
    classes = {'conv2d': imp('flowket.machines.conv_net_autoregressive_2D').ConvNetAutoregressive2D,
               'conv1d': imp('flowket.machines.simple_conv_net_autoregressive_1D').SimpleConvNetAutoregressive1D,
               'cconv1d': imp('flowket.machines.complex_values_simple_conv_net_autoregressive_1D').ComplexValuesSimpleConvNetAutoregressive1D}
    rng = np.random.default_rng(20261025)
    out = {}
    for name, kind, shape, kw, spec in [
        ('conv2d_4x3', 'conv2d', (4, 3), dict(depth=2, num_of_channels=8), nets.Conv2DSpec(4, 3, 2, 8)),
        ('conv2d_3x3_d3', 'conv2d', (3, 3), dict(depth=3, num_of_channels=4), nets.Conv2DSpec(3, 3, 3, 4)),
        ('conv1d_10', 'conv1d', (10,), dict(depth=4, num_of_channels=8, max_dilation_rate=2, add_skip_connections=True),
         nets.Conv1DSpec(10, 4, 8, max_dilation_rate=2, add_skip_connections=True)),
        ('cconv1d_8', 'cconv1d', (8,), dict(depth=3, num_of_channels=4, max_dilation_rate=2),
         nets.ComplexConv1DSpec(8, 3, 4, max_dilation_rate=2)),
    ]:
        B = 24
        params = [p + 0.3 * torch.randn(p.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(7 + i)) * (p.dim() == 1)
                  for i, p in enumerate(nets.init_params(spec, seed=5, dtype=torch.float64))]
        uniforms = rng.random((B,) + shape)

        def build_sampler(uniform_queue):
            tf_standin.inject_weights([p.numpy() for p in params])
            tf_standin.BATCH[0] = B
            del tf_standin.UNIFORMS[:]
            tf_standin.UNIFORMS.extend(uniform_queue)
            tf_standin.MULTINOMIAL_CALLS[0] = 0
            x = torch.zeros((B,) + shape, dtype=torch.float64)          # the machine is built on a placeholder batch
            tf_standin.InputLayer(x, dtype='int8')
            tf_standin.RECORDING[0] = True
            machine = classes[kind](x, **kw)
            model = tf_standin.GraphModel(x, machine.conditional_log_probs)
            tf_standin.RECORDING[0] = False
            return fast_cls(model, B)

        probe = build_sampler([])                                        # pass 1: in which order are the sites visited?
        order = [node.spatial_location for node in probe.sampling_order if node.layer is probe.input_layer]
        assert sorted(order) == sorted(np.ndindex(*shape)) and tf_standin.MULTINOMIAL_CALLS[0] == len(order)
        sampler = build_sampler([uniforms[(slice(None),) + site] for site in order])          # pass 2: the real draw
        sigma = np.asarray(next(sampler)).astype(np.int8)
        assert sigma.shape == (B,) + shape and not tf_standin.UNIFORMS
        out[name + '/params'] = nets.flatten_params(params).numpy()
        out[name + '/uniforms'] = uniforms
        out[name + '/sigma'] = sigma
        out[name + '/site_order'] = np.array(order, dtype=np.int64)
        out[name + '/graph_nodes'] = np.int64(probe.dependencies_graph.graph.number_of_nodes())
    path = os.path.join(OUT, 'reference_fast_sampler.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, len(out), 'arrays', {k: int(out[k]) for k in out if k.endswith('graph_nodes')})


def complex_sr_pipeline():
    """tests/golden/reference_complex_sr_pipeline.npz: the stochastic-reconfiguration update of the complex 1-D machine
    (BASELINE configs[3]) computed by the reference's own code end to end: ComplexValuesSimpleConvNetAutoregressive1D forward,
    `Machine.predictions_jacobian` (machines/abstract_machine.py:24-28; the Jacobian primitive is autograd over the
    stand-in's torch graph), the complex assembly of `ComplexValuesOptimizer.get_predictions_jacobian`
    (complex_values_optimizer.py:8-9,72-76), then the SR methods as in `sr_algebra`."""
    import torch
    from oracle import nets, tf_standin
    cconv_cls = load_reference_machines()[2]
    for name, path in [('flowket.optimizers', '/optimizers'),
                       ('flowket.optimizers.stochastic_reconfiguration', '/optimizers/stochastic_reconfiguration')]:
        mod = types.ModuleType(name)
        mod.__path__ = [REF + path]
        sys.modules[name] = mod
    le = types.ModuleType('flowket.optimizers.stochastic_reconfiguration.linear_equations')
    le.conjugate_gradient = None
    sys.modules[le.__name__] = le
    sr_cls = importlib.import_module('flowket.optimizers.stochastic_reconfiguration.optimizer').ComplexValuesStochasticReconfiguration
    rng = np.random.default_rng(20261026)
    spec = nets.ComplexConv1DSpec(9, 3, 4, max_dilation_rate=2)
    kw = dict(depth=3, num_of_channels=4, max_dilation_rate=2)
    B = 30
    params = [p + 0.3 * torch.randn(p.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(11 + i)) * (p.dim() == 1)
              for i, p in enumerate(nets.init_params(spec, seed=3, dtype=torch.float64))]
    sigma = rng.choice([-1, 1], size=(B, 9)).astype(np.int8)
    leaves = tf_standin.inject_weights([p.numpy() for p in params], requires_grad=True)
    machine = cconv_cls(torch.from_numpy(sigma.astype(np.float64)), **kw)
    e_loc = rng.normal(size=B) * 2 - 5 + 1j * rng.normal(size=B)
    y_true = torch.from_numpy(np.conj(e_loc - e_loc.mean()) / B)
    real = [tf_standin.Variable(w.detach()) for w in leaves[0::2]]
    imag = [tf_standin.Variable(w.detach()) for w in leaves[1::2]]
    me = types.SimpleNamespace(
        batch_size=torch.tensor(complex(B)), diag_shift=0.05, lr=0.01, use_cholesky=True, add_s_matrix_stats=False,
        use_energy_loss=False, iterative_solver=False, predictions_jacobian=machine.predictions_jacobian,
        predictions_keras_model=types.SimpleNamespace(output=machine.predictions, targets=[y_true], weights=leaves),
        model_real_weights=real, model_imag_weights=imag)
    for name in ('get_predictions_jacobian', 'get_wave_function_jacobian_minus_mean', 'get_energy_grad', '_update_s_matrix_stats',
                 'compute_wave_function_gradient_covariance_inverse_multiplication',
                 'compute_wave_function_gradient_covariance_inverse_multiplication_directly', 'apply_complex_gradient'):
        setattr(me, name, types.MethodType(getattr(sr_cls, name), me))
    jac = me.get_predictions_jacobian().detach()
    o_bar = me.get_wave_function_jacobian_minus_mean().detach()
    energy_grad = me.get_energy_grad(None, o_bar)
    delta = me.compute_wave_function_gradient_covariance_inverse_multiplication(energy_grad, o_bar)
    updates = me.apply_complex_gradient(delta * (-1.0 + 0j))
    new_flat = []
    for (_, new_r), (_, new_i) in zip(updates[:len(real)], updates[len(real):]):
        new_flat += [np.asarray(new_r.detach()).reshape(-1), np.asarray(new_i.detach()).reshape(-1)]
    out = {'params': nets.flatten_params(params).numpy(), 'sigma': sigma, 'y_true': y_true.numpy(), 'local_energy': e_loc,
           'jacobian': jac.numpy(), 'delta': delta.detach().numpy()[:, 0], 'diag_shift': np.float64(0.05), 'lr': np.float64(0.01),
           'new_params': np.concatenate(new_flat)}
    path = os.path.join(OUT, 'reference_complex_sr_pipeline.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, len(out), 'arrays; complex parameters:', jac.shape[1])


COMMANDS = {
    'numpy_half': main,                          # operators / local energy / bit conventions / ED anchors
    'edge': edge_cases,                          # find_conn over all states of degenerate lattices
    'mcmc': mcmc_diagnostics,                    # calc_r_hat_value / sum_correlations
    'j1j2': j1j2_graph,                          # coloured edge list + bond operators handed to netket
    'exact': exact_variational,                  # ExactVariational / ExactObservable
    'vmc': variational_monte_carlo,              # VariationalMonteCarlo + MiniBatchGenerator
    'callbacks': callbacks_logs,                 # stats callbacks, evaluate, exact_evaluate, BadEigenStateStopping
    'sampler': autoregressive_sampler,           # AutoregressiveSampler.__next__ around the oracle network
    'complex_ops': complex_ops,                  # lncosh / complex_log / ensemble ops
    'machines': machines,                        # the three machine classes + gradients (oracle/tf_standin.py)
    'fast_sampler': fast_sampler,                # FastAutoregressiveSampler + DependencyGraph + topologies
    'ensembles': ensembles,                      # symmetrisation ensembles around the 2-D machine
    'sr': sr_algebra,
    'sr_pipeline': complex_sr_pipeline,          # complex machine -> Jacobian -> SR update, end to end                            # ComplexValuesStochasticReconfiguration methods
    'cg': conjugate_gradient_solver,             # the vendored conjugate-gradient solver
    'weights': export_pretrained_weights,        # experiments/weights/ising_*.h5 -> npz
}


if __name__ == '__main__':
    # every generator installs its own module stand-ins: run each in a fresh interpreter
    #   python -m oracle.make_golden            -> numpy_half
    #   python -m oracle.make_golden machines   -> one fixture
    #   python -m oracle.make_golden all        -> everything, one subprocess per fixture
    names = sys.argv[1:] or ['numpy_half']
    if names == ['all']:
        import subprocess
        for name in COMMANDS:
            subprocess.check_call([sys.executable, '-m', 'oracle.make_golden', name])
    else:
        for name in names:
            COMMANDS[name]()
