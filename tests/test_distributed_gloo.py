"""CPU tier: the N > 1 path (energy-statistics allreduce, gradient allreduce, sample sharding) with
world_size-2 gloo processes."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from flowket_b200.optimization import DistributedVariationalMonteCarlo
    from flowket_b200.optimizers import allreduce_sum_
    rng = np.random.default_rng(123)
    full = rng.normal(size=64) * 3 - 20 + 1j * rng.normal(size=64) * 0.1     # the global batch of local energies
    shard = full[rank * 32:(rank + 1) * 32]
    mean, var, count = DistributedVariationalMonteCarlo.reduce_stats(shard)
    g = torch.full((5,), float(rank + 1))
    allreduce_sum_(g)
    out[rank] = (mean, var, count, g.tolist(), complex(full.mean()), float(np.var(full.real)))
    dist.destroy_process_group()


def test_world_size_2_energy_statistics_and_gradient_allreduce():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29651, out), nprocs=world, join=True)
    for rank in range(world):
        mean, var, count, g, want_mean, want_var = out[rank]
        assert count == 64
        assert mean == pytest.approx(want_mean, rel=1e-12)     # a true global mean: sum / count, not a mean of means
        assert var == pytest.approx(want_var, rel=1e-10)
        assert g == [3.0] * 5


def test_philox_shard_offsets_partition_the_global_batch():
    """rank r draws global sample indices [r*B, (r+1)*B): the host bookkeeping of the sampler"""
    from flowket_b200 import Input, Model
    from flowket_b200.machines import ConvNetAutoregressive2D
    from flowket_b200.samplers import FastAutoregressiveSampler
    inp = Input(shape=(4, 4))
    cond = Model(inp, ConvNetAutoregressive2D(inp, depth=2, num_of_channels=8).conditional_log_probs)
    offs = [FastAutoregressiveSampler(cond, 32, seed=7, sample_offset=r * 32).sample_offset for r in range(4)]
    assert offs == [0, 32, 64, 96]
    s = FastAutoregressiveSampler(cond, 100, mini_batch_size=32)
    assert s._effective_batch() == 96          # batch % mini_batch samples are dropped (fast_autoregressive.py:31-33)
    assert s.copy_with_new_batch_size(16).batch_size == 16


def _shard_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from flowket_b200 import Input, Model
    from flowket_b200.machines import ConvNetAutoregressive2D
    from flowket_b200.operators import Heisenberg
    from flowket_b200.optimization import DistributedVariationalMonteCarlo
    from flowket_b200.samplers import FastAutoregressiveSampler
    inp = Input(shape=(4, 4))
    m = ConvNetAutoregressive2D(inp, depth=2, num_of_channels=8)
    cond = Model(inp, m.conditional_log_probs)
    # a straight port of a reference Horovod script: nobody sets sample_offset
    sampler = FastAutoregressiveSampler(cond, 32, seed=7)
    vmc = DistributedVariationalMonteCarlo(Model(inp, m.predictions), Heisenberg(hilbert_state_shape=[4, 4], pbc=False), sampler)
    first = sampler.sample_offset
    bigger = sampler.copy_with_new_batch_size(128)
    vmc.set_sampler(bigger)
    out[rank] = (first, bigger.sample_offset, vmc.global_batch_size)
    dist.destroy_process_group()


def test_distributed_vmc_gives_every_rank_its_own_slice_of_the_philox_stream():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_shard_worker, args=(world, 29655, out), nprocs=world, join=True)
    assert out[0] == (0, 0, 256) and out[1] == (32, 128, 256)


def _sr_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from flowket_b200.optimizers import sr_delta, real_sr_delta
    rng = np.random.default_rng(7)
    B, P = 48, 11
    O = torch.from_numpy(rng.normal(size=(B, P)) + 1j * rng.normal(size=(B, P)))
    e = torch.from_numpy(rng.normal(size=B) * 2 - 5 + 1j * rng.normal(size=B))
    lo, hi = (0, 20) if rank == 0 else (20, 48)          # ragged shards
    res = {}
    for iterative in (False, True):
        # complex-parameter SR: y_true = conj(E_loc - E_global) / B_local, as DistributedVariationalMonteCarlo hands it over
        y_full = torch.conj(e - e.mean()) / B
        want = sr_delta(O, y_full, 0.05, iterative, 1e-12, 500)
        y_loc = torch.conj(e[lo:hi] - e.mean()) / (hi - lo)
        got = sr_delta(O[lo:hi], y_loc, 0.05, iterative, 1e-12, 500, distributed=True)
        res['complex', iterative] = float((got - want).abs().max() / want.abs().max())
        want = real_sr_delta(O.real, O.imag, e, 0.05, iterative, 1e-12, 500)
        got = real_sr_delta(O.real[lo:hi], O.imag[lo:hi], e[lo:hi], 0.05, iterative, 1e-12, 500, distributed=True)
        res['real', iterative] = float((got - want).abs().max() / want.abs().max())
    out[rank] = res
    dist.destroy_process_group()


def test_world_size_2_stochastic_reconfiguration_matches_single_process():
    """SR system sharded over two ranks (means, right-hand side, P x P matrix or CG products allreduced) == the same system
    assembled from the whole batch by one process"""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_sr_worker, args=(world, 29653, out), nprocs=world, join=True)
    for rank in range(world):
        for key, err in out[rank].items():
            assert err < 1e-8, (rank, key, err)


def _sample_space_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from flowket_b200.optimizers import real_sr_delta, sample_space_sr_delta
    rng = np.random.default_rng(11)
    B, P = 24, 203                                           # P >> 2B and not a multiple of the slice alignment
    R = torch.from_numpy(rng.normal(size=(B, P)))
    I = torch.from_numpy(rng.normal(size=(B, P)))
    e = torch.from_numpy(rng.normal(size=B) * 2 - 5 + 1j * rng.normal(size=B))
    want = real_sr_delta(R, I, e, 0.05)                      # P x P system, whole batch, one process
    bounds = [0, 7, 24] if world == 2 else [0, 5, 13, 24]    # ragged shards
    lo, hi = bounds[rank], bounds[rank + 1]
    Rb, Ib, eb = R - R.mean(0), I - I.mean(0), e - e.mean()
    X = torch.cat([Rb[lo:hi], Ib[lo:hi]])
    ep = torch.cat([eb.real[lo:hi], eb.imag[lo:hi]])
    res = {}
    got = sample_space_sr_delta(X, ep, 0.05, distributed=True)
    res['fp64'] = float((got - want).abs().max() / want.abs().max())
    got = sample_space_sr_delta(X.float(), ep.float(), 0.05, distributed=True, low_precision=torch.bfloat16)
    res['bf16'] = float((got.double() - want).abs().max() / want.abs().max())
    # identical on every rank
    ref = got.clone()
    dist.broadcast(ref, 0)
    res['replicated'] = float((got - ref).abs().max())
    # the replicated system factorised once, block columns dealt to the ranks, instead of once per rank
    got = sample_space_sr_delta(X, ep, 0.05, distributed=True, shared_cholesky=True)
    res['shared_cholesky'] = float((got - want).abs().max() / want.abs().max())
    from flowket_b200.optimizers import distributed_cholesky_solve
    A = torch.from_numpy(rng.normal(size=(45, 60)))
    T = A @ A.T / 60 + 0.05 * torch.eye(45, dtype=torch.float64)
    rhs = torch.from_numpy(rng.normal(size=45))
    T_before = T.clone()
    w = distributed_cholesky_solve(T, rhs, block=7)            # 7 block columns (the last one ragged) over 2 or 3 ranks
    assert torch.equal(T, T_before)                            # the replicated system is read, never written
    res['cholesky_solve'] = float((w - torch.linalg.solve(T, rhs)).abs().max())
    w0 = w.clone()
    dist.broadcast(w0, 0)
    res['cholesky_replicated'] = float((w - w0).abs().max())
    # one process, same function
    single = sample_space_sr_delta(torch.cat([Rb, Ib]), torch.cat([eb.real, eb.imag]), 0.05)
    res['single'] = float((single - want).abs().max() / want.abs().max())
    out[rank] = res
    dist.destroy_process_group()


@pytest.mark.parametrize('world,port', [(2, 29655), (3, 29657)])
def test_sample_space_stochastic_reconfiguration_sharded_matches_single_process(world, port):
    """all-to-all re-shard of X to parameter-major, partial Grams, allreduce, replicated solve == the
    P x P solve of the whole batch (push-through identity), for ragged shards and any world size"""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_sample_space_worker, args=(world, port, out), nprocs=world, join=True)
    for rank in range(world):
        res = out[rank]
        assert res['fp64'] < 1e-9 and res['single'] < 1e-9, (rank, res)
        assert res['bf16'] < 2e-2, (rank, res)
        assert res['replicated'] == 0.0, (rank, res)
        assert res['shared_cholesky'] < 1e-9 and res['cholesky_solve'] < 1e-10 and res['cholesky_replicated'] == 0.0, (rank, res)


class _TableModel(object):
    """log-amplitude table behind Model.predict / input_shape (host stand-in for the CUDA forward)"""

    def __init__(self, shape, seed=0):
        from flowket_b200.exact.utils import vector_to_machine
        n = int(np.prod(shape))
        rng = np.random.default_rng(seed)
        self.vector = rng.normal(scale=0.5, size=2 ** n) + 1j * rng.uniform(-3, 3, size=2 ** n)
        self._f = vector_to_machine(self.vector)
        self.input_shape = (None,) + tuple(shape)

    def predict(self, x, batch_size=None):
        return self._f(np.asarray(x))


def _exact_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from flowket_b200.optimization import ExactVariational, DistributedExactVariational
    from oracle import operators as oops
    shape = (3, 2)
    model = _TableModel(shape, seed=4)
    op = oops.OracleOperator('heisenberg', shape, pbc=False)
    whole = ExactVariational(model, op, 16)
    whole.machine_updated()
    mine = DistributedExactVariational(model, op, 16)
    assert (mine.slice_hi - mine.slice_lo) * world == 64 and mine.num_of_batch_until_full_cycle == 64 // world // 16
    gen = mine.to_generator()
    xs, ys = zip(*[next(gen) for _ in range(mine.num_of_batch_until_full_cycle)])
    lo, hi = mine.slice_lo, mine.slice_hi
    res = {
        'energy': abs(mine.energy_observable.current_energy - whole.energy_observable.current_energy),
        'variance': abs(mine.energy_observable.current_local_energy_variance - whole.energy_observable.current_local_energy_variance),
        'probs': float(np.abs(mine.probs - whole.probs).max()),
        'states': bool(np.array_equal(np.concatenate(xs), whole.states[lo:hi])),
        'coefficients': float(np.abs(np.concatenate(ys) - whole.energy_grad_coefficients[lo:hi]).max()),
        'tables': mine.energy_observable.states_idx_local_connections.shape[1],
    }
    out[rank] = res
    dist.destroy_process_group()


def test_exact_variational_sharded_over_two_ranks_matches_the_whole_enumeration():
    """the 2^N states shard over the ranks; energy, variance, probabilities and the gradient coefficients of each
    slice equal those of the single-process enumeration, and the connection tables shrink by the number of ranks"""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_exact_worker, args=(world, 29659, out), nprocs=world, join=True)
    for rank in range(world):
        res = out[rank]
        assert res['energy'] < 1e-12 and res['variance'] < 1e-10 and res['probs'] < 1e-15, (rank, res)
        assert res['states'] and res['coefficients'] < 1e-13 and res['tables'] == 32, (rank, res)


def _deal_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from flowket_b200.optimizers.sample_space_sr import deal_local_energies, split_solve_shares
    from oracle import operators as oops, local_energy as oeloc
    shape = (3, 3)
    op = oops.OracleOperator('heisenberg', shape, pbc=False)
    rng = np.random.default_rng(5)
    Bl = 12
    full = rng.choice([-1, 1], size=(Bl * world, 9)).astype(np.int8)      # the global batch, rank-major
    w = rng.normal(size=9) + 1j * rng.normal(size=9) * 0.3

    def log_psi(x):       # any deterministic wave function of the configuration
        x = np.asarray(x, np.float64).reshape(len(x), -1)
        return x @ w + 0.1 * (x[:, :-1] * x[:, 1:]).sum(axis=1)

    def fn(sig):          # the local energies of the samples it is handed (oracle: local_energy.py)
        x = sig.numpy().reshape((-1,) + shape)
        return torch.as_tensor(np.asarray(oeloc.local_values(op, lambda c: log_psi(c)[:, None], x), np.complex128))

    calls = []
    mine = torch.as_tensor(full[rank * Bl:(rank + 1) * Bl])
    want = fn(torch.as_tensor(full))
    res = {}
    for rho in (0.0, 0.2, 0.6, 5.0):
        counts = split_solve_shares(world, Bl * world, rho, solver_rank=0, kappa=1.2)
        got = deal_local_energies(mine, fn, counts, rank, world, after_gather=lambda: calls.append(rank))
        res[rho] = (counts, float((got - want).abs().max()))
    res['calls'] = len(calls)
    out[rank] = res
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_split_solve_deals_the_local_energies_over_the_ranks(world):
    """sample_space_sr.deal_local_energies (the N > 1 half of the split solve): whatever the shares -- equal, small or empty
    solver share, remainders -- every rank ends up with the local energies of the whole global batch in global order, equal
    to evaluating them in one process"""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_deal_worker, args=(world, 29663 + world, out), nprocs=world, join=True)
    for rank in range(world):
        res = out[rank]
        assert res['calls'] == 4
        for rho in (0.0, 0.2, 0.6, 5.0):
            counts, err = res[rho]
            assert sum(counts) == 12 * world and err < 1e-12, (rank, rho, counts, err)
        assert res[5.0][0][0] == 0 and res[0.0][0][0] > 0
