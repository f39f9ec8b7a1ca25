"""CPU tier: the dependency analysis behind the planned row-trimmed local-energy forward.
log psi of every connected configuration, recomputed from its first changed row downwards with two halo rows taken from
the sample's own activations, equals the full forward (oracle/nets.py) to fp64 round-off."""
import numpy as np
import pytest
import torch

from oracle import nets, operators as oops, prefix_reuse as pr


@pytest.mark.parametrize('shape,depth,channels,wn,opkind,opkw', [
    ((6, 5), 3, 8, True, 'heisenberg', dict(pbc=False)),
    ((4, 4), 4, 6, False, 'heisenberg', dict(pbc=True)),       # periodic bonds connect the last row to the first: r0 = 0
    ((5, 4), 2, 8, True, 'ising', dict(pbc=False, h=1.5)),
    ((4, 4), 3, 4, False, 'j1j2', dict(pbc=False, j2=0.5)),
])
def test_row_trimmed_forward_equals_full_forward(shape, depth, channels, wn, opkind, opkw):
    spec = nets.Conv2DSpec(shape[0], shape[1], depth, channels, weights_normalization=wn)
    params = nets.init_params(spec, seed=3, dtype=torch.float64, bias_scale=0.3)
    rng = np.random.default_rng(5)
    sigma = rng.choice([-1, 1], size=(3,) + shape)
    op = oops.OracleOperator(opkind, shape, **opkw)
    conn, mel, use = op.find_conn(sigma.astype(np.float64))
    cond_base, cache = pr.forward_with_cache(spec, params, sigma)
    # the cached forward is the ordinary forward
    want_base = nets.log_psi(spec, params, sigma)
    got_base = pr._select(cond_base, sigma).sum(dim=(1, 2))
    assert (got_base - want_base).abs().max().item() < 1e-12
    checked = 0
    for b in range(sigma.shape[0]):
        cfgs = conn[np.asarray(use[:, b], bool), b].astype(np.int64)
        want = nets.log_psi(spec, params, cfgs)
        r0 = pr.first_changed_row(np.broadcast_to(sigma[b], cfgs.shape), cfgs)
        cache_b = [{k: v[b:b + 1] for k, v in blk.items()} for blk in cache]
        for row in np.unique(r0):
            sel = np.nonzero(r0 == row)[0]
            if row == shape[0]:                     # the sample itself (connection 0)
                got = got_base[b].expand(len(sel))
            else:
                n = len(sel)
                cache_rep = [{k: v.expand((n,) + v.shape[1:]) for k, v in blk.items()} for blk in cache_b]
                got = pr.log_psi_from_row(spec, params, cfgs[sel], int(row), cache_rep, cond_base[b:b + 1].expand(
                    (n,) + cond_base.shape[1:]), np.broadcast_to(sigma[b], (n,) + shape))
            assert (got - want[sel]).abs().max().item() < 1e-11, (b, row)
            checked += len(sel)
    assert checked == int(np.asarray(use, bool).sum())


def test_row_work_saved_on_the_headline_lattice():
    """10x10 OBC Heisenberg: share of lattice rows a row-trimmed forward has to recompute, averaged over the used
    connections of random samples (the number quoted in DESIGN.md section 7)."""
    rng = np.random.default_rng(0)
    shape = (10, 10)
    sigma = rng.choice([-1, 1], size=(64,) + shape)
    conn, mel, use = oops.OracleOperator('heisenberg', shape, pbc=False).find_conn(sigma.astype(np.float64))
    use = np.asarray(use, bool)
    rows = []
    for b in range(sigma.shape[0]):
        cfgs = conn[use[:, b], b]
        r0 = pr.first_changed_row(np.broadcast_to(sigma[b], cfgs.shape), cfgs)
        rows.append(shape[0] - r0)                 # rows to recompute (0 for the sample itself)
    rows = np.concatenate(rows)
    share = rows.mean() / shape[0]
    fits_m64 = np.mean(rows * (shape[1] + 2) <= 64)          # configurations whose recomputed rows fit one M = 64 tile
    print('rows recomputed: %.1f %% of the lattice on average; %.1f %% of the configurations fit an M = 64 tile' % (
        100 * share, 100 * fits_m64))
    assert 0.50 < share < 0.60
    assert 0.40 < fits_m64 < 0.60
