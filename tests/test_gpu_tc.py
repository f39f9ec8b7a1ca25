"""Tensor-core (tcgen05, bf16 operands / fp32 accumulation) engine vs the fp32 engine and the fp64 oracle.
Stated tolerance of the bf16 path (north_star allows a separately stated tolerance): see TOL below."""
import numpy as np
import pytest
import torch

from oracle import nets
from tests.helpers import make_pair, random_sigma

pytestmark = pytest.mark.gpu

# |log psi_tc - log psi_fp64| <= TOL_ABS + TOL_REL * |log psi|   (bf16 rounding of activations and weights)
TOL_ABS, TOL_REL = 0.05, 2e-3

CASES = [((4, 4), 3), ((6, 6), 4), ((10, 10), 5), ((10, 10), 20), ((12, 12), 3), ((16, 16), 2), ((5, 7), 2)]


@pytest.mark.parametrize('shape,depth', CASES)
def test_tc_log_psi(shape, depth):
    from flowket_b200 import FK_ENGINE_TC
    model, _, spec, params = make_pair('conv2d', shape, depth, 32, seed=7)
    sigma = random_sigma(301, shape, seed=2)   # odd count: exercises the idle second pipeline
    ref32 = model.predict(sigma)[:, 0]
    model.engine = FK_ENGINE_TC
    got = model.predict(sigma)[:, 0]
    want = nets.log_psi_numpy(spec, params, sigma[:64])[:, 0]
    err32 = np.abs(got - ref32)
    print('shape', shape, 'depth', depth, 'max |tc - fp32| =', err32.max(), 'mean', err32.mean(),
          '|log psi| ~', np.abs(ref32).mean())
    assert np.all(np.abs(got[:64] - want) <= TOL_ABS + TOL_REL * np.abs(want))
    assert np.all(err32 <= TOL_ABS + TOL_REL * np.abs(ref32))


def test_tc_local_energy():
    from flowket_b200 import FK_ENGINE_TC
    from flowket_b200.observables.monte_carlo import Observable
    from flowket_b200.operators import Heisenberg
    shape = (10, 10)
    model, _, spec, params = make_pair('conv2d', shape, 6, 32, seed=9)
    sigma = random_sigma(64, shape, seed=3)
    obs = Observable(Heisenberg(hilbert_state_shape=list(shape), pbc=False))
    e32 = obs.local_values(model, sigma)
    model.engine = FK_ENGINE_TC
    etc = obs.local_values(model, sigma)
    rel = np.abs(etc - e32) / np.abs(e32)
    print('E_loc: max rel |tc - fp32| =', rel.max(), 'mean', rel.mean())
    assert rel.max() < 5e-2
    assert abs(etc.mean() - e32.mean()) / abs(e32.mean()) < 5e-3
