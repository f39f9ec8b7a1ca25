"""J1J2: constructor of flowket/operators/j1j2.py:58-69.  The reference builds a netket GraphOperator; netket is
not a dependency here: the edge list of j1j2.py:14-35 is restated and each bond contributes
sigma_a.sigma_b * J_c (diagonal +-J_c, off-diagonal 2 J_c for anti-parallel pairs, no sign rotation).  The
connection layout is the compacted one of NetketOperatorWrapper (netket_operator.py:46-66)."""
import numpy as np

from .operator import Operator
from .. import _lib


def j1j2_edges(hilbert_state_shape, pbc=False):
    L1, L2 = hilbert_state_shape
    e = []
    for h in range(L1):
        for w in range(L2 - 1):
            e.append((w + L2 * h, w + 1 + L2 * h, 1))
            if h < L1 - 1:
                e.append((w + L2 * h, w + L2 * (h + 1), 1))
                e.append((w + L2 * h, w + 1 + L2 * (h + 1), 2))
            elif pbc:
                e.append((w + L2 * h, w, 1))
                e.append((w + L2 * h, w + 1, 2))
            if h > 0:
                e.append((w + L2 * h, w + 1 + L2 * (h - 1), 2))
            elif pbc:
                e.append((w + L2 * h, w + 1 + L2 * (L1 - 1), 2))
        w = L2 - 1
        if pbc:
            e.append((L2 - 1 + L2 * h, L2 * h, 1))
            e.append((w + L2 * h, L2 * ((h + 1) % L1), 2))
            e.append((w + L2 * h, L2 * ((L1 + h - 1) % L1), 2))
        if h < L1 - 1:
            e.append((w + L2 * h, w + L2 * (h + 1), 1))
        elif pbc:
            e.append((w + L2 * h, w, 1))
    return e


class J1J2Operator(Operator):
    mel_dtype = np.complex128   # NetketOperatorWrapper returns complex128 matrix elements

    def __init__(self, hilbert_state_shape, j2=0.5, pbc=False):
        assert len(hilbert_state_shape) == 2
        super(J1J2Operator, self).__init__(hilbert_state_shape)
        self.j2 = j2
        self.pbc = pbc
        self.max_number_of_local_connections = int(np.prod(hilbert_state_shape)) * len(hilbert_state_shape) * 2 + 1
        self.estimated_number_of_local_connections = self.max_number_of_local_connections

    def terms(self):
        J = {1: 1.0, 2: float(self.j2)}
        out = []
        for a, b, colour in j1j2_edges(self.hilbert_state_shape, self.pbc):
            if J[colour] == 0.0:
                continue
            out.append((a, b, _lib.FK_TERM_EXCHANGE, -1, J[colour], 2.0 * J[colour]))
        return out, _lib.FK_OP_J1J2, 1, 0


def j1j2_two_dim_operator(hilbert_state_shape, j2=0.5, pbc=False):
    return J1J2Operator(tuple(hilbert_state_shape), j2=j2, pbc=pbc)


J1J2 = j1j2_two_dim_operator
