"""Heisenberg: constructor and semantics of flowket/operators/heisenberg.py:6-121.
H = sum_<ab> sigma_a.sigma_b (Pauli matrices); connection k = 1 + ((i*W + j)*ndim + d) exchanges the two sites of
the bond to (i+1, j) (d = 0) or (i, j+1) (d = 1); matrix element -2 with the Marshall rotation, else +2."""
import numpy as np

from .operator import OperatorOnGrid
from .. import _lib


class Heisenberg(OperatorOnGrid):
    def __init__(self, total_sz=0.0, unitary_rotation=True, **kwargs):
        super(Heisenberg, self).__init__(**kwargs)
        self.off_diag = -2.0 if unitary_rotation else 2.0
        self.total_sz = total_sz
        self.total_size = int(np.prod(self.hilbert_state_shape))
        self.dim = len(self.hilbert_state_shape)
        assert self.dim <= 2
        self.max_number_of_local_connections = self.total_size * self.dim + 1

    def terms(self):
        shape = self.hilbert_state_shape
        H, W = (shape[0], 1) if self.dim == 1 else shape   # 1-D lattices are (N, 1) (heisenberg.py:58-59)
        out = []
        for i in range(H):
            for j in range(W):
                d = 0
                dirs = ([(1, 0)] if H > 1 else []) + ([(0, 1)] if W > 1 else [])
                for di, dj in dirs:
                    slot = 1 + (i * W + j) * self.dim + d
                    d += 1
                    i2, j2 = i + di, j + dj
                    if i2 >= H or j2 >= W:
                        if not self.pbc:
                            continue
                        i2, j2 = i2 % H, j2 % W
                    out.append((i * W + j, i2 * W + j2, _lib.FK_TERM_EXCHANGE, slot, 1.0, self.off_diag))
        return out, _lib.FK_OP_HEISENBERG, 0, 0

    def random_states(self, num_of_states):
        if self.total_sz is not None:
            size = self.total_size // 2
            states = np.zeros((num_of_states, self.total_size))
            base = np.concatenate([np.ones(size + int(self.total_sz)), np.full(size - int(self.total_sz), -1)])
            for i in range(num_of_states):
                states[i] = np.random.permutation(base)
            return states.reshape((num_of_states,) + self.hilbert_state_shape)
        return super(Heisenberg, self).random_states(num_of_states)

    def use_state(self, state):
        if self.total_sz is None:
            return True
        return state.sum() == self.total_sz
