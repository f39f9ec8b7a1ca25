"""Ising: constructor and semantics of flowket/operators/ising.py:8-46.
H = -j sum sz sz - h sum sx; connection k = 1 + flat site index flips that site (mel = -h); all connections used;
mel[0] accumulated in float32 like the reference."""
import numpy as np

from .operator import OperatorOnGrid
from .. import _lib


class Ising(OperatorOnGrid):
    def __init__(self, h=1.0, j=1.0, **kwargs):
        super(Ising, self).__init__(**kwargs)
        self.h = h
        self.j = j
        assert len(self.hilbert_state_shape) <= 2
        self.max_number_of_local_connections = int(np.prod(self.hilbert_state_shape)) + 1

    def terms(self):
        shape = self.hilbert_state_shape
        out = []
        if len(shape) == 1:
            N = shape[0]
            for i in range(N):
                nb = i + 1 if i + 1 < N else (0 if self.pbc else -1)
                out.append((i, nb, _lib.FK_TERM_DIAG, -1, -self.j, 0.0))
        else:
            H, W = shape
            for i in range(H):
                for j in range(W):
                    if H > 1:
                        nb = (i + 1) * W + j if i + 1 < H else (j if self.pbc else -1)
                        out.append((i * W + j, nb, _lib.FK_TERM_DIAG, -1, -self.j, 0.0))
                    if W > 1:
                        nb = i * W + j + 1 if j + 1 < W else (i * W if self.pbc else -1)
                        out.append((i * W + j, nb, _lib.FK_TERM_DIAG, -1, -self.j, 0.0))
        for site in range(int(np.prod(shape))):
            out.append((site, -1, _lib.FK_TERM_FLIP, 1 + site, 0.0, -self.h))
        return out, _lib.FK_OP_ISING, 0, 1
