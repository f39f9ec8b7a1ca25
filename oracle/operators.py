"""CPU restatement (numpy) of the reference's Hamiltonians' `find_conn`.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Restates (file:line relative to /root/reference/src/flowket):
  * Heisenberg / HeisenbergFindConn   operators/heisenberg.py:6-121
  * Ising.find_conn                    operators/ising.py:17-46
  * J1J2 edge list + netket wrapper    operators/j1j2.py:7-66, operators/netket_operator.py:46-66
    (netket itself is an un-vendored, un-pinned dependency: the per-bond matrix
    sigma.sigma = zz + xx + yy gives diagonal +-J_c and off-diagonal 2*J_c for
    anti-parallel pairs; netket's `get_conn` returns the diagonal entry first and
    then one entry per bond with a non-zero off-diagonal element, in edge order.
    The *order* is "parity unpinned"; values are pinned by exact diagonalisation.)

Common output contract (operators/operator.py:14-21):
  all_conn[C, B, *shape] float64, mel[C, B] float64 (complex128 for J1J2), use_conn[C, B] bool,
  with connection 0 = the sample itself and mel[0] = diagonal element.
"""
import numpy as np


def _as_2d(sample, shape):
    sample = np.asarray(sample)
    if len(shape) == 1:
        return sample.reshape(sample.shape[0], shape[0], 1), (shape[0], 1)
    return sample, tuple(shape)


def heisenberg_find_conn(sample, hilbert_state_shape, pbc=True, unitary_rotation=True):
    """Connection k = 1 + ((i*W + j)*ndim + d): d=0 bond to (i+1, j), d=1 bond to (i, j+1).

    For a 1-D shape (N,) the lattice is (N, 1) and ndim = 1 (only the d=0 bond exists)."""
    shape_in = tuple(hilbert_state_shape)
    ndim = len(shape_in)
    s, (H, W) = _as_2d(sample, shape_in)
    B = s.shape[0]
    off_diag = -2.0 if unitary_rotation else 2.0
    C = H * W * ndim + 1
    conn = np.broadcast_to(s.astype(np.float64)[None], (C, B, H, W)).copy()
    mel = np.zeros((C, B), np.float64)
    use = np.zeros((C, B), bool)
    use[0] = True
    diag = np.zeros(B, np.float64)
    for i in range(H):
        for j in range(W):
            d = 0
            dirs = []
            if H > 1:
                dirs.append((1, 0))
            if W > 1:
                dirs.append((0, 1))
            for (di, dj) in dirs:
                k = 1 + (i * W + j) * ndim + d
                d += 1
                i2, j2 = i + di, j + dj
                if i2 >= H or j2 >= W:
                    if not pbc:
                        continue  # use stays False, mel 0, conn = sample
                    i2, j2 = i2 % H, j2 % W
                a, b = s[:, i, j].astype(np.float64), s[:, i2, j2].astype(np.float64)
                diag += a * b
                use[k] = a != b
                # the two sites are swapped unconditionally (a no-op when equal); the write order
                # matters only in the degenerate extent-1 wrap (i2,j2)==(i,j)
                conn[k, :, i, j] = b
                conn[k, :, i2, j2] = a
                mel[k] = np.where(use[k], off_diag, 0.0)
    mel[0] = diag
    if ndim == 1:
        conn = conn.reshape(C, B, H)
    return conn, mel, use


def ising_find_conn(sample, hilbert_state_shape, h=1.0, j=1.0, pbc=True):
    """Connection k = 1 + flat site index: that site flipped, mel = -h; all connections used.

    mel[0] = -j * sum_sites s(site) * (s(down) + s(right)), neighbours wrapped (pbc) or 0 (obc);
    the reference accumulates this in float32 (operators/ising.py:25)."""
    shape_in = tuple(hilbert_state_shape)
    s = np.asarray(sample)
    B = s.shape[0]
    N = int(np.prod(shape_in))
    conn = np.broadcast_to(s[None], (N + 1,) + s.shape).copy()
    flat = conn.reshape(N + 1, B, N)
    for site in range(N):
        flat[site + 1, :, site] *= -1
    pad = ((0, 0),) + ((0, 1),) * len(shape_in)
    padded = np.pad(s, pad, mode='wrap') if pbc else np.pad(s, pad, mode='constant', constant_values=0)
    acc = np.zeros(s.shape, np.float32)
    if len(shape_in) == 1:
        for i in range(shape_in[0]):
            acc[:, i] -= j * s[:, i] * padded[:, i + 1]
    else:
        H, W = shape_in
        for i in range(H):
            for jj in range(W):
                if H > 1:
                    acc[:, i, jj] -= j * s[:, i, jj] * padded[:, i + 1, jj]
                if W > 1:
                    acc[:, i, jj] -= j * s[:, i, jj] * padded[:, i, jj + 1]
    mel = np.full((N + 1, B), -h, dtype=np.float64)
    mel[0] = acc.sum(axis=tuple(range(1, s.ndim)))
    return conn, mel, np.ones((N + 1, B), bool)


def j1j2_edges(hilbert_state_shape, pbc=False):
    """Edge list [(site_a, site_b, colour)] in the order of operators/j1j2.py:14-35 (colour 1 = NN, 2 = NNN)."""
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


def j1j2_find_conn(sample, hilbert_state_shape, j2=0.5, pbc=False):
    """Compacted layout of NetketOperatorWrapper.new_netket_find_conn (netket_operator.py:46-66):
    per sample, slot 0 = self, slots 1..n = one per anti-parallel bond in edge order (mel = 2*J_c),
    remaining slots zero-filled with mel = 0, use = (mel != 0), use[0] = True."""
    shape = tuple(hilbert_state_shape)
    s = np.asarray(sample)
    B = s.shape[0]
    N = int(np.prod(shape))
    C = N * len(shape) * 2 + 1
    J = {1: 1.0, 2: float(j2)}
    edges = j1j2_edges(shape, pbc)
    flat = s.reshape(B, N).astype(np.float64)
    conn = np.zeros((C, B, N), np.float64)
    mel = np.zeros((C, B), np.complex128)
    conn[0] = flat
    fill = np.ones(B, np.int64)
    rows = np.arange(B)
    for (a, b, col) in edges:
        prod = flat[:, a] * flat[:, b]
        mel[0] += J[col] * prod
        anti = (prod < 0) & (J[col] != 0.0)
        idx = rows[anti]
        slot = fill[anti]
        new = flat[idx].copy()
        new[:, a], new[:, b] = flat[idx, b], flat[idx, a]
        conn[slot, idx] = new
        mel[slot, idx] = 2.0 * J[col]
        fill[anti] += 1
    use = mel != 0.0
    use[0] = True
    return conn.reshape((C, B) + shape), mel, use


class OracleOperator(object):
    """Small object exposing the reference Operator protocol on top of the functions above."""

    def __init__(self, kind, hilbert_state_shape, **kw):
        self.kind, self.kw = kind, kw
        self.hilbert_state_shape = tuple(hilbert_state_shape)
        n = int(np.prod(self.hilbert_state_shape))
        self.max_number_of_local_connections = {
            'heisenberg': n * len(self.hilbert_state_shape) + 1, 'ising': n + 1,
            'j1j2': n * len(self.hilbert_state_shape) * 2 + 1}[kind]

    def find_conn(self, sample):
        f = {'heisenberg': heisenberg_find_conn, 'ising': ising_find_conn, 'j1j2': j1j2_find_conn}[self.kind]
        return f(sample, self.hilbert_state_shape, **self.kw)
