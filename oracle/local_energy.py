"""CPU restatement (numpy) of the reference's Monte-Carlo local-energy estimator.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Restates (relative to /root/reference/src/flowket):
  * get_flat_local_connections_log_values / Observable.local_values*   observables/monte_carlo/operator.py:6-54
  * BaseObservable.estimate                                           observables/monte_carlo/observable.py:10-14
  * VariationalMonteCarlo.loss_coefficients / next_batch               optimization/variational_monte_carlo.py:38-50
"""
import numpy as np


def flat_used_connections(local_connections, all_use_conn):
    """Sample-major ragged list of the used connections (self first)."""
    moved = np.moveaxis(local_connections, 1, 0)
    flat = moved.reshape((-1,) + local_connections.shape[2:])
    return flat[all_use_conn.astype(bool).T.flatten()]


def local_values_unbalanced(wave_function, local_connections, hamiltonian_values, all_use_conn):
    use = all_use_conn.astype(bool)
    B = use.shape[1]
    log_values = wave_function(flat_used_connections(local_connections, use))[:, 0]
    counts = use.sum(axis=0).astype(np.int64)
    out = np.zeros(B, np.complex128)
    start = 0
    for b in range(B):
        seg = log_values[start:start + counts[b]]
        start += counts[b]
        ratio = np.exp(seg - seg[0])
        out[b] = np.multiply(hamiltonian_values[use[:, b], b], ratio).sum()
    return out


def local_values_balanced(wave_function, local_connections, hamiltonian_values):
    C, B = local_connections.shape[:2]
    flat = local_connections.reshape((C * B,) + local_connections.shape[2:])
    log_values = wave_function(flat)[:, 0].reshape(C, B)
    return np.multiply(hamiltonian_values, np.exp(log_values - log_values[0])).sum(axis=0)


def local_values(operator, wave_function, configurations):
    conn, mel, use = operator.find_conn(configurations)
    if use.mean() < 0.95:
        return local_values_unbalanced(wave_function, conn, mel, use)
    return local_values_balanced(wave_function, conn, mel)


def estimate(operator, wave_function, configurations):
    lv = local_values(operator, wave_function, configurations)
    return np.mean(lv), np.var(np.real(lv)), lv


def loss_coefficients(local_energy, mean_energy, batch_size):
    """y_b = conj(E_loc_b - E_mean) / B (variational_monte_carlo.py:38-40,50)."""
    return np.conj(local_energy - mean_energy) / batch_size
