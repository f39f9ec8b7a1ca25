"""CPU restatement (numpy) of the reference's exact-enumeration path.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Restates (relative to /root/reference/src/flowket):
  * bit conventions (bit k of the state index <-> flattened site k, bit 1 <-> +1)   exact/utils.py:39-58
  * complex_norm_log_fsum_exp                                                      exact/utils.py:75-78
  * vector_to_machine                                                              exact/utils.py:95-99
  * ExactObservable._build_local_connections / update_local_energy                 optimization/exact_variational.py:44-90
  * ExactVariational._update_wave_function_arrays / _update_local_energy           optimization/exact_variational.py:132-144
"""
import numpy as np
import scipy.sparse
import scipy.sparse.linalg


def all_states(number_of_spins):
    idx = np.arange(2 ** number_of_spins, dtype=np.int64)
    return np.stack([2 * ((idx >> k) & 1) - 1 for k in range(number_of_spins)], axis=1).astype(np.int8)


def states_to_index(states):
    flat = np.asarray(states).reshape(states.shape[0], -1)
    weights = (1 << np.arange(flat.shape[1], dtype=np.int64))
    return ((flat == 1) * weights).sum(axis=1)


def vector_to_machine(log_wave_function_vector):
    def machine(batch):
        return log_wave_function_vector[states_to_index(np.asarray(batch))][..., None]
    return machine


class ExactVariationalOracle(object):
    def __init__(self, log_psi_fn, operator, input_shape, batch_size):
        """log_psi_fn(states[n,*shape]) -> complex ndarray [n]"""
        self.log_psi_fn, self.operator = log_psi_fn, operator
        self.input_shape = tuple(input_shape)
        self.number_of_spins = int(np.prod(self.input_shape))
        self.num_of_states = 2 ** self.number_of_spins
        batch_size = min(batch_size, self.num_of_states)
        if self.num_of_states % batch_size != 0:
            raise Exception('In exact the batch size must divide the total number of states in the system')
        self.batch_size = batch_size
        self.states = all_states(self.number_of_spins).reshape((self.num_of_states,) + self.input_shape)
        C = operator.max_number_of_local_connections
        self.idx_conn = np.zeros((C, self.num_of_states), np.int64)
        self.ham = np.zeros((C, self.num_of_states), np.complex128)
        for i in range(0, self.num_of_states, batch_size):
            conn, mel, _use = operator.find_conn(self.states[i:i + batch_size])
            self.ham[:conn.shape[0], i:i + batch_size] = mel
            c = conn.reshape(conn.shape[0] * conn.shape[1], -1)
            self.idx_conn[:conn.shape[0], i:i + batch_size] = states_to_index(c).reshape(conn.shape[0], -1)

    def machine_updated(self):
        self.wave_function = np.zeros(self.num_of_states, np.complex128)
        for i in range(0, self.num_of_states, self.batch_size):
            self.wave_function[i:i + self.batch_size] = self.log_psi_fn(self.states[i:i + self.batch_size])
        psi_squared = 2.0 * self.wave_function
        re = np.real(psi_squared)
        m = re.max()
        log_norm = np.log(np.sum(np.exp(re - m))) + m
        self.wave_function_norm_squared = np.exp(log_norm)
        self.log_probs = re - log_norm
        self.probs = np.exp(self.log_probs)
        log_values = self.wave_function[self.idx_conn]                     # [C, S]
        val = np.exp(np.conj(log_values) + log_values[0])
        self.energies = (np.conj(self.ham) * val).sum(axis=0) / self.wave_function_norm_squared
        self.current_energy = self.energies.sum()
        naive = (self.ham * np.exp(log_values - log_values[0])).sum(axis=0)
        self.naive_energies = naive
        self.current_local_energy_variance = float(
            (np.real(naive - self.current_energy) ** 2 * self.probs).sum())
        self.energy_grad_coefficients = self.energies - self.probs * self.current_energy
        return self.current_energy


def sparse_hamiltonian(operator, input_shape, batch_size=4096):
    """H as a scipy CSR matrix built from find_conn over all 2^N states (row = state, col = connection)."""
    n = int(np.prod(input_shape))
    S = 2 ** n
    states = all_states(n).reshape((S,) + tuple(input_shape))
    rows, cols, vals = [], [], []
    for i in range(0, S, batch_size):
        conn, mel, use = operator.find_conn(states[i:i + batch_size])
        C, B = mel.shape
        idx = states_to_index(conn.reshape(C * B, -1)).reshape(C, B)
        r = np.broadcast_to(np.arange(i, i + B)[None], (C, B))
        m = use & (mel != 0)
        rows.append(r[m]); cols.append(idx[m]); vals.append(np.real(mel[m]))
    return scipy.sparse.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(S, S))


def ground_state(operator, input_shape):
    Hm = sparse_hamiltonian(operator, input_shape)
    w, v = scipy.sparse.linalg.eigsh(Hm, k=1, which='SA')
    return float(w[0]), v[:, 0]
