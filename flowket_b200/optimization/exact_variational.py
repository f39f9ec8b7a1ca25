"""ExactVariational / ExactObservable: full enumeration of the 2^N states
(flowket/optimization/exact_variational.py:10-161).  log psi of all states comes from the CUDA forward
(model.predict); connection *indices* come from the device find_conn; the O(C * 2^N) gathers stay in numpy
exactly like the reference (this is the reference's CPU-runnable configuration, cfg 1)."""
import time

import numpy as np

from ..exact.utils import binary_array_to_decimal_array, decimal_array_to_binary_array, fsum, \
    complex_norm_log_fsum_exp


class ExactObservable(object):
    def __init__(self, exact_variational, operator, calculate_variance_of_the_local_operator=False):
        self.exact_variational = exact_variational
        self.operator = operator
        self.calculate_variance_of_the_local_operator = calculate_variance_of_the_local_operator
        ev = exact_variational
        self.energies = np.zeros((ev.num_of_states,), dtype=np.complex128)
        self.naive_energies = np.zeros_like(self.energies)
        self._build_local_connections()

    def _build_local_connections(self):
        ev = self.exact_variational
        C = self.operator.max_number_of_local_connections
        if C is None:
            C = self.calculate_max_number_of_local_connections()
        self.states_idx_local_connections = np.zeros((C, ev.num_of_states), dtype=np.int64)
        self.states_hamiltonian_values = np.zeros((C, ev.num_of_states), dtype=np.complex128)
        for i in range(0, ev.num_of_states, ev.batch_size):
            conn, mel, _use = self.operator.find_conn(ev.states[i:i + ev.batch_size, ...])
            self.states_hamiltonian_values[:conn.shape[0], i:i + ev.batch_size] = mel
            bits = (conn.reshape(conn.shape[0] * conn.shape[1], ev.number_of_spins) + 1) // 2
            self.states_idx_local_connections[:conn.shape[0], i:i + ev.batch_size] = \
                binary_array_to_decimal_array(bits).reshape(conn.shape[0], -1)

    def calculate_max_number_of_local_connections(self):
        ev = self.exact_variational
        return max(len(self.operator.find_conn(ev.states[i:i + ev.batch_size, ...])[0])
                   for i in range(0, ev.num_of_states, ev.batch_size))

    def update_local_energy(self):
        ev = self.exact_variational
        for i in range(0, ev.num_of_states, ev.batch_size):
            sl = slice(i, i + ev.batch_size)
            log_values = ev.wave_function[self.states_idx_local_connections[:, sl]]
            val_mult = np.exp(np.conj(log_values) + log_values[0, :])
            self.energies[sl] = (np.conj(self.states_hamiltonian_values[:, sl]) * val_mult).sum(axis=0) \
                / ev.wave_function_norm_squared
            if self.calculate_variance_of_the_local_operator:
                self.naive_energies[sl] = (self.states_hamiltonian_values[:, sl]
                                           * np.exp(log_values - log_values[0, :])).sum(axis=0)
        self.current_energy = fsum(self.energies)
        if self.calculate_variance_of_the_local_operator:
            d = np.real(self.naive_energies - self.current_energy)
            self.current_local_energy_variance = float(fsum(d * d * ev.probs))


class ExactVariational(object):
    def __init__(self, model, operator, batch_size):
        self.model = model
        self.operator = operator
        self.wave_function_callable = lambda states: [model.predict(states[0])]
        self._build_wave_function_arrays(tuple(model.input_shape[1:]))
        self._set_batch_size(batch_size)
        self.energy_observable = ExactObservable(self, operator, calculate_variance_of_the_local_operator=True)

    def _build_wave_function_arrays(self, input_size):
        self.input_size = input_size
        self.number_of_spins = int(np.prod(self.input_size))
        self.num_of_states = 2 ** self.number_of_spins
        self.wave_function = np.zeros((self.num_of_states,), dtype=np.complex128)
        self.psi_squared = np.zeros_like(self.wave_function)
        self.probs = np.zeros((self.num_of_states,), dtype=np.float64)
        self.log_probs = np.zeros_like(self.probs)
        self.probs_mult_energy_mean = np.zeros_like(self.wave_function)
        self.energy_grad_coefficients = np.zeros_like(self.wave_function)
        self.states = decimal_array_to_binary_array(np.arange(self.num_of_states), self.number_of_spins, False) \
            .reshape((self.num_of_states,) + self.input_size)
        self.wave_function_norm_squared = None

    def _set_batch_size(self, batch_size):
        if batch_size > self.num_of_states:
            batch_size = self.num_of_states
        if self.num_of_states % batch_size != 0:
            raise Exception('In exact the batch size must divide the total number of states in the system')
        self.batch_size = batch_size
        self.num_of_batch_until_full_cycle = self.num_of_states // self.batch_size

    def _update_wave_function_arrays(self):
        for i in range(0, self.num_of_states, self.batch_size):
            self.wave_function[i:i + self.batch_size] = \
                self.wave_function_callable([self.states[i:i + self.batch_size, ...]])[0][:, 0]
        np.multiply(self.wave_function, 2.0, out=self.psi_squared)
        log_norm = complex_norm_log_fsum_exp(self.psi_squared)
        self.wave_function_norm_squared = np.exp(log_norm)
        np.subtract(np.real(self.psi_squared), log_norm, out=self.log_probs)
        np.exp(self.log_probs, out=self.probs)

    def _update_local_energy(self):
        self.energy_observable.update_local_energy()
        np.multiply(self.probs.astype(np.complex128), self.energy_observable.current_energy,
                    out=self.probs_mult_energy_mean)
        np.subtract(self.energy_observable.energies, self.probs_mult_energy_mean, out=self.energy_grad_coefficients)

    def machine_updated(self):
        self.machine_updated_start_time = time.time()
        self._update_wave_function_arrays()
        self.wave_function_update_end_time = time.time()
        self._update_local_energy()
        self.local_energy_update_end_time = time.time()

    def to_generator(self):
        while True:
            self.machine_updated()
            for i in range(0, self.num_of_states, self.batch_size):
                yield self.states[i:i + self.batch_size], self.energy_grad_coefficients[i:i + self.batch_size]

    def __iter__(self):
        return self.to_generator()
