"""ExactVariational with the 2^N states sharded over the ranks of torch.distributed ("ExactVariational
shards the 2^N states the same way").  The reference enumerates everything in one process
(flowket/optimization/exact_variational.py:75-161); here rank r owns the contiguous slice [r 2^N / G, (r+1) 2^N / G):

  * log psi of the slice comes from this rank's GPU (model.predict); one allreduce of the zero-padded table gives every rank the
    whole wave function (2^N complex numbers -- the connected states of a slice live anywhere in the table), from which the
    norm, probabilities and log-probabilities follow redundantly on every rank;
  * find_conn, the connection index table and the local energies are built for the slice only (the C x 2^N tables are the
    memory hog of the exact path: they shrink by G);
  * <H> and the variance are sums over slices: one allreduce of three numbers;
  * the generator yields the slice's (states, coefficients) mini-batches; with Trainer(distributed=True) or
    convert_to_accumulate_gradient_optimizer(use_horovod=True) the accumulated gradients are summed over the ranks, which is
    the exact gradient of the whole enumeration.

Every rank runs `machine_updated()` at the same time (it contains collectives).  Without an initialised process group the
class degenerates to ExactVariational."""
import time

import numpy as np

from .exact_variational import ExactVariational, ExactObservable
from ..exact.utils import fsum, complex_norm_log_fsum_exp


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


def _allreduce_numpy(array):
    """sum of a numpy array over the ranks (fp64 / complex128 kept exactly as sums of the per-rank values)"""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return array
    import torch
    arr = np.ascontiguousarray(array)
    as_real = arr.view(np.float64) if np.iscomplexobj(arr) else arr
    t = torch.from_numpy(as_real.copy())
    if dist.get_backend() == 'nccl':
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    out = t.cpu().numpy()
    return out.view(np.complex128).reshape(arr.shape) if np.iscomplexobj(arr) else out.reshape(arr.shape)


class DistributedExactObservable(ExactObservable):
    """ExactObservable over this rank's slice of the states; `current_energy` / variance are global."""

    def _build_local_connections(self):
        ev = self.exact_variational
        C = self.operator.max_number_of_local_connections
        if C is None:
            C = int(_allreduce_numpy(np.array([float(self.calculate_max_number_of_local_connections())]))[0])   # upper bound
        n_local = ev.slice_hi - ev.slice_lo
        self.states_idx_local_connections = np.zeros((C, n_local), dtype=np.int64)
        self.states_hamiltonian_values = np.zeros((C, n_local), dtype=np.complex128)
        from ..exact.utils import binary_array_to_decimal_array
        for i in range(0, n_local, ev.batch_size):
            conn, mel, _use = self.operator.find_conn(ev.states[ev.slice_lo + i:ev.slice_lo + i + ev.batch_size, ...])
            self.states_hamiltonian_values[:conn.shape[0], i:i + ev.batch_size] = mel
            bits = (conn.reshape(conn.shape[0] * conn.shape[1], ev.number_of_spins) + 1) // 2
            self.states_idx_local_connections[:conn.shape[0], i:i + ev.batch_size] = \
                binary_array_to_decimal_array(bits).reshape(conn.shape[0], -1)

    def calculate_max_number_of_local_connections(self):
        ev = self.exact_variational
        return max(len(self.operator.find_conn(ev.states[i:i + ev.batch_size, ...])[0])
                   for i in range(ev.slice_lo, ev.slice_hi, ev.batch_size))

    def update_local_energy(self):
        ev = self.exact_variational
        lo = ev.slice_lo
        n_local = ev.slice_hi - lo
        for i in range(0, n_local, ev.batch_size):
            sl = slice(i, i + ev.batch_size)
            log_values = ev.wave_function[self.states_idx_local_connections[:, sl]]
            val_mult = np.exp(np.conj(log_values) + log_values[0, :])
            self.energies[lo + i:lo + i + ev.batch_size] = \
                (np.conj(self.states_hamiltonian_values[:, sl]) * val_mult).sum(axis=0) / ev.wave_function_norm_squared
            if self.calculate_variance_of_the_local_operator:
                self.naive_energies[lo + i:lo + i + ev.batch_size] = \
                    (self.states_hamiltonian_values[:, sl] * np.exp(log_values - log_values[0, :])).sum(axis=0)
        local = slice(lo, ev.slice_hi)
        self.current_energy = complex(_allreduce_numpy(np.array([fsum(self.energies[local])], dtype=np.complex128))[0])
        if self.calculate_variance_of_the_local_operator:
            d = np.real(self.naive_energies[local] - self.current_energy)
            self.current_local_energy_variance = float(_allreduce_numpy(np.array([fsum(d * d * ev.probs[local])]))[0])


class DistributedExactVariational(ExactVariational):
    def __init__(self, model, operator, batch_size):
        dist = _dist()
        self.world_size = dist.get_world_size() if dist else 1
        self.rank = dist.get_rank() if dist else 0
        self.model = model
        self.operator = operator
        self.wave_function_callable = lambda states: [model.predict(states[0])]
        self._build_wave_function_arrays(tuple(model.input_shape[1:]))
        if self.num_of_states % self.world_size != 0:
            raise Exception('the number of ranks must divide the total number of states in the system')
        per_rank = self.num_of_states // self.world_size
        self.slice_lo, self.slice_hi = self.rank * per_rank, (self.rank + 1) * per_rank
        self._set_batch_size(batch_size)
        self.energy_observable = DistributedExactObservable(self, operator, calculate_variance_of_the_local_operator=True)

    def _set_batch_size(self, batch_size):
        per_rank = self.num_of_states // self.world_size
        if batch_size > per_rank:
            batch_size = per_rank
        if per_rank % batch_size != 0:
            raise Exception('In exact the batch size must divide the number of states of a rank (%d)' % per_rank)
        self.batch_size = batch_size
        self.num_of_batch_until_full_cycle = per_rank // self.batch_size      # mini-batches of *this rank* per enumeration

    def _update_wave_function_arrays(self):
        self.wave_function[:] = 0
        for i in range(self.slice_lo, self.slice_hi, self.batch_size):
            self.wave_function[i:i + self.batch_size] = \
                self.wave_function_callable([self.states[i:i + self.batch_size, ...]])[0][:, 0]
        self.wave_function[:] = _allreduce_numpy(self.wave_function)          # disjoint slices: the sum is the gather
        np.multiply(self.wave_function, 2.0, out=self.psi_squared)
        log_norm = complex_norm_log_fsum_exp(self.psi_squared)
        self.wave_function_norm_squared = np.exp(log_norm)
        np.subtract(np.real(self.psi_squared), log_norm, out=self.log_probs)
        np.exp(self.log_probs, out=self.probs)

    def _update_local_energy(self):
        self.energy_observable.update_local_energy()
        local = slice(self.slice_lo, self.slice_hi)
        self.energy_grad_coefficients[:] = 0
        self.energy_grad_coefficients[local] = self.energy_observable.energies[local] - \
            self.probs[local] * self.energy_observable.current_energy

    def machine_updated(self):
        self.machine_updated_start_time = time.time()
        self._update_wave_function_arrays()
        self.wave_function_update_end_time = time.time()
        self._update_local_energy()
        self.local_energy_update_end_time = time.time()

    def to_generator(self):
        while True:
            self.machine_updated()
            for i in range(self.slice_lo, self.slice_hi, self.batch_size):
                yield self.states[i:i + self.batch_size], self.energy_grad_coefficients[i:i + self.batch_size]
