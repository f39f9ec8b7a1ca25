"""Exact enumeration of the 2^N basis states (BASELINE configs[0]): `ExactVariational` / `ExactObservable` with the
attribute contract of flowket/optimization/exact_variational.py:10-161, built around DEVICE-RESIDENT tables.

Layout.  The log psi table (complex128 [2^N]), the probabilities and, per observable, the connection tables -- index of
every connected state and its matrix element, [C, n] -- live in HBM as torch tensors.  One machine update is

    forward of the owned states (model.predict_device, CUDA)  ->  table           (all-gather when the states are sharded)
    log-sum-exp of 2 Re log psi (fp64)                          ->  norm, probs
    fk_exact_local_energy: one thread per state gathers its column of the index table from the psi table and sums
    the probability-weighted and the naive local energy (fp64)  ->  energies
    three fp64 sums                                             ->  <H>, variance, gradient coefficients

and nothing crosses PCIe except the scalars.  The connection tables are built once by fk_exact_conn_table, which emits
indices straight from the operator's term table (the reference materialises every connected configuration on the host
and converts it back to an index).  numpy views of the reference's public attributes (`probs`, `wave_function`,
`energy_grad_coefficients`, `energies`, ...) are read-through properties that copy on first use after an update.

Sharding.  With `rank`/`world_size` > 1 (see distributed_exact_variational.py) a process owns
the contiguous slice [slice_lo, slice_hi) of the states: it evaluates and keeps connection tables for that slice
only, the psi table is all-gathered, and <H> / variance are all-reduced.

Models without a device route (any object with `input_shape` and `predict`, e.g. a tabulated wave function) and
operators without a device term table run the same pipeline on CPU tensors with the gather written as torch indexing;
that is how the reference class is used with arbitrary callables and how the host-side tests pin this class against
the reference's own (tests/test_host_logic.py)."""
import ctypes
import time

import numpy as np
import torch

from .. import _lib


def _process_group():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) else None


def _collective_device(tensor_device):
    """NCCL needs CUDA tensors, gloo CPU tensors"""
    dist = _process_group()
    if dist is None:
        return tensor_device
    return torch.device('cuda') if dist.get_backend() == 'nccl' else torch.device('cpu')


def _all_reduce_sum(t, world_size):
    """sum over the ranks when the enumeration is sharded (world_size > 1); the identity otherwise"""
    dist = _process_group()
    if dist is None or world_size == 1:
        return t
    buf = t.to(_collective_device(t.device)).contiguous()
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return buf.to(t.device)


class _Cached(object):
    """read-through numpy view of a tensor attribute: `name` -> owner._t[name] copied to the host once per update"""

    def __init__(self, name):
        self.name = name

    def __get__(self, owner, owner_type=None):
        if owner is None:
            return self
        host = owner._host
        if self.name not in host:
            t = owner._t[self.name]
            host[self.name] = t.detach().cpu().numpy() if t is not None else None
        return host[self.name]


class ExactObservable(object):
    """<O> over the full enumeration for any operator: connection tables on the device, per-state energies from the
    owner's psi table (exact_variational.py:10-66)."""

    energies = _Cached('energies')                    # p(s) conj(O_loc(s)); zero outside the owned slice
    naive_energies = _Cached('naive_energies')        # O_loc(s)
    states_idx_local_connections = _Cached('index')   # [C, owned states]
    states_hamiltonian_values = _Cached('mel')

    def __init__(self, exact_variational, operator, calculate_variance_of_the_local_operator=False):
        self.exact_variational = exact_variational
        self.operator = operator
        self.calculate_variance_of_the_local_operator = calculate_variance_of_the_local_operator
        self.current_energy = None
        self.current_local_energy_variance = None
        self._host = {}
        self._t = {'energies': None, 'naive_energies': None}
        self._t['index'], self._t['mel'] = self._connection_tables()

    # ---- built once ------------------------------------------------------------------------------------------------
    def _connection_tables(self):
        ev = self.exact_variational
        n = ev.slice_hi - ev.slice_lo
        if ev.device.type == 'cuda' and hasattr(self.operator, 'device_desc'):
            if self.operator.max_number_of_local_connections is None:
                raise ValueError('the operator does not state max_number_of_local_connections')
            lib = _lib.require_cuda()
            desc = self.operator.device_desc()
            C = int(desc.max_conn)
            index = torch.empty((C, n), dtype=torch.int64, device=ev.device)
            mel = torch.empty((C, n), dtype=torch.float64, device=ev.device)
            _lib.check(lib.fk_exact_conn_table(ctypes.byref(desc), ev.slice_lo, n, index.data_ptr(), mel.data_ptr(),
                                               _lib.stream_ptr()))
            return index, mel
        # host protocol: find_conn on windows of the owned states, configurations -> indices with torch
        C = self.operator.max_number_of_local_connections
        windows = [self._host_window(lo) for lo in range(ev.slice_lo, ev.slice_hi, ev.batch_size)]
        if C is None:
            C = max(idx.shape[0] for idx, _ in windows)
            if ev.world_size > 1:
                C = int(_all_reduce_sum(torch.tensor([float(C)], dtype=torch.float64), ev.world_size)[0].item())   # an upper bound
        index = torch.zeros((C, n), dtype=torch.int64, device=ev.device)
        complex_mel = any(torch.is_complex(m) for _, m in windows)
        mel = torch.zeros((C, n), dtype=torch.complex128 if complex_mel else torch.float64, device=ev.device)
        for w, (idx, m) in enumerate(windows):
            cols = slice(w * ev.batch_size, w * ev.batch_size + idx.shape[1])
            index[:idx.shape[0], cols] = idx
            mel[:m.shape[0], cols] = m
        return index, mel

    def _host_window(self, lo):
        ev = self.exact_variational
        conn, mel, _use = self.operator.find_conn(ev.states[lo:lo + ev.batch_size])
        conn = torch.as_tensor(np.asarray(conn))
        bits = (conn.reshape(conn.shape[0], conn.shape[1], ev.number_of_spins) == 1).to(torch.int64)
        index = (bits << torch.arange(ev.number_of_spins, dtype=torch.int64)).sum(dim=-1)
        mel = torch.as_tensor(np.asarray(mel))
        mel = mel.to(torch.complex128) if torch.is_complex(mel) and bool((mel.imag != 0).any()) else \
            (mel.real if torch.is_complex(mel) else mel).to(torch.float64)
        return index.to(ev.device), mel.to(ev.device)

    def calculate_max_number_of_local_connections(self):
        return int(self._t['index'].shape[0])

    # ---- every update ----------------------------------------------------------------------------------------------
    def update_local_energy(self):
        ev = self.exact_variational
        table = ev._t['wave_function']
        index, mel = self._t['index'], self._t['mel']
        n = index.shape[1]
        want_naive = self.calculate_variance_of_the_local_operator
        if table.device.type == 'cuda' and not torch.is_complex(mel):
            lib = _lib.require_cuda()
            weighted = torch.empty(n, dtype=torch.complex128, device=table.device)
            naive = torch.empty(n, dtype=torch.complex128, device=table.device) if want_naive else None
            _lib.check(lib.fk_exact_local_energy(table.data_ptr(), index.data_ptr(), mel.data_ptr(), index.shape[0], n,
                                                 float(ev._log_norm), weighted.data_ptr(),
                                                 naive.data_ptr() if want_naive else None, _lib.stream_ptr()))
        else:
            gathered = table[index]                                   # [C, n]
            own = gathered[0]
            weighted = (mel.conj() * torch.exp(gathered.conj() + own - ev._log_norm)).sum(dim=0)
            naive = (mel * torch.exp(gathered - own)).sum(dim=0) if want_naive else None
        owned = slice(ev.slice_lo, ev.slice_hi)
        full = torch.zeros(ev.num_of_states, dtype=torch.complex128, device=table.device)
        full[owned] = weighted
        self._t['energies'] = full
        totals = [weighted.real.sum(), weighted.imag.sum()]
        self.current_energy = complex(*[float(v) for v in _all_reduce_sum(torch.stack(totals), ev.world_size)])
        if want_naive:
            full_naive = torch.zeros_like(full)
            full_naive[owned] = naive
            self._t['naive_energies'] = full_naive
            spread = (naive.real - self.current_energy.real) ** 2 * ev._t['probs'][owned]
            self.current_local_energy_variance = float(_all_reduce_sum(spread.sum().reshape(1), ev.world_size)[0])
        self._host = {k: v for k, v in self._host.items() if k in ('index', 'mel')}


class ExactVariational(object):
    """Generator of (states, gradient coefficients) mini-batches over the whole enumeration with exact <H> and
    variance (exact_variational.py:69-161).  `rank` / `world_size` shard the states (default: everything here)."""

    wave_function = _Cached('wave_function')                          # log psi of all states
    probs = _Cached('probs')
    log_probs = _Cached('log_probs')
    energy_grad_coefficients = _Cached('energy_grad_coefficients')    # zero outside the owned slice

    def __init__(self, model, operator, batch_size, rank=0, world_size=1):
        self.model = model
        self.operator = operator
        self.rank, self.world_size = int(rank), int(world_size)
        self.input_size = tuple(model.input_shape[1:])
        self.number_of_spins = int(np.prod(self.input_size))
        self.num_of_states = 2 ** self.number_of_spins
        if self.num_of_states % self.world_size != 0:
            raise Exception('the number of ranks must divide the total number of states in the system')
        per_rank = self.num_of_states // self.world_size
        self.slice_lo, self.slice_hi = self.rank * per_rank, (self.rank + 1) * per_rank
        batch_size = min(int(batch_size), per_rank)
        if per_rank % batch_size != 0:
            raise Exception('In exact the batch size must divide the total number of states in the system'
                            if self.world_size == 1 else
                            'In exact the batch size must divide the number of states of a rank (%d)' % per_rank)
        self.batch_size = batch_size
        self.num_of_batch_until_full_cycle = per_rank // batch_size      # mini-batches of THIS process per enumeration
        self.on_device = hasattr(model, 'predict_device')
        self.device = torch.device('cuda') if self.on_device else torch.device('cpu')
        self._host = {}
        self._t = {'wave_function': None, 'probs': None, 'log_probs': None, 'energy_grad_coefficients': None}
        self._log_norm = None
        self.wave_function_norm_squared = None
        self._sigma = self._enumerate(self.slice_lo, self.slice_hi)      # int8 [owned, *lattice]
        self._states_host = None
        self.energy_observable = ExactObservable(self, operator, calculate_variance_of_the_local_operator=True)

    # ---- states ----------------------------------------------------------------------------------------------------
    def _enumerate(self, lo, hi):
        n = hi - lo
        if self.device.type == 'cuda':
            lib = _lib.require_cuda()
            sigma = torch.empty((n, self.number_of_spins), dtype=torch.int8, device=self.device)
            _lib.check(lib.fk_exact_states(lo, n, self.number_of_spins, sigma.data_ptr(), _lib.stream_ptr()))
        else:
            index = torch.arange(lo, hi, dtype=torch.int64)
            sigma = (((index[:, None] >> torch.arange(self.number_of_spins)) & 1) * 2 - 1).to(torch.int8)
        return sigma.reshape((n,) + self.input_size)

    @property
    def states(self):
        """all 2^N states, float64 [2^N, *lattice] on the host like the reference's array (built on first use)"""
        if self._states_host is None:
            index = np.arange(self.num_of_states, dtype=np.int64)
            bits = (index[:, None] >> np.arange(self.number_of_spins, dtype=np.int64)) & 1
            self._states_host = (2.0 * bits - 1.0).reshape((self.num_of_states,) + self.input_size)
        return self._states_host

    # ---- one machine update ----------------------------------------------------------------------------------------
    def _evaluate_owned(self):
        out = torch.empty(self.slice_hi - self.slice_lo, dtype=torch.complex128, device=self.device)
        for lo in range(0, out.shape[0], self.batch_size):
            window = self._sigma[lo:lo + self.batch_size]
            if self.on_device:
                values = self.model.predict_device(window, batch_size=self.batch_size)
            else:
                values = torch.as_tensor(np.asarray(self.model.predict(window.numpy().astype(np.float64))))
            out[lo:lo + self.batch_size] = values.reshape(window.shape[0], -1)[:, 0].to(torch.complex128)
        return out

    def _gather_table(self, owned):
        dist = _process_group()
        if dist is None or self.world_size == 1:
            return owned
        where = _collective_device(owned.device)
        parts = torch.view_as_real(owned.to(where)).contiguous()
        whole = torch.empty((self.world_size * parts.shape[0], 2), dtype=parts.dtype, device=where)
        dist.all_gather_into_tensor(whole, parts)
        return torch.view_as_complex(whole.reshape(-1, 2)).to(owned.device)

    def machine_updated(self):
        self.machine_updated_start_time = time.time()
        table = self._gather_table(self._evaluate_owned())
        two_re = 2.0 * table.real
        log_norm = torch.logsumexp(two_re, dim=0)
        self._log_norm = float(log_norm)
        self.wave_function_norm_squared = float(np.exp(self._log_norm))
        log_probs = two_re - log_norm
        self._t.update(wave_function=table, log_probs=log_probs, probs=torch.exp(log_probs))
        self._host = {}
        if table.is_cuda:
            torch.cuda.synchronize()
        self.wave_function_update_end_time = time.time()
        observable = self.energy_observable
        observable.update_local_energy()
        owned = slice(self.slice_lo, self.slice_hi)
        coefficients = torch.zeros_like(observable._t['energies'])
        coefficients[owned] = observable._t['energies'][owned] - self._t['probs'][owned] * observable.current_energy
        self._t['energy_grad_coefficients'] = coefficients
        if table.is_cuda:
            torch.cuda.synchronize()
        self.local_energy_update_end_time = time.time()

    # ---- generator protocol ----------------------------------------------------------------------------------------
    def to_generator(self):
        """yields (states [batch, *lattice], coefficients [batch]) over the owned states, refreshing the tables at the
        start of every pass (exact_variational.py:148-156); host arrays, like every generator of this package"""
        while True:
            self.machine_updated()
            states, coefficients = self.states, self.energy_grad_coefficients
            for lo in range(self.slice_lo, self.slice_hi, self.batch_size):
                yield states[lo:lo + self.batch_size], coefficients[lo:lo + self.batch_size]

    def __iter__(self):
        return self.to_generator()
