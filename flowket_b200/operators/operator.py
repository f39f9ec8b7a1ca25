"""Operator base classes with the reference's protocol (flowket/operators/operator.py:6-46).  Every operator
lowers itself to a device term table (fk_operator_t); `find_conn` is the materialising drop-in (fk_find_conn),
while the local-energy kernel generates the connections on the fly from the same table."""
import abc
import ctypes

import numpy as np

from .. import _lib


class Operator(abc.ABC):
    def __init__(self, hilbert_state_shape):
        super(Operator, self).__init__()
        self.hilbert_state_shape = tuple(hilbert_state_shape)
        self.max_number_of_local_connections = None
        self._desc = None
        self._terms_dev = None

    # ---- term table ------------------------------------------------------------------------------------
    @abc.abstractmethod
    def terms(self):
        """-> (list of (site_a, site_b, kind, slot, diag_coef, off_coef), kind_id, compact, diag_fp32)"""

    def host_table(self):
        terms, kind, compact, diag_fp32 = self.terms()
        arr = (_lib.FkTerm * max(len(terms), 1))()
        for i, (a, b, k, slot, dc, oc) in enumerate(terms):
            arr[i] = _lib.FkTerm(int(a), int(b), int(k), int(slot), float(dc), float(oc))
        return arr, len(terms), kind, compact, diag_fp32

    def device_desc(self):
        """fk_operator_t whose `terms` member points to a CUDA copy of the table."""
        if self._desc is None:
            import torch
            _lib.require_cuda()
            arr, n, kind, compact, diag_fp32 = self.host_table()
            raw = np.frombuffer(arr, dtype=np.uint8, count=ctypes.sizeof(_lib.FkTerm) * max(n, 1)).copy()
            self._terms_dev = torch.from_numpy(raw).cuda()
            self._desc = _lib.FkOperator(kind, int(np.prod(self.hilbert_state_shape)),
                                         int(self.max_number_of_local_connections), n, int(compact), int(diag_fp32),
                                         ctypes.c_void_p(self._terms_dev.data_ptr()))
        return self._desc

    # ---- reference protocol ----------------------------------------------------------------------------
    mel_dtype = np.float64

    def find_conn_device(self, sample):
        """-> (conn int8 [C,B,*shape], mel float64 [C,B], use bool [C,B]) as CUDA tensors"""
        import torch
        lib = _lib.require_cuda()
        desc = self.device_desc()
        if isinstance(sample, torch.Tensor):
            s = sample.to('cuda').to(torch.int8)
        else:
            s = torch.from_numpy(np.ascontiguousarray(np.asarray(sample).astype(np.int8))).cuda()
        B = s.shape[0]
        shape = tuple(s.shape[1:])
        s = s.reshape(B, -1).contiguous()
        C, N = desc.max_conn, desc.num_sites
        assert s.shape[1] == N, 'sample has %d sites, operator %d' % (s.shape[1], N)
        conn = torch.empty((C, B, N), dtype=torch.int8, device='cuda')
        mel = torch.empty((C, B), dtype=torch.float64, device='cuda')
        use = torch.empty((C, B), dtype=torch.uint8, device='cuda')
        _lib.check(lib.fk_find_conn(ctypes.byref(desc), ctypes.c_void_p(s.data_ptr()), B,
                                    ctypes.c_void_p(conn.data_ptr()), ctypes.c_void_p(mel.data_ptr()),
                                    ctypes.c_void_p(use.data_ptr()), _lib.stream_ptr()))
        return conn.reshape((C, B) + shape), mel, use.bool()

    def find_conn(self, sample):
        """Reference layout and dtypes: all_conn[C,B,*shape] float64, mel[C,B], use_conn[C,B] bool (host)."""
        conn, mel, use = self.find_conn_device(sample)
        return (conn.cpu().numpy().astype(np.float64), mel.cpu().numpy().astype(self.mel_dtype),
                use.cpu().numpy())

    def use_state(self, state):
        return True

    def random_states(self, num_of_states):
        return np.random.choice([-1, 1], size=(num_of_states,) + self.hilbert_state_shape)


def cube_shape(number_of_spins_in_each_dimention=20, cube_dimention=1, column_or_row=True):
    if cube_dimention == 1:
        return [number_of_spins_in_each_dimention, 1] if column_or_row else [1, number_of_spins_in_each_dimention]
    return [number_of_spins_in_each_dimention, ] * cube_dimention


class OperatorOnGrid(Operator, abc.ABC):
    def __init__(self, hilbert_state_shape=None, pbc=True):
        if hilbert_state_shape is None:
            hilbert_state_shape = cube_shape()
        super(OperatorOnGrid, self).__init__(hilbert_state_shape)
        self.pbc = pbc
