"""ctypes binding of libflowket_b200.so (the C ABI declared in include/flowket_b200.h).

There is no CPU fallback: importing the package works without the library (so that host-only logic can be
tested on a CPU box), but the first call into the device path raises if the CUDA extension is missing.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libflowket_b200.so')

c_void_p, c_int, c_int64, c_uint64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64

FK_NET_CONV2D, FK_NET_CONV1D, FK_NET_CCONV1D = 0, 1, 2
FK_FLAG_WEIGHT_NORM, FK_FLAG_EXP_NORM, FK_FLAG_SKIP = 1, 2, 4
FK_OP_HEISENBERG, FK_OP_ISING, FK_OP_J1J2 = 0, 1, 2
FK_TERM_EXCHANGE, FK_TERM_FLIP, FK_TERM_DIAG = 0, 1, 2
FK_ENGINE_FP32, FK_ENGINE_TC, FK_ENGINE_TC_EXACT = 0, 1, 2


class FkTerm(ctypes.Structure):
    _fields_ = [('site_a', ctypes.c_int32), ('site_b', ctypes.c_int32), ('kind', ctypes.c_int32),
                ('slot', ctypes.c_int32), ('diag_coef', ctypes.c_double), ('off_coef', ctypes.c_double)]


class FkOperator(ctypes.Structure):
    _fields_ = [('kind', ctypes.c_int32), ('num_sites', ctypes.c_int32), ('max_conn', ctypes.c_int32),
                ('num_terms', ctypes.c_int32), ('compact', ctypes.c_int32), ('diag_fp32', ctypes.c_int32),
                ('terms', c_void_p)]


# name -> (restype, argtypes); every symbol include/flowket_b200.h declares
SIGNATURES = {
    'fk_last_error': (ctypes.c_char_p, []),
    'fk_version': (c_int, []),
    'fk_launch_count': (c_int64, []),
    'fk_net_create': (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    'fk_net_destroy': (c_int, [c_void_p]),
    'fk_net_num_params': (c_int, [c_void_p, ctypes.POINTER(c_int64)]),
    'fk_net_set_params': (c_int, [c_void_p, c_void_p, c_void_p]),
    'fk_log_psi_workspace_bytes': (c_int64, [c_void_p, c_int64, c_int]),
    'fk_log_psi': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_void_p, c_int64, c_void_p]),
    'fk_cond_log_probs': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p]),
    'fk_sample_workspace_bytes': (c_int64, [c_void_p, c_int64]),
    'fk_sample': (c_int, [c_void_p, c_void_p, c_uint64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int64,
                          c_void_p]),
    'fk_sample_tc_workspace_bytes': (c_int64, [c_void_p, c_int64]),
    'fk_sample_tc': (c_int, [c_void_p, c_void_p, c_uint64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int64,
                             c_void_p]),
    'fk_sample_naive_workspace_bytes': (c_int64, [c_void_p, c_int64]),
    'fk_sample_naive': (c_int, [c_void_p, c_void_p, c_uint64, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                c_int64, c_void_p]),
    'fk_find_conn': (c_int, [ctypes.POINTER(FkOperator), c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    'fk_local_energy_workspace_bytes': (c_int64, [c_void_p, ctypes.POINTER(FkOperator), c_int64, c_int]),
    'fk_local_energy': (c_int, [c_void_p, ctypes.POINTER(FkOperator), c_void_p, c_int64, c_void_p, c_void_p,
                                ctypes.POINTER(c_int64), c_int, c_void_p, c_int64, c_void_p]),
    'fk_grad_workspace_bytes': (c_int64, [c_void_p, c_int64, c_int]),
    'fk_grad_weighted': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p]),
    'fk_grad_weighted_tc_workspace_bytes': (c_int64, [c_void_p, c_int64]),
    'fk_grad_weighted_tc': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p]),
    'fk_grad_per_sample': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    'fk_grad_per_sample_tc_workspace_bytes': (c_int64, [c_void_p, c_int64]),
    'fk_grad_per_sample_tc': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    'fk_sr_gram_workspace_bytes': (c_int64, [c_int64, c_int64, c_int]),
    'fk_sr_gram': (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_int64, c_void_p]),
    'fk_sr_gram_tc_workspace_bytes': (c_int64, [c_int64, c_int64, c_int, c_int]),
    'fk_sr_gram_tc': (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p, c_int64, c_void_p]),
    'fk_jacobian_rows_tc_workspace_bytes': (c_int64, [c_void_p, c_int64]),
    'fk_jacobian_rows_tc': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64,
                                    c_void_p]),
    'fk_sr_gram_xxt_workspace_bytes': (c_int64, [c_int64]),
    'fk_sr_gram_xxt': (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, ctypes.c_float, c_void_p, c_int64,
                               c_void_p, c_int64, c_void_p]),
    'fk_sr_centre_shift_workspace_bytes': (c_int64, [c_int64]),
    'fk_sr_centre_shift': (c_int, [c_void_p, c_int64, c_int64, c_int64, ctypes.c_double, c_void_p, c_void_p, c_int64,
                                   c_void_p]),
    'fk_sr_xt_w': (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    'fk_sr_solver_create': (c_int, [ctypes.POINTER(c_void_p)]),
    'fk_sr_solver_destroy': (c_int, [c_void_p]),
    'fk_sr_solve_workspace_bytes': (c_int64, [c_void_p, c_int64]),
    'fk_sr_solve': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p]),
    'fk_sr_solve_mixed_workspace_bytes': (c_int64, [c_void_p, c_int64]),
    'fk_sr_solve_mixed': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_int64,
                                  c_void_p]),
    'fk_sr_factor_mixed': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p]),
    'fk_sr_solve_factored': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int64, c_void_p]),
    'fk_exact_states': (c_int, [c_int64, c_int64, c_int, c_void_p, c_void_p]),
    'fk_exact_index': (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    'fk_exact_conn_table': (c_int, [ctypes.POINTER(FkOperator), c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    'fk_exact_local_energy': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, ctypes.c_double, c_void_p, c_void_p,
                                      c_void_p]),
}


class FlowketB200Error(RuntimeError):
    pass


_lib = None


def load(path=None):
    """Load the shared library (no CUDA call is made). Raises if it is missing -- there is no CPU fallback."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise FlowketB200Error(
            'flowket_b200: CUDA extension %s not found. Build it with `python -c "import __graft_entry__ as g; '
            'g.build()"` (or `make -C flowket_b200/csrc`). There is no CPU fallback.' % path)
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise FlowketB200Error(load().fk_last_error().decode('utf-8', 'replace'))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise FlowketB200Error('flowket_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.')
    return load()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
