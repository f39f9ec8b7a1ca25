"""Device-side machine wrapper: owns the fk_net handle, the flat fp32 parameter vector (torch CUDA tensor)
and grow-only workspaces.  torch is used for device memory and streams only; all arithmetic on the hot path
runs in libflowket_b200.so."""
import ctypes

import numpy as np

from . import _lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class DeviceNet(object):
    def __init__(self, kind, H, W, depth, channels, kernel_size, max_dilation, flags, device=None):
        import torch
        self.lib = _lib.require_cuda()
        self.torch = torch
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.H, self.W, self.sites = H, W, H * W
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.fk_net_create(ctypes.byref(handle), kind, H, W, depth, channels, kernel_size,
                                              max_dilation if max_dilation else 0, flags))
        self.handle = handle
        n = ctypes.c_int64()
        _lib.check(self.lib.fk_net_num_params(self.handle, ctypes.byref(n)))
        self.num_params = n.value
        self._ws = {}

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                self.lib.fk_net_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---- helpers --------------------------------------------------------------------------------------
    def workspace(self, name, nbytes):
        t = self._ws.get(name)
        if t is None or t.numel() < nbytes:
            self._ws[name] = None
            t = self.torch.empty(int(nbytes), dtype=self.torch.uint8, device=self.device)
            self._ws[name] = t
        return t

    def to_sigma(self, x):
        """any array-like of +-1 (or 0 for unsampled sites) -> contiguous int8 CUDA tensor [n, sites]"""
        torch = self.torch
        if isinstance(x, torch.Tensor):
            t = x.to(device=self.device)
        else:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(x).astype(np.int8, copy=False))).to(self.device)
        return t.to(torch.int8).reshape(t.shape[0], self.sites).contiguous()

    # ---- entry points -----------------------------------------------------------------------------------
    def set_params(self, flat):
        assert flat.is_cuda and flat.dtype == self.torch.float32 and flat.numel() == self.num_params
        with self.torch.cuda.device(self.device):
            _lib.check(self.lib.fk_net_set_params(self.handle, _ptr(flat.contiguous()), _lib.stream_ptr()))

    def log_psi(self, sigma, engine=_lib.FK_ENGINE_FP32, max_chunk=None):
        torch = self.torch
        n = sigma.shape[0]
        out = torch.empty((n, 2), dtype=torch.float32, device=self.device)
        if n == 0:
            return torch.view_as_complex(out)
        chunk = n if max_chunk is None else min(n, max_chunk)
        if engine == _lib.FK_ENGINE_FP32:
            chunk = min(chunk, 16384)
        nbytes = self.lib.fk_log_psi_workspace_bytes(self.handle, chunk, engine)
        ws = self.workspace('fwd', nbytes)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.fk_log_psi(self.handle, _ptr(sigma), n, _ptr(out), engine, _ptr(ws), ws.numel(),
                                           _lib.stream_ptr()))
        return torch.view_as_complex(out)

    def cond_log_probs(self, sigma, max_chunk=16384):
        torch = self.torch
        n = sigma.shape[0]
        out = torch.empty((n, self.sites, 2), dtype=torch.float32, device=self.device)
        if n == 0:
            return out
        nbytes = self.lib.fk_log_psi_workspace_bytes(self.handle, min(n, max_chunk), _lib.FK_ENGINE_FP32)
        ws = self.workspace('fwd', nbytes)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.fk_cond_log_probs(self.handle, _ptr(sigma), n, _ptr(out), _ptr(ws), ws.numel(),
                                                  _lib.stream_ptr()))
        return out

    def sample(self, batch_size, uniforms=None, seed=0, sample_offset=0, naive=False, return_p0=False,
               engine=_lib.FK_ENGINE_FP32):
        torch = self.torch
        sigma = torch.empty((batch_size, self.sites), dtype=torch.int8, device=self.device)
        p0 = torch.empty((batch_size, self.sites), dtype=torch.float32, device=self.device) if return_p0 else None
        if uniforms is not None:
            uniforms = uniforms.to(device=self.device, dtype=torch.float64).reshape(batch_size, self.sites).contiguous()
        if naive:
            size_fn, fn = self.lib.fk_sample_naive_workspace_bytes, self.lib.fk_sample_naive
        elif engine in (_lib.FK_ENGINE_TC, _lib.FK_ENGINE_TC_EXACT):
            size_fn, fn = self.lib.fk_sample_tc_workspace_bytes, self.lib.fk_sample_tc
        else:
            size_fn, fn = self.lib.fk_sample_workspace_bytes, self.lib.fk_sample
        nbytes = size_fn(self.handle, batch_size)
        if nbytes < 0:
            raise _lib.FlowketB200Error('the tensor-core sampler supports ConvNetAutoregressive2D with 32 channels only')
        ws = self.workspace('sample', nbytes)
        with torch.cuda.device(self.device):
            _lib.check(fn(self.handle, _ptr(uniforms), seed, sample_offset, batch_size, _ptr(sigma), _ptr(p0), _ptr(ws),
                          ws.numel(), _lib.stream_ptr()))
        return (sigma, p0) if return_p0 else sigma

    def local_energy(self, op_desc, sigma, engine=_lib.FK_ENGINE_FP32, count=True):
        """-> (E_loc complex128 [B], stats float64 [4] = (sum Re, sum Im, sum Re^2, count), n_conn).
        count=False: n_conn is None and, on the tensor-core engines, the call never waits for the device."""
        torch = self.torch
        B = sigma.shape[0]
        eloc = torch.empty((B, 2), dtype=torch.float64, device=self.device)
        stats = torch.zeros(4, dtype=torch.float64, device=self.device)
        n_conn = ctypes.c_int64(0) if count else None
        if B:
            nbytes = self.lib.fk_local_energy_workspace_bytes(self.handle, ctypes.byref(op_desc), B, engine)
            ws = self.workspace('eloc', nbytes)
            with torch.cuda.device(self.device):
                _lib.check(self.lib.fk_local_energy(self.handle, ctypes.byref(op_desc), _ptr(sigma), B, _ptr(eloc),
                                                    _ptr(stats), ctypes.byref(n_conn) if count else None, engine, _ptr(ws),
                                                    ws.numel(), _lib.stream_ptr()))
        return torch.view_as_complex(eloc), stats, (n_conn.value if count else None)

    def grad_weighted(self, sigma, y, engine=_lib.FK_ENGINE_FP32):
        """sum_b 2 Re(log psi_b y_b) differentiated w.r.t. the flat parameter vector; y complex64 [B]"""
        torch = self.torch
        B = sigma.shape[0]
        y2 = torch.view_as_real(y.to(device=self.device, dtype=torch.complex64).contiguous()).contiguous()
        grad = torch.empty(self.num_params, dtype=torch.float32, device=self.device)
        if engine in (_lib.FK_ENGINE_TC, _lib.FK_ENGINE_TC_EXACT):
            nbytes = self.lib.fk_grad_weighted_tc_workspace_bytes(self.handle, B)
            if nbytes < 0:
                raise _lib.FlowketB200Error('the tensor-core gradient supports ConvNetAutoregressive2D (32 channels, '
                                            'kernel 3) on lattices up to 10x10')
            ws = self.workspace('grad_tc', nbytes)
            fn = self.lib.fk_grad_weighted_tc
        else:
            ws = self.workspace('grad', self.lib.fk_grad_workspace_bytes(self.handle, B, 0))
            fn = self.lib.fk_grad_weighted
        with torch.cuda.device(self.device):
            _lib.check(fn(self.handle, _ptr(sigma), _ptr(y2), B, _ptr(grad), _ptr(ws), ws.numel(), _lib.stream_ptr()))
        return grad

    def grad_per_sample(self, sigma, imag=True, engine=_lib.FK_ENGINE_FP32):
        torch = self.torch
        B = sigma.shape[0]
        O_re = torch.empty((B, self.num_params), dtype=torch.float32, device=self.device)
        O_im = torch.empty((B, self.num_params), dtype=torch.float32, device=self.device) if imag else None
        if engine in (_lib.FK_ENGINE_TC, _lib.FK_ENGINE_TC_EXACT):
            nbytes = self.lib.fk_grad_per_sample_tc_workspace_bytes(self.handle, B)
            if nbytes < 0:
                raise _lib.FlowketB200Error('tensor-core per-sample gradient: machine / lattice not supported')
            ws = self.workspace('grad_ps_tc', nbytes)
            fn = self.lib.fk_grad_per_sample_tc
        else:
            ws = self.workspace('grad_ps', self.lib.fk_grad_workspace_bytes(self.handle, B, 1))
            fn = self.lib.fk_grad_per_sample
        with torch.cuda.device(self.device):
            _lib.check(fn(self.handle, _ptr(sigma), B, _ptr(O_re), _ptr(O_im), _ptr(ws), ws.numel(), _lib.stream_ptr()))
        return O_re, O_im


def sr_gram(A, transpose_a, engine=_lib.FK_ENGINE_FP32, precise=True):
    """G = A^T A (transpose_a) or A A^T on the device: fp32 CUDA-core kernel (fk_sr_gram) or the tcgen05 GEMM
    (fk_sr_gram_tc; fp16 hi/lo operands when `precise`, fp32 accumulation)."""
    import torch
    lib = _lib.require_cuda()
    A = A.contiguous()
    rows, cols = A.shape
    M = cols if transpose_a else rows
    G = torch.empty((M, M), dtype=torch.float32, device=A.device)
    with torch.cuda.device(A.device):
        if engine in (_lib.FK_ENGINE_TC, _lib.FK_ENGINE_TC_EXACT):
            nbytes = lib.fk_sr_gram_tc_workspace_bytes(rows, cols, int(transpose_a), int(bool(precise)))
            ws = torch.empty(int(nbytes), dtype=torch.uint8, device=A.device)
            _lib.check(lib.fk_sr_gram_tc(_ptr(A), rows, cols, int(transpose_a), int(bool(precise)), _ptr(G), _ptr(ws),
                                         ws.numel(), _lib.stream_ptr()))
        else:
            nbytes = lib.fk_sr_gram_workspace_bytes(rows, cols, int(transpose_a))
            ws = torch.empty(int(nbytes), dtype=torch.uint8, device=A.device)
            _lib.check(lib.fk_sr_gram(_ptr(A), rows, cols, int(transpose_a), _ptr(G), _ptr(ws), ws.numel(),
                                      _lib.stream_ptr()))
    return G
