"""GPU tool: fk_sr_solve (fp64 potrf) vs fk_sr_solve_mixed (fp32 potrf + fp64 refinement) at the headline size; stage costs."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flowket_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
lib = _lib.require_cuda()
dev = torch.device('cuda')
torch.manual_seed(0)
X = torch.randn(n, 4096, device=dev) * torch.logspace(0, -3, 4096, device=dev)
S = ((X @ X.T) / (n // 2)).double()
S.diagonal().add_(0.05)
S = (0.5 * (S + S.T)).contiguous()
b = torch.randn(n, device=dev, dtype=torch.float64)
h = ctypes.c_void_p()
_lib.check(lib.fk_sr_solver_create(ctypes.byref(h)))


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


ws64 = torch.empty(lib.fk_sr_solve_workspace_bytes(h, n), dtype=torch.uint8, device=dev)
wsm = torch.empty(lib.fk_sr_solve_mixed_workspace_bytes(h, n), dtype=torch.uint8, device=dev)
info = torch.zeros(1, dtype=torch.int32, device=dev)
Sc = torch.empty_like(S)


def solve64():
    Sc.copy_(S)
    x = b.clone()
    _lib.check(lib.fk_sr_solve(h, Sc.data_ptr(), x.data_ptr(), n, info.data_ptr(), ws64.data_ptr(), ws64.numel(), _lib.stream_ptr()))
    return x


def mixed(k):
    def f():
        x = b.clone()
        resid = torch.zeros(k + 2, dtype=torch.float64, device=dev)
        _lib.check(lib.fk_sr_solve_mixed(h, S.data_ptr(), x.data_ptr(), n, k, info.data_ptr(), resid.data_ptr(), wsm.data_ptr(),
                                         wsm.numel(), _lib.stream_ptr()))
        return x, resid
    return f


t_copy, _ = timed(lambda: Sc.copy_(S))
t64, x64 = timed(solve64)
print('n %d: fp64 potrf + potrs %.2f ms (of which the 8 n^2-byte copy %.2f ms)' % (n, t64, t_copy))
for k in (0, 1, 2, 3, 4):
    t, (x, resid) = timed(mixed(k))
    hist = torch.sqrt(resid / resid[0]).cpu().numpy()
    print('mixed, %d refinements: %.2f ms, |x - x64| / |x64| %.2e, residual history %s' % (
        k, t, float((x - x64).norm() / x64.norm()), ['%.1e' % v for v in hist]))
S32 = S.float()
t, _ = timed(lambda: torch.linalg.cholesky_ex(S32))
print('torch.linalg.cholesky_ex fp32: %.2f ms' % t)
t, _ = timed(lambda: torch.linalg.cholesky_ex(S))
print('torch.linalg.cholesky_ex fp64: %.2f ms' % t)
t, _ = timed(lambda: S.float())
print('fp64 -> fp32 conversion by torch: %.2f ms' % t)
L32 = torch.linalg.cholesky_ex(S32)[0]
rhs32 = b.float().reshape(-1, 1)
t, _ = timed(lambda: torch.cholesky_solve(rhs32, L32))
print('torch.cholesky_solve fp32, one right-hand side: %.2f ms' % t)
t, _ = timed(lambda: S @ b)
print('fp64 matvec (torch): %.2f ms' % t)
