"""Probe: fp64 vs fp32 Cholesky (+ iterative refinement) of the 2B x 2B sample-space SR system on one GPU."""
import sys
import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
torch.manual_seed(0)
X = torch.randn(n, 4096, device='cuda') * torch.logspace(0, -3, 4096, device='cuda')      # decaying spectrum
T = (X @ X.T) / (n // 2)
T.diagonal().add_(0.05)
b = torch.randn(n, device='cuda', dtype=torch.float64)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


def solve64():
    L = torch.linalg.cholesky(T.double())
    return torch.cholesky_solve(b.reshape(-1, 1), L).reshape(-1)


def solve32(refine=2):
    L, info = torch.linalg.cholesky_ex(T)
    w = torch.cholesky_solve(b.float().reshape(-1, 1), L).reshape(-1).double()
    for _ in range(refine):
        r = b - (T @ w.float()).double()            # residual in fp32 (the fp64-residual variant is below)
        w = w + torch.cholesky_solve(r.float().reshape(-1, 1), L).reshape(-1).double()
    return w


t64, w64 = timed(solve64)
print('n = %d  cond ~ %.2e' % (n, float(torch.linalg.matrix_norm(T, 2) / 0.05)))
print('fp64 cholesky + solve: %.2f ms' % t64)
for refine in (0, 1, 2, 3):
    t32, w32 = timed(lambda: solve32(refine))
    print('fp32 cholesky + %d refinement(s): %.2f ms, rel err vs fp64 %.2e' % (
        refine, t32, float((w32 - w64).norm() / w64.norm())))
# refinement with the residual in fp64 (T promoted block-wise)
def solve32_r64(refine=2, rows=4096):
    L, info = torch.linalg.cholesky_ex(T)
    w = torch.cholesky_solve(b.float().reshape(-1, 1), L).reshape(-1).double()
    for _ in range(refine):
        r = b.clone()
        for i in range(0, n, rows):
            r[i:i + rows] -= T[i:i + rows].double() @ w
        w = w + torch.cholesky_solve(r.float().reshape(-1, 1), L).reshape(-1).double()
    return w
for refine in (1, 2):
    t32, w32 = timed(lambda: solve32_r64(refine))
    print('fp32 cholesky + %d fp64-residual refinement(s): %.2f ms, rel err vs fp64 %.2e' % (
        refine, t32, float((w32 - w64).norm() / w64.norm())))
