"""GPU probe: how many conjugate-gradient iterations (reference defaults: tol 1e-3 relative to |r0|, <= 200) does the
sample-space SR system of the headline machine need?  Builds S with the device pipeline, then runs CG with torch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import torch
from flowket_b200 import Input, Model, _lib, FK_ENGINE_TC
from flowket_b200.machines import ConvNetAutoregressive2D
from flowket_b200.operators import Heisenberg
from flowket_b200.observables.monte_carlo import Observable

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
lib = _lib.require_cuda()
inp = Input(shape=(10, 10), dtype='int8')
m = ConvNetAutoregressive2D(inp, depth=20, num_of_channels=32, seed=0)
model = Model(inputs=inp, outputs=m.predictions)
model.engine = FK_ENGINE_TC
net = m.device_net()
sg = net.sample(B, seed=3, engine=FK_ENGINE_TC)
obs = Observable(Heisenberg(hilbert_state_shape=[10, 10], pbc=False))
e = obs.local_values_device(model, sg)
P = net.num_params
nkb = (P + 63) // 64
R = 2 * B
X = torch.empty((nkb, R, 64), dtype=torch.bfloat16, device=sg.device)
wsb = lib.fk_jacobian_rows_tc_workspace_bytes(net.handle, B)
ws = torch.empty(wsb, dtype=torch.uint8, device=sg.device)
_lib.check(lib.fk_jacobian_rows_tc(net.handle, sg.data_ptr(), B, X.data_ptr(), R, 0, B, ws.data_ptr(), wsb, _lib.stream_ptr()))
G = torch.empty((R, R), dtype=torch.float32, device=sg.device)
wsb = lib.fk_sr_gram_xxt_workspace_bytes(R)
ws = torch.empty(wsb, dtype=torch.uint8, device=sg.device)
_lib.check(lib.fk_sr_gram_xxt(X.data_ptr(), R, P, R, 1, 0, 1.0, G.data_ptr(), R, ws.data_ptr(), wsb, _lib.stream_ptr()))
S = torch.empty((R, R), dtype=torch.float64, device=sg.device)
wsb = lib.fk_sr_centre_shift_workspace_bytes(R)
ws = torch.empty(wsb, dtype=torch.uint8, device=sg.device)
_lib.check(lib.fk_sr_centre_shift(G.data_ptr(), R, R, 1, 0.05, S.data_ptr(), ws.data_ptr(), wsb, _lib.stream_ptr()))
e = e - e.mean()
rhs = torch.cat([e.real, e.imag]) / B
print('diag of S: min %.3g max %.3g mean %.3g' % (S.diagonal().min().item(), S.diagonal().max().item(), S.diagonal().mean().item()))
for dtype in (torch.float64, torch.float32):
    Sd = S.to(dtype)
    b = rhs.to(dtype)
    x = torch.zeros_like(b); r = b.clone(); p = r.clone(); gamma = torch.dot(r, r)
    r0 = r.norm().item()
    it = 0
    hist = []
    while it < 400:
        z = Sd @ p
        alpha = gamma / torch.dot(p, z)
        x += alpha * p; r -= alpha * z
        g2 = torch.dot(r, r)
        p = r + (g2 / gamma) * p
        gamma = g2
        it += 1
        rn = r.norm().item() / r0
        if it in (10, 20, 50, 100, 200, 300, 400) or rn < 1e-3:
            hist.append((it, rn))
        if rn < 1e-6:
            break
    first = next((i for i, v in hist if v < 1e-3), None)
    print(dtype, 'iterations to 1e-3:', first, 'history', [(i, '%.2e' % v) for i, v in hist][:12])
xd = torch.linalg.solve(S, rhs)
print('CG(1e-6) vs direct:', ((x.double() - xd).norm() / xd.norm()).item())
ev = torch.linalg.eigvalsh(S[:4096, :4096])
print('eig range of a 4096 principal block: %.3g .. %.3g' % (ev.min().item(), ev.max().item()))


# ---- preconditioned CG: block-Jacobi with the diagonal blocks a rank would own in the sharded step (fp32 Cholesky of each block)
def pcg(Sd, b, Minv, tol, max_iter=2000):
    x = torch.zeros_like(b); r = b.clone(); z = Minv(r); p = z.clone(); rz = torch.dot(r, z)
    r0 = r.norm().item()
    for it in range(1, max_iter + 1):
        q = Sd @ p
        alpha = rz / torch.dot(p, q)
        x += alpha * p; r -= alpha * q
        if r.norm().item() <= tol * r0:
            return x, it
        z = Minv(r); rz2 = torch.dot(r, z); p = z + (rz2 / rz) * p; rz = rz2
    return x, max_iter


b64 = rhs.double()
for nblk in (2, 4, 8, 16, 64):
    bs = R // nblk
    Ls = [torch.linalg.cholesky(S[i * bs:(i + 1) * bs, i * bs:(i + 1) * bs].float()).double() for i in range(nblk)]

    def Minv(r, Ls=Ls, bs=bs):
        out = torch.empty_like(r)
        for i, L in enumerate(Ls):
            out[i * bs:(i + 1) * bs] = torch.cholesky_solve(r[i * bs:(i + 1) * bs].reshape(-1, 1), L).reshape(-1)
        return out
    res = []
    for tol in (1e-3, 1e-6, 1e-9):
        x, it = pcg(S, b64, Minv, tol)
        res.append((tol, it, ((x - xd).norm() / xd.norm()).item()))
    print('block-Jacobi PCG, %d blocks of %d: ' % (nblk, bs) + ', '.join('tol %.0e: %d its (err %.1e)' % t for t in res), flush=True)
x, it = pcg(S, b64, lambda r: r / S.diagonal(), 1e-6)
print('Jacobi PCG: %d its to 1e-6' % it)
