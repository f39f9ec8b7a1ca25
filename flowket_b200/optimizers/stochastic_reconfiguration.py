"""Stochastic reconfiguration: algebra of
flowket/optimizers/stochastic_reconfiguration/optimizer.py:33-124 and linear_equations.py:34-137.

  Obar = O - mean_b O;  F = Obar^H conj(y_true), y_true = conj(E_loc - E)/B;  S = Obar^H Obar / B + lambda I;
  delta = S^-1 F (Cholesky / solve, or CG with tol relative to |r0|);  W <- W - lr * delta.

Per-sample Jacobians come from the device (fk_grad_per_sample); the P x P contraction runs through fk_sr_gram
on the stacked real matrix [Re Obar ; Im Obar].  ComplexValuesStochasticReconfiguration assembles complex
parameters pairwise from (real, imag) variables as the reference does (complex_values_optimizer.py:8-9,48-58,72-76).
StochasticReconfiguration is the real-parameter variant the reference lacks:
S = Re(Obar^H Obar)/B + lambda I, F = Re(Obar^H (E_loc - E))/B."""
import numpy as np


def conjugate_gradient(apply, rhs, tol=1e-3, max_iter=200):
    """torch port of the TF contrib solver the reference vendors: stops when |r| <= tol * |r0| or at max_iter."""
    import torch
    x = torch.zeros_like(rhs)
    r = rhs.clone()
    p = r.clone()
    gamma = torch.vdot(r, r)
    tol_abs = tol * float(torch.linalg.vector_norm(r))
    i = 0
    while (max_iter is None or i < max_iter) and float(torch.linalg.vector_norm(r)) > tol_abs:
        z = apply(p)
        alpha = gamma / torch.vdot(p, z)
        x = x + alpha * p
        r = r - alpha * z
        gamma_new = torch.vdot(r, r)
        p = r + (gamma_new / gamma) * p
        gamma = gamma_new
        i += 1
    return x, i, float(torch.linalg.vector_norm(r))


class _SRBase(object):
    def __init__(self, model, lr=0.01, diag_shift=0.05, iterative_solver=True, conjugate_gradient_tol=1e-3,
                 iterative_solver_max_iterations=200, use_cholesky=True):
        self.model, self.machine = model, model.machine
        self.lr, self.diag_shift = lr, diag_shift
        self.iterative_solver = iterative_solver
        self.conjugate_gradient_tol = conjugate_gradient_tol
        self.iterative_solver_max_iterations = iterative_solver_max_iterations
        self.use_cholesky = use_cholesky
        self.conjugate_gradient_iterations = 0
        self.conjugate_gradient_residual_norm = 0.0

    def jacobian(self, sigma):
        net = self.machine.device_net()
        return net.grad_per_sample(net.to_sigma(sigma), imag=True)

    def solve(self, O_bar, rhs):
        """(Obar^H Obar / B + lambda I)^-1 rhs for a complex (or real) centred Jacobian on the device."""
        import torch
        from .._device import sr_gram
        B = O_bar.shape[0]
        if self.iterative_solver:
            def apply(v):
                return O_bar.conj().T @ (O_bar @ v) / B + self.diag_shift * v
            x, it, res = conjugate_gradient(apply, rhs, self.conjugate_gradient_tol, self.iterative_solver_max_iterations)
            self.conjugate_gradient_iterations, self.conjugate_gradient_residual_norm = it, res
            return x
        if O_bar.is_complex():
            # Obar^H Obar = (R^T R + I^T I) + i (R^T I - I^T R): one real Gram of the stacked matrix [R | I]
            R, I = O_bar.real.contiguous().float(), O_bar.imag.contiguous().float()
            G = sr_gram(torch.cat([R, I], dim=1), transpose_a=True)      # [2P, 2P]
            P = R.shape[1]
            S = torch.complex(G[:P, :P] + G[P:, P:], G[:P, P:] - G[P:, :P]) / B
        else:
            S = sr_gram(O_bar.float(), transpose_a=True) / B
        S = S + self.diag_shift * torch.eye(S.shape[0], dtype=S.dtype, device=S.device)
        if self.use_cholesky:
            L = torch.linalg.cholesky(S)
            return torch.cholesky_solve(rhs.reshape(-1, 1).to(S.dtype), L).reshape(-1)
        return torch.linalg.solve(S, rhs.to(S.dtype))


class ComplexValuesStochasticReconfiguration(_SRBase):
    """For machines whose parameters are complex (real, imag) pairs with W = real - i*imag
    (ComplexValuesSimpleConvNetAutoregressive1D)."""

    def __init__(self, predictions_keras_model, predictions_jacobian=None, **kwargs):
        super(ComplexValuesStochasticReconfiguration, self).__init__(predictions_keras_model, **kwargs)
        specs = self.machine.weight_specs()
        self._pairs = []   # (offset_real, offset_imag, size) per complex weight
        off = 0
        offsets = []
        for _, shape, _ in specs:
            offsets.append((off, int(np.prod(shape))))
            off += int(np.prod(shape))
        assert len(specs) % 2 == 0
        for i in range(0, len(specs), 2):
            assert specs[i][0].replace('_real', '') == specs[i + 1][0].replace('_imag', ''), 'not a complex machine'
            self._pairs.append((offsets[i][0], offsets[i + 1][0], offsets[i][1]))

    def _complex_index(self, device):
        import torch
        re = torch.cat([torch.arange(o, o + n) for o, _, n in self._pairs]).to(device)
        im = torch.cat([torch.arange(o, o + n) for _, o, n in self._pairs]).to(device)
        return re, im

    def complex_jacobian(self, sigma):
        """complex(dRe f/da, dRe f/db) per (a, b) pair == d log psi / dW for holomorphic log psi
        (complex_values_optimizer.py:8-9,72-76)."""
        import torch
        O_re, _ = self.jacobian(sigma)
        re, im = self._complex_index(O_re.device)
        return torch.complex(O_re[:, re], O_re[:, im])

    def compute_update(self, sigma, y_true):
        """delta (complex [P_c]) for a batch; y_true = conj(E_loc - E)/B as produced by VariationalMonteCarlo."""
        import torch
        O = self.complex_jacobian(sigma)
        O_bar = O - O.mean(dim=0, keepdim=True)
        y = torch.as_tensor(np.asarray(y_true, np.complex64)).to(O.device)
        F = O_bar.conj().T @ torch.conj(y)
        return self.solve(O_bar, F)

    def apply_complex_gradient(self, delta):
        """a += lr Re(conj(-delta)), b += lr Im(conj(-delta))  =>  W <- W - lr delta (complex_values_optimizer.py:48-58)"""
        import torch
        params = self.machine.flat_params_device()
        re, im = self._complex_index(params.device)
        conj_g = torch.conj(-delta)
        params[re] += self.lr * conj_g.real.float()
        params[im] += self.lr * conj_g.imag.float()
        self.machine.params_updated()

    def step(self, sigma, y_true):
        delta = self.compute_update(sigma, y_true)
        self.apply_complex_gradient(delta)
        return delta


class StochasticReconfiguration(_SRBase):
    """Real-parameter SR (ConvNetAutoregressive2D, SimpleConvNetAutoregressive1D)."""

    def compute_update(self, sigma, local_energy):
        import torch
        O_re, O_im = self.jacobian(sigma)
        B = O_re.shape[0]
        e = torch.as_tensor(np.asarray(local_energy, np.complex128)).to(O_re.device)
        e = e - e.mean()
        R = (O_re - O_re.mean(dim=0, keepdim=True))
        I = (O_im - O_im.mean(dim=0, keepdim=True))
        F = (R.T @ e.real.float() + I.T @ e.imag.float()) / B          # Re(Obar^H (E - Ebar)) / B
        stacked = torch.cat([R, I], dim=0)                             # Re(Obar^H Obar) = R^T R + I^T I
        return self.solve(stacked, F) if not self.iterative_solver else self._solve_real(stacked, F, B)

    def _solve_real(self, stacked, F, B):
        def apply(v):
            return stacked.T @ (stacked @ v) / B + self.diag_shift * v
        x, it, res = conjugate_gradient(apply, F, self.conjugate_gradient_tol, self.iterative_solver_max_iterations)
        self.conjugate_gradient_iterations, self.conjugate_gradient_residual_norm = it, res
        return x

    def solve(self, stacked, rhs):
        import torch
        from .._device import sr_gram
        B = stacked.shape[0] // 2
        S = sr_gram(stacked.float(), transpose_a=True) / B
        S = S + self.diag_shift * torch.eye(S.shape[0], dtype=S.dtype, device=S.device)
        L = torch.linalg.cholesky(S)
        return torch.cholesky_solve(rhs.reshape(-1, 1), L).reshape(-1)

    def step(self, sigma, local_energy):
        delta = self.compute_update(sigma, local_energy)
        params = self.machine.flat_params_device()
        params.add_(delta.float(), alpha=-self.lr)
        self.machine.params_updated()
        return delta
