"""CPU restatement (numpy) of the reference's stochastic-reconfiguration algebra.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Restates (relative to /root/reference/src/flowket):
  * get_wave_function_jacobian_minus_mean, get_energy_grad, direct / iterative solve
        optimizers/stochastic_reconfiguration/optimizer.py:33-124
  * conjugate_gradient (tol relative to |r0|, max_iter)
        optimizers/stochastic_reconfiguration/linear_equations.py:34-137
  * complex Jacobian assembly complex(dRe f/da, dRe f/db) and W <- W - lr*delta
        optimizers/complex_values_optimizer.py:8-9,48-58,72-76
Real-parameter SR (ConvNetAutoregressive2D / SimpleConvNetAutoregressive1D) has no reference
implementation ("parity unpinned"): defined here as S = Re<Obar^* Obar>/B + lambda I,
F = Re(Obar^H conj(y)) with Obar the centred Jacobian of the *complex* log psi w.r.t. real parameters.
"""
import numpy as np


def centre(O):
    return O - O.mean(axis=0, keepdims=True)


def energy_grad(O_bar, y_true):
    """F = Obar^H conj(y_true), y_true = conj(E_loc - E)/B (optimizer.py:102-108)."""
    return O_bar.conj().T @ np.conj(y_true)


def s_matrix(O_bar, diag_shift):
    B = O_bar.shape[0]
    return O_bar.conj().T @ O_bar / B + diag_shift * np.eye(O_bar.shape[1], dtype=O_bar.dtype)


def solve_direct(O_bar, rhs, diag_shift):
    return np.linalg.solve(s_matrix(O_bar, diag_shift), rhs)


def conjugate_gradient(apply, rhs, tol=1e-3, max_iter=200):
    x = np.zeros_like(rhs)
    r = rhs.copy()
    p = r.copy()
    gamma = np.vdot(r, r)
    tol_abs = tol * np.linalg.norm(r)
    i = 0
    while (max_iter is None or i < max_iter) and np.linalg.norm(r) > tol_abs:
        z = apply(p)
        alpha = gamma / np.vdot(p, z)
        x = x + alpha * p
        r = r - alpha * z
        gamma_new = np.vdot(r, r)
        p = r + (gamma_new / gamma) * p
        gamma = gamma_new
        i += 1
    return x, i, np.linalg.norm(r)


def solve_iterative(O_bar, rhs, diag_shift, tol=1e-3, max_iter=200):
    B = O_bar.shape[0]
    return conjugate_gradient(lambda v: O_bar.conj().T @ (O_bar @ v) / B + diag_shift * v, rhs, tol, max_iter)


def real_sr_system(O_re, O_im, eloc, diag_shift):
    """Real-parameter SR: O = O_re + i O_im is d log psi / d theta (theta real).
    S = Re(Obar^H Obar)/B + lambda I,  F = Re(Obar^H (E_loc - E))/B."""
    B = O_re.shape[0]
    Ob = centre(O_re + 1j * O_im)
    S = np.real(Ob.conj().T @ Ob) / B + diag_shift * np.eye(Ob.shape[1])
    F = np.real(Ob.conj().T @ (eloc - eloc.mean())) / B
    return S, F
