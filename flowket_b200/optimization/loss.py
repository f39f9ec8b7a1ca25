"""flowket/optimization/loss.py:4-5.  The CUDA path differentiates this loss directly (fk_grad_weighted);
the function is kept for scripts that pass it to `compile` and for host-side checks."""
import numpy


def loss_for_energy_minimization(y_true, y_pred):
    return 2.0 * numpy.real(numpy.multiply(y_pred, y_true))
