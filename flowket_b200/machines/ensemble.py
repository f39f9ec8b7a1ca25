"""Symmetrisation ensembles (flowket/machines/ensemble.py:14-69): psi_sym(sigma) from the predictions of a base model on
the images of sigma under a symmetry group.

  probabilistic (default):  Re = 1/2 logsumexp_k(2 Re log psi_k) - 1/2 log K,   Im = angle(mean_k exp(i Im log psi_k))
  average:                  log(mean_k psi_k)

`make_2d_obc_invariants` = the dihedral group D4 (four rotations and their left-right flips, ensemble.py:42-49),
`make_up_down_invariant` = global spin flip (ensemble.py:63-69, nests around the previous ensemble exactly as the
reference does), `make_pbc_invariants` = all lattice translations (ensemble.py:52-60).  The result duck-types the Keras
model the reference returns: `.predict(x)`, `.input_shape`; everything is evaluated on the device through the base
model's engine (`predict_device`), the transforms are index permutations of int8 configurations."""
import itertools
import math

import numpy as np


def probabilistic_ensemble_op(x):
    """x: complex [n, K] log-amplitudes -> complex [n, 1]   (ensemble.py:14-21)"""
    import torch
    re = 0.5 * torch.logsumexp(2.0 * x.real, dim=-1, keepdim=True) - 0.5 * math.log(x.shape[-1])
    phase = torch.exp(torch.complex(torch.zeros_like(x.imag), x.imag)).mean(dim=-1, keepdim=True)
    return torch.complex(re, torch.atan2(phase.imag, phase.real))


def average_ensemble_op(x):
    """log(mean_k exp(x_k))   (ensemble.py:24-25)"""
    import torch
    m = x.real.amax(dim=-1, keepdim=True)
    mean = torch.exp(x - m).mean(dim=-1, keepdim=True)
    return torch.complex(torch.log(mean.abs()) + m, torch.atan2(mean.imag, mean.real))


class EnsembleModel(object):
    """predictions of `base` on `transforms(sigma)`, combined; `base` is a flowket_b200 Model or another EnsembleModel"""
    output_kind = 'predictions'

    def __init__(self, base, transforms, probabilistic=True, name='ensemble'):
        self.base, self.transforms, self.probabilistic, self.name = base, list(transforms), probabilistic, name

    @property
    def machine(self):
        return self.base.machine

    @property
    def engine(self):
        return self.base.engine

    @engine.setter
    def engine(self, value):
        self.base.engine = value

    @property
    def input_shape(self):
        return self.base.input_shape

    @property
    def ensemble_size(self):
        inner = self.base.ensemble_size if isinstance(self.base, EnsembleModel) else 1
        return len(self.transforms) * inner

    def predict_device(self, x, batch_size=None):
        """x: [n, *shape] (+-1; numpy or torch) -> complex64 CUDA tensor [n, 1]"""
        import torch
        if not isinstance(x, torch.Tensor):
            x = torch.from_numpy(np.ascontiguousarray(np.asarray(x).astype(np.int8)))
        x = x.to('cuda').to(torch.int8)
        n, K = x.shape[0], len(self.transforms)
        if n == 0:
            return torch.zeros((0, 1), dtype=torch.complex64, device='cuda')
        # [n, K, *shape] -> [(n K), *shape]: the K images of one configuration are consecutive (ensemble.py:33-37)
        images = torch.stack([t(x) for t in self.transforms], dim=1).reshape((n * K,) + tuple(x.shape[1:])).contiguous()
        y = self.base.predict_device(images, batch_size=batch_size).reshape(n, K)
        op = probabilistic_ensemble_op if self.probabilistic else average_ensemble_op
        return op(y.to(torch.complex64))

    def predict(self, x, batch_size=None, **_unused):
        return self.predict_device(x, batch_size=batch_size).cpu().numpy()

    __call__ = predict


def _rot90(k):
    import torch
    return lambda x: torch.rot90(x, k, dims=(1, 2))      # tf.image.rot90: counter-clockwise, like numpy/torch


def make_2d_obc_invariants(keras_input_layer, predictions_model, probabilistic=True):
    shape = tuple(keras_input_layer.shape)
    assert len(shape) == 2 and shape[0] == shape[1], 'the D4 ensemble needs a square 2-D lattice'
    import torch
    rotations = [_rot90(k) for k in range(4)]
    flipped = [(lambda x, r=r: torch.flip(r(x), dims=(2,))) for r in rotations]      # FlipLeftRight of every rotation
    return EnsembleModel(predictions_model, rotations + flipped, probabilistic, name='obc_invariants')


def make_up_down_invariant(keras_input_layer, predictions_model, probabilistic=True):
    return EnsembleModel(predictions_model, [lambda x: x, lambda x: -x], probabilistic, name='up_down_invariant')


def make_pbc_invariants(keras_input_layer, predictions_model, apply_also_obc_invariants=True, probabilistic=True):
    import torch
    shape = tuple(keras_input_layer.shape)
    if apply_also_obc_invariants:
        predictions_model = make_2d_obc_invariants(keras_input_layer, predictions_model, probabilistic)
    dims = tuple(range(1, len(shape) + 1))
    rolls = [(lambda x, s=s: torch.roll(x, shifts=s, dims=dims)) for s in itertools.product(*[range(d) for d in shape])]
    return EnsembleModel(predictions_model, rolls, probabilistic, name='pbc_invariants')
