"""Samplers with the reference's constructor signatures and iterator protocol
(flowket/deepar/samplers/base_sampler.py:4-25, fast_autoregressive.py:13-35, autoregressive.py:10-48,
flowket/samplers/__init__.py:8-14).  `__next__` returns a host ndarray like the reference; `next_device()`
returns the int8 CUDA tensor so VariationalMonteCarlo never leaves the device between sampling and E_loc."""
import abc
import copy

import numpy as np


class Sampler(abc.ABC):
    def __init__(self, input_size, batch_size, mini_batch_size=None):
        super(Sampler, self).__init__()
        self.input_size = tuple(input_size)
        self._set_batch_size(batch_size, mini_batch_size=mini_batch_size)

    def _set_batch_size(self, batch_size, mini_batch_size=None):
        if mini_batch_size is None:
            mini_batch_size = batch_size
        if batch_size < mini_batch_size:
            mini_batch_size = batch_size
        self.batch_size = batch_size
        self.mini_batch_size = mini_batch_size

    def __iter__(self):
        return self

    @abc.abstractmethod
    def __next__(self):
        pass


class _DeviceAutoregressiveSampler(Sampler):
    naive = False

    def __init__(self, conditional_log_probs_machine, batch_size, mini_batch_size=None, seed=1234, sample_offset=0,
                 engine=None, **kwargs):
        super(_DeviceAutoregressiveSampler, self).__init__(
            input_size=conditional_log_probs_machine.input_shape[1:], batch_size=batch_size,
            mini_batch_size=mini_batch_size)
        self.conditional_log_probs_machine = conditional_log_probs_machine
        self.machine = conditional_log_probs_machine.machine
        self.seed = seed
        self.sample_offset = sample_offset   # global index of this rank's first sample (multi-GPU sharding)
        self.shard_rank = None               # set by DistributedVariationalMonteCarlo: offset = shard_rank * batch_size
        self.engine = engine                 # None: follow the model's engine; FK_ENGINE_FP32 / FK_ENGINE_TC
        self._draws = 0
        self.last_p0 = None

    def copy_with_new_batch_size(self, batch_size, mini_batch_size=None):
        new_sampler = copy.copy(self)
        new_sampler._set_batch_size(batch_size, mini_batch_size)
        if getattr(self, 'shard_rank', None) is not None:      # keep the shards disjoint at the new batch size
            new_sampler.sample_offset = self.shard_rank * batch_size
        return new_sampler

    def _effective_batch(self):
        # FastAutoregressiveSampler drops batch % mini_batch samples (fast_autoregressive.py:31-33)
        if self.mini_batch_size < self.batch_size:
            return (self.batch_size // self.mini_batch_size) * self.mini_batch_size
        return self.batch_size

    def next_device(self, uniforms=None, return_p0=False):
        """int8 CUDA tensor [B, *input_size].  `uniforms` (float64 [B, *input_size]) switches to the explicit
        random numbers of AutoregressiveSampler (autoregressive.py:31); otherwise Philox4x32-10 keyed by
        (seed + draw counter; global sample index, site)."""
        net = self.machine.device_net()
        B = self._effective_batch()
        if uniforms is not None:
            import torch
            uniforms = torch.as_tensor(np.asarray(uniforms, np.float64)) if not hasattr(uniforms, 'is_cuda') else uniforms
            B = uniforms.shape[0]
        engine = getattr(self, 'engine', None)
        if engine is None:
            engine = getattr(self.conditional_log_probs_machine, 'engine', 0)
        res = net.sample(B, uniforms=uniforms, seed=self.seed + self._draws, sample_offset=self.sample_offset,
                         naive=self.naive, return_p0=return_p0, engine=engine)
        self._draws += 1
        if return_p0:
            sigma, self.last_p0 = res
        else:
            sigma = res
        return sigma.reshape((B,) + self.input_size)

    def __next__(self):
        return self.next_device().cpu().numpy()


class FastAutoregressiveSampler(_DeviceAutoregressiveSampler):
    """Cached incremental exact sampler (one persistent CUDA kernel; fk_sample)."""
    naive = False


class AutoregressiveSampler(_DeviceAutoregressiveSampler):
    """One full forward per site (fk_sample_naive); +-1 convention of flowket.samplers.AutoregressiveSampler."""
    naive = True

    def __init__(self, conditional_log_probs_machine, batch_size, use_progress_bar=False, autoregressive_ordering=None,
                 zero_base=False, **kwargs):
        if autoregressive_ordering is not None:
            raise NotImplementedError('only the raster ordering is implemented on the B200 path')
        if zero_base:
            raise NotImplementedError('zero_base=True (0/1 spins) is outside the hot path; use the +-1 sampler')
        super(AutoregressiveSampler, self).__init__(conditional_log_probs_machine, batch_size, **kwargs)
        self.use_progress_bar = use_progress_bar
        self.zero_base = zero_base


from .exact_sampler import ExactSampler, WaveFunctionSampler  # noqa: E402
from .metropolis_hastings import MetropolisHastingsSampler, MetropolisHastingsLocal, MetropolisHastingsUniform, \
    MetropolisHastingsHamiltonian, MetropolisHastingsExchange, MetropolisHastingsGlobal  # noqa: E402

__all__ = ['Sampler', 'FastAutoregressiveSampler', 'AutoregressiveSampler', 'ExactSampler', 'WaveFunctionSampler',
           'MetropolisHastingsSampler', 'MetropolisHastingsLocal', 'MetropolisHastingsUniform',
           'MetropolisHastingsHamiltonian', 'MetropolisHastingsExchange', 'MetropolisHastingsGlobal']
