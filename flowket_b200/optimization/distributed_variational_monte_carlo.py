"""Multi-GPU VMC: replaces HorovodVariationalMonteCarlo (flowket/optimization/horovod_variational_monte_carlo.py:10-26).

One process per GPU (torchrun); every rank samples and evaluates its own shard; ONE fp64 allreduce(sum) of
{sum Re E_loc, sum Im E_loc, sum (Re E_loc)^2, count} replaces Horovod's two scalar averages, so the mean is a
true global mean (sum / count), independent of the number of ranks, and the variance is reduced too."""
import numpy

from .variational_monte_carlo import VariationalMonteCarlo


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


class DistributedVariationalMonteCarlo(VariationalMonteCarlo):
    def __init__(self, model, operator, sampler, **kwargs):
        super(DistributedVariationalMonteCarlo, self).__init__(model, operator, sampler, **kwargs)
        dist = _dist()
        self.world_size = dist.get_world_size() if dist else 1
        self.rank = dist.get_rank() if dist else 0
        self.global_batch_size = self.batch_size * self.world_size
        self._shard_sampler(self.sampler)

    def _shard_sampler(self, sampler):
        """Every rank must draw its own slice of the global batch.  The device samplers key their Philox stream by
        (seed, global sample index): a rank's samples are the indices [rank * B, (rank + 1) * B).  A reference Horovod
        script leaves that to per-process unseeded RNGs; here a sampler whose offset was never set is given its rank's
        slice, and an explicit offset that collides with another rank's slice is refused."""
        if self.world_size <= 1 or not hasattr(sampler, 'sample_offset'):
            return
        want = self.rank * sampler.batch_size
        if getattr(sampler, 'shard_rank', None) is None and sampler.sample_offset == 0:
            sampler.sample_offset = want
            sampler.shard_rank = self.rank
        elif sampler.sample_offset != want and getattr(sampler, 'shard_rank', None) != self.rank:
            if self.rank > 0 and sampler.sample_offset < want:
                raise ValueError('rank %d: sampler.sample_offset = %d overlaps the samples of lower ranks (expected %d)'
                                 % (self.rank, sampler.sample_offset, want))

    def set_sampler(self, sampler, mini_batch_size=None):
        res = super(DistributedVariationalMonteCarlo, self).set_sampler(sampler, mini_batch_size)
        self.global_batch_size = self.batch_size * self.world_size
        self._shard_sampler(sampler)
        return res

    @staticmethod
    def reduce_stats(local_energy):
        """local complex E_loc array -> (global mean, global var(Re), global count) via one allreduce."""
        import torch
        lv = numpy.asarray(local_energy)
        stats = torch.tensor([lv.real.sum(), lv.imag.sum(), (lv.real ** 2).sum(), float(lv.size)], dtype=torch.float64)
        dist = _dist()
        if dist is not None:
            if dist.get_backend() == 'nccl':
                stats = stats.cuda()
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)
            stats = stats.cpu()
        s_re, s_im, s_sq, n = [float(v) for v in stats]
        mean = complex(s_re / n, s_im / n)
        return mean, s_sq / n - (s_re / n) ** 2, int(n)

    def _update_batch_local_energy(self):
        _, _, self.current_local_energy = self._estimate()
        self.current_energy, self.current_local_energy_variance, self.global_count = \
            self.reduce_stats(self.current_local_energy)

    def _accept_local_values(self, lv):
        self.current_local_energy = lv
        self.current_energy, self.current_local_energy_variance, self.global_count = self.reduce_stats(lv)

    def next_batch(self):
        batch, _ = super(DistributedVariationalMonteCarlo, self).next_batch()
        return batch, self.loss_coefficients() / self.global_batch_size


HorovodVariationalMonteCarlo = DistributedVariationalMonteCarlo
