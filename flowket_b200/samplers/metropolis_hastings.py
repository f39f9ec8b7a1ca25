"""Metropolis-Hastings samplers (flowket/samplers/metropolis_hastings.py:15-195) -- a statistical cross-check of
the exact autoregressive samplers, not part of the hot path.  Proposals and the
accept/reject step are host numpy over `num_of_chains` parallel chains; log psi of every candidate batch is one
call of `machine.predict`, i.e. the CUDA forward when `machine` is a flowket_b200 Model.

Differences from the reference, on purpose: randomness comes from a private `numpy.random.Generator(seed)`
instead of the global numpy state, the chatty prints are gone, and r-hat is computed without a Python loop
over lags when there is a single lag."""
import abc

import numpy

from . import Sampler


def sum_correlations(correlations):
    """Geyer's initial-positive-sequence truncation: stop before the first adjacent pair with a negative sum
    (metropolis_hastings.py:8-12)."""
    for i in range(1, correlations.shape[0] // 2):
        if correlations[2 * i - 1] + correlations[2 * i] < 0:
            return correlations[:2 * i].sum()
    return correlations.sum()


class MetropolisHastingsSampler(Sampler):
    def __init__(self, machine, batch_size, num_of_chains=1, unused_sampels=0, discard_ratio=10, seed=None, **kwargs):
        super(MetropolisHastingsSampler, self).__init__(input_size=machine.input_shape[1:], batch_size=batch_size,
                                                        **kwargs)
        if self.batch_size % num_of_chains != 0:
            raise Exception('Num of samplers must divide the batch size')
        self.machine = machine
        self.num_of_chains = num_of_chains
        self.unused_sampels = unused_sampels      # sweeps thrown away between two kept samples (reference spelling)
        self.discard_ratio = discard_ratio        # warm-up = samples_per_chain // discard_ratio kept-sample periods
        self.rng = numpy.random.default_rng(seed)
        self.sample = self.rng.choice([-1, 1], size=(num_of_chains,) + self.input_size)
        self.candidates = numpy.copy(self.sample)
        self.accepts = numpy.zeros((num_of_chains,), dtype=bool)
        self.acceptance_ratio = 1.0
        self.sample_machine_values = None

    def _log_psi(self, configurations):
        return numpy.asarray(self.machine.predict(configurations, batch_size=self.mini_batch_size))[:, 0]

    def machine_updated(self):
        self.sample_machine_values = self._log_psi(self.sample)

    @abc.abstractmethod
    def _sweep(self):
        """One proposal + accept/reject for every chain; returns the number of accepted moves."""

    def _accept(self, log_acceptance, candidates_machine_values):
        numpy.greater(numpy.exp(numpy.minimum(log_acceptance, 0.0)), self.rng.uniform(size=self.num_of_chains),
                      out=self.accepts)
        self.sample[self.accepts, ...] = self.candidates[self.accepts, ...]
        self.sample_machine_values[self.accepts] = candidates_machine_values[self.accepts]
        return int(self.accepts.sum())

    def warn_up(self, num_of_iterations):
        for _ in range(num_of_iterations * self.unused_sampels):
            self._sweep()

    def __next__(self):
        per_chain = self.batch_size // self.num_of_chains
        batch = numpy.empty((self.num_of_chains, per_chain) + self.input_size)
        self.machine_updated()
        if self.discard_ratio > 0:
            self.warn_up(per_chain // self.discard_ratio)
        accepted = 0
        for i in range(per_chain):
            for _ in range(self.unused_sampels + 1):
                accepted += self._sweep()
            batch[:, i, ...] = self.sample
        self.acceptance_ratio = accepted / float(self.batch_size * (self.unused_sampels + 1))
        return batch.reshape((self.batch_size,) + self.input_size)      # chain-major, like the reference

    def calc_r_hat_value(self, estimated_values):
        """Gelman-Rubin potential scale reduction, the pooled variance, the truncated autocorrelation sum and the
        effective sample size of a per-sample estimate laid out chain-major (BDA3 p. 285;
        metropolis_hastings.py:66-91).  -> (r_hat, variance, correlations_sum, effective_sample_size)"""
        n = self.batch_size // self.num_of_chains
        chains = numpy.asarray(estimated_values).reshape((self.num_of_chains, n))
        chain_means = chains.mean(axis=1)
        with numpy.errstate(divide='ignore', invalid='ignore'):
            between = n / (self.num_of_chains - 1.0) * numpy.square(chain_means - chain_means.mean()).sum()
        if n == 1:
            return 1.0, between, 0.0, self.batch_size
        within = numpy.square(chains - chain_means[:, None]).sum(axis=1).mean() / (n - 1)
        variance = ((n - 1) * within + between) / n
        r_hat = numpy.sqrt(variance / within)
        variogram = numpy.array([numpy.square(chains[:, t:] - chains[:, :n - t]).mean() for t in range(1, n)])
        correlations_sum = sum_correlations(1 - variogram / (2 * variance))
        return r_hat, variance, correlations_sum, self.batch_size / (1 + 2 * correlations_sum)


class MetropolisHastingsSymmetricProposal(MetropolisHastingsSampler):
    """q(s'|s) = q(s|s'): accept with min(1, |psi(s')|^2 / |psi(s)|^2) (metropolis_hastings.py:101-117)."""

    @abc.abstractmethod
    def _next_candidates(self):
        pass

    def _sweep(self):
        self._next_candidates()
        values = self._log_psi(self.candidates)
        if not numpy.all(numpy.isfinite(values)):
            raise Exception('candidates_machine_values has not finite element')
        return self._accept(2.0 * (numpy.real(values) - numpy.real(self.sample_machine_values)), values)


class MetropolisHastingsLocal(MetropolisHastingsSymmetricProposal):
    """Flip one uniformly chosen spin per chain (metropolis_hastings.py:122-127)."""

    def _next_candidates(self):
        numpy.copyto(self.candidates, self.sample)
        flat = self.candidates.reshape(self.num_of_chains, -1)
        site = self.rng.integers(flat.shape[1], size=self.num_of_chains)
        flat[numpy.arange(self.num_of_chains), site] *= -1


class MetropolisHastingsExchange(MetropolisHastingsSymmetricProposal):
    """Swap a uniformly chosen site with a site displaced by -1/0/+1 along every axis, periodic wrap
    (conserves total S_z; metropolis_hastings.py:133-145)."""

    def _next_candidates(self):
        chains = numpy.arange(self.num_of_chains)
        first, second = [chains], [chains]
        for dim_size in self.sample.shape[1:]:
            position = self.rng.integers(dim_size, size=self.num_of_chains)
            move = self.rng.integers(-1, 2, size=self.num_of_chains)
            first.append(position)
            second.append((position + move) % dim_size)
        first, second = tuple(first), tuple(second)
        numpy.copyto(self.candidates, self.sample)
        held = self.candidates[first].copy()
        self.candidates[first] = self.candidates[second]
        self.candidates[second] = held


class MetropolisHastingsUniform(MetropolisHastingsSymmetricProposal):
    """Independent uniform proposals (metropolis_hastings.py:163-167)."""

    def _next_candidates(self):
        self.candidates = self.rng.choice([-1, 1], size=self.sample.shape)


class MetropolisHastingsGlobal(MetropolisHastingsSymmetricProposal):
    """Proposals drawn from another sampler, one chain per sample of its batch (metropolis_hastings.py:148-157).
    Like the reference it uses the symmetric acceptance rule, i.e. it assumes the proposal distribution is flat."""

    def __init__(self, machine, batch_size, global_sampler, **kwargs):
        super(MetropolisHastingsGlobal, self).__init__(machine, batch_size, num_of_chains=global_sampler.batch_size,
                                                       **kwargs)
        self.global_sampler = global_sampler

    def _next_candidates(self):
        self.candidates = numpy.asarray(next(self.global_sampler))


class MetropolisHastingsHamiltonian(MetropolisHastingsSampler):
    """Propose one of the operator's connected configurations uniformly; the Hastings factor is the ratio of the
    numbers of connections of the two states (metropolis_hastings.py:173-195).  `find_conn` is the device
    kernel when `hamiltonian` is a flowket_b200 operator."""

    def __init__(self, machine, batch_size, hamiltonian, **kwargs):
        super(MetropolisHastingsHamiltonian, self).__init__(machine, batch_size, **kwargs)
        self.hamiltonian = hamiltonian
        self.sample = numpy.array(self.hamiltonian.random_states(self.num_of_chains), dtype=self.sample.dtype)
        self.candidates = numpy.copy(self.sample)

    def _sweep(self):
        all_conn, _, use_conn = self.hamiltonian.find_conn(self.sample)
        use_conn = numpy.asarray(use_conn, dtype=bool)
        num_of_conn = use_conn.sum(axis=0)
        # k-th used connection of every chain, k uniform in [0, n_conn): rank the used slots with a cumulative sum
        pick = (self.rng.uniform(size=self.num_of_chains) * num_of_conn).astype(numpy.int64)
        pick = numpy.minimum(pick, num_of_conn - 1)
        rank = numpy.cumsum(use_conn, axis=0) - 1
        slot = numpy.argmax(use_conn & (rank == pick[None, :]), axis=0)
        self.candidates = numpy.array(all_conn[slot, numpy.arange(self.num_of_chains), ...], dtype=self.sample.dtype)
        _, _, candidates_use_conn = self.hamiltonian.find_conn(self.candidates)
        candidates_num_of_conn = numpy.asarray(candidates_use_conn, dtype=bool).sum(axis=0)
        values = self._log_psi(self.candidates)
        log_acceptance = 2.0 * (numpy.real(values) - numpy.real(self.sample_machine_values)) \
            + numpy.log(num_of_conn / candidates_num_of_conn)
        return self._accept(log_acceptance, values)
