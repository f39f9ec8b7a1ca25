"""Observable: E_loc(sigma) = sum_k mel_k exp(log psi(sigma'_k) - log psi(sigma))
(flowket/observables/monte_carlo/operator.py:13-54).

Two routes, same numbers:
  * wave_function is a flowket_b200 Model (or its bound predict): the fused device pipeline fk_local_energy --
    connections generated on the fly from the operator's term table, log psi of all connected configurations
    on the GPU, ratio + segmented sum in the epilogue.  Nothing is materialised, nothing touches the host.
  * any other callable psi(configs) -> ndarray[n,1] (e.g. exact.utils.vector_to_machine): the reference's
    generic protocol, with find_conn on the device (fk_find_conn) and the callable evaluated as given."""
import functools

import numpy

from .observable import BaseObservable


def _model_of(wave_function):
    from ...keras_shim import Model
    if isinstance(wave_function, Model):
        return wave_function
    if isinstance(wave_function, functools.partial):
        return _model_of(wave_function.func)
    owner = getattr(wave_function, '__self__', None)
    if isinstance(owner, Model) and owner.output_kind == 'predictions':
        return owner
    return None


def _ensemble_of(wave_function):
    from ...machines.ensemble import EnsembleModel
    if isinstance(wave_function, EnsembleModel):
        return wave_function
    if isinstance(wave_function, functools.partial):
        return _ensemble_of(wave_function.func)
    owner = getattr(wave_function, '__self__', None)
    return owner if isinstance(owner, EnsembleModel) else None


class Observable(BaseObservable):
    def __init__(self, operator):
        super(Observable, self).__init__()
        self.operator = operator
        self.last_stats = None
        self.last_num_connections = None
        self.count_connections = True    # False: the device route never reads the connection count back (no host sync)

    # ---- device route -------------------------------------------------------------------------------------
    def local_values_device(self, model, configurations):
        """-> complex128 CUDA tensor [B]; also records the fp64 statistics for the multi-GPU allreduce."""
        net = model.machine.device_net()
        sigma = net.to_sigma(configurations)
        eloc, stats, n_conn = net.local_energy(self.operator.device_desc(), sigma, engine=model.engine,
                                               count=self.count_connections)
        self.last_stats, self.last_num_connections = stats, n_conn
        return eloc

    def per_sample_device(self, model):
        """The local energies as a function of the samples, for the split solve of the sharded SR step
        (optimizers/sample_space_sr.py): the pipeline decides which rank evaluates which samples of the global batch."""
        from ...optimizers.sample_space_sr import PerSampleLocalEnergy
        return PerSampleLocalEnergy(lambda configurations: self.local_values_device(model, configurations))

    def local_values_device_generic(self, predict_device, configurations, chunk=1 << 18):
        """Any device wave function (e.g. a symmetrisation ensemble): connections materialised on the device by
        fk_find_conn, log psi of the used ones through `predict_device`, ratios in complex64 and the segmented sum in
        complex128 with torch -- same numbers as the host protocol below, nothing leaves the GPU."""
        import torch
        conn, mel, use = self.operator.find_conn_device(configurations)       # [C,B,*shape], [C,B], [C,B]
        C, B = mel.shape
        shape = tuple(conn.shape[2:])
        flat = conn.permute(1, 0, *range(2, conn.dim())).reshape((B * C,) + shape)     # sample-major: conn 0 first
        used = use.t().reshape(-1)
        idx = torch.nonzero(used, as_tuple=False).reshape(-1)
        logs = torch.empty(idx.numel(), dtype=torch.complex64, device=flat.device)
        for i in range(0, idx.numel(), chunk):
            logs[i:i + chunk] = predict_device(flat[idx[i:i + chunk]]).reshape(-1)
        counts = use.sum(dim=0)
        starts = torch.cumsum(counts, 0) - counts
        sample_of = torch.repeat_interleave(torch.arange(B, device=flat.device), counts)
        ratios = torch.exp(logs - logs[starts][sample_of])                             # complex64, as the reference
        weighted = mel.t().reshape(-1)[idx].to(torch.complex128) * ratios.to(torch.complex128)
        eloc = torch.zeros(B, dtype=torch.complex128, device=flat.device)
        eloc.index_add_(0, sample_of, weighted)
        self.last_num_connections = int(idx.numel())
        self.last_stats = torch.stack([eloc.real.sum(), eloc.imag.sum(), (eloc.real ** 2).sum(),
                                       torch.tensor(float(B), dtype=torch.float64, device=flat.device)])
        return eloc

    # ---- generic route (reference protocol) ------------------------------------------------------------------
    def local_values_optimized_for_unbalanced_local_connections(self, wave_function, local_connections,
                                                                hamiltonian_values, all_use_conn):
        use = numpy.asarray(all_use_conn).astype(bool)
        batch_size = use.shape[1]
        moved = numpy.moveaxis(local_connections, 1, 0).reshape((-1,) + local_connections.shape[2:])
        flat_log_values = numpy.asarray(wave_function(moved[use.T.flatten()]))[:, 0]
        counts = use.sum(axis=0)
        ends = numpy.cumsum(counts)
        starts = ends - counts
        ratios = numpy.exp(flat_log_values - numpy.repeat(flat_log_values[starts], counts))
        weighted = numpy.asarray(hamiltonian_values).T[use.T] * ratios
        local_values = numpy.zeros((batch_size,), dtype=numpy.complex128)
        numpy.add.at(local_values, numpy.repeat(numpy.arange(batch_size), counts), weighted)
        return local_values

    def local_values_optimized_for_balanced_local_connections(self, wave_function, local_connections,
                                                              hamiltonian_values):
        flat_conn = local_connections.reshape((-1,) + local_connections.shape[2:])
        log_values = numpy.asarray(wave_function(flat_conn))[:, 0].reshape(local_connections.shape[0:2])
        return numpy.multiply(hamiltonian_values, numpy.exp(log_values - log_values[0, :])).sum(axis=0)

    def local_values(self, wave_function, configurations):
        model = _model_of(wave_function)
        if model is not None:
            return self.local_values_device(model, configurations).cpu().numpy()
        ensemble = _ensemble_of(wave_function)
        if ensemble is not None:
            return self.local_values_device_generic(ensemble.predict_device, configurations).cpu().numpy()
        local_connections, hamiltonian_values, all_use_conn = self.operator.find_conn(configurations)
        if all_use_conn.mean() < 0.95:
            return self.local_values_optimized_for_unbalanced_local_connections(
                wave_function, local_connections, hamiltonian_values, all_use_conn)
        return self.local_values_optimized_for_balanced_local_connections(wave_function, local_connections,
                                                                          hamiltonian_values)
