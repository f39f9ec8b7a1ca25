"""ExactVariational with the 2^N states sharded over the ranks of torch.distributed ("ExactVariational
shards the 2^N states the same way").  The reference enumerates everything in one process
(flowket/optimization/exact_variational.py:75-161); here rank r owns the contiguous slice [r 2^N / G, (r+1) 2^N / G):

  * log psi of the slice comes from this rank's GPU (model.predict); one allreduce of the zero-padded table gives every rank the
    whole wave function (2^N complex numbers -- the connected states of a slice live anywhere in the table), from which the
    norm, probabilities and log-probabilities follow redundantly on every rank;
  * find_conn, the connection index table and the local energies are built for the slice only (the C x 2^N tables are the
    memory hog of the exact path: they shrink by G);
  * <H> and the variance are sums over slices: one allreduce of three numbers;
  * the generator yields the slice's (states, coefficients) mini-batches; with Trainer(distributed=True) or
    convert_to_accumulate_gradient_optimizer(use_horovod=True) the accumulated gradients are summed over the ranks, which is
    the exact gradient of the whole enumeration.

Every rank runs `machine_updated()` at the same time (it contains collectives).  Without an initialised process group the
class degenerates to ExactVariational.  The sharding itself (slice ownership, all-gather of the table, all-reduce of the
sums) is implemented once, in exact_variational.py; this module binds it to the process group."""
from .exact_variational import ExactVariational, ExactObservable


def _rank_and_world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class DistributedExactObservable(ExactObservable):
    """ExactObservable over this rank's slice of the states; `current_energy` / variance are global."""


class DistributedExactVariational(ExactVariational):
    def __init__(self, model, operator, batch_size):
        rank, world = _rank_and_world()
        super(DistributedExactVariational, self).__init__(model, operator, batch_size, rank=rank, world_size=world)
