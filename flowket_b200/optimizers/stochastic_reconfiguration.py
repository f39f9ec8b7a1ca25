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


def sr_delta(O, y_true, diag_shift=0.05, iterative_solver=False, conjugate_gradient_tol=1e-3, max_iterations=200,
             distributed=False, gram=None, stats=None, y_over_global_batch=False):
    """delta = (Obar^H Obar / B + lambda I)^-1 Obar^H conj(y)   for the rows `O` [B_local, P] (complex or real torch tensor, any
    device) and `y_true` = conj(E_loc - E) / B_local of this rank (optimizer.py:33-124).
    distributed: the batch is sharded over the ranks of torch.distributed -- the mean of O, the right-hand side and either the
    P x P matrix (direct) or every matrix-vector product (CG) are summed over the ranks (NCCL on CUDA tensors, gloo on CPU);
    every rank returns the same delta as a single process would on the concatenated batch.
    y_over_global_batch: y_true is already divided by the global batch (DistributedVariationalMonteCarlo.next_batch) instead
    of this rank's batch (the reference's per-rank convention).
    gram: callable X [B, K] real -> X^T X (e.g. the device kernels); default torch matmul."""
    import torch
    from .trainer import allreduce_sum_
    red = allreduce_sum_ if distributed else (lambda t: t)
    B_local = O.shape[0]
    B = int(red(torch.tensor([float(B_local)], dtype=torch.float64, device=O.device)).item())
    mean = red(O.sum(dim=0)) / B
    O_bar = O - mean
    y = torch.as_tensor(y_true).to(device=O.device, dtype=O.dtype)
    if not y_over_global_batch:
        y = y * (float(B_local) / B)
    F = red(O_bar.conj().T @ torch.conj(y))
    if iterative_solver:
        def apply(v):
            return red(O_bar.conj().T @ (O_bar @ v)) / B + diag_shift * v
        x, it, res = conjugate_gradient(apply, F, conjugate_gradient_tol, max_iterations)
        if stats is not None:
            stats['iterations'], stats['residual_norm'] = it, res
        return x
    if gram is None:
        gram = lambda X: X.T @ X   # noqa: E731
    if O_bar.is_complex():
        # Obar^H Obar = (R^T R + I^T I) + i (R^T I - I^T R): one real Gram of the stacked matrix [R | I]
        R, I = O_bar.real.contiguous(), O_bar.imag.contiguous()
        G = red(gram(torch.cat([R, I], dim=1)))
        P = R.shape[1]
        S = torch.complex(G[:P, :P] + G[P:, P:], G[:P, P:] - G[P:, :P]) / B
    else:
        S = red(gram(O_bar)) / B
    S = S + diag_shift * torch.eye(S.shape[0], dtype=S.dtype, device=S.device)
    L = torch.linalg.cholesky(S)
    return torch.cholesky_solve(F.reshape(-1, 1).to(S.dtype), L).reshape(-1)


def real_sr_delta(R, I, e, diag_shift=0.05, iterative_solver=False, conjugate_gradient_tol=1e-3, max_iterations=200,
                  distributed=False, gram=None, stats=None):
    """Real-parameter SR for a complex log psi:  S = Re(Obar^H Obar)/B + lambda I = (Rbar^T Rbar + Ibar^T Ibar)/B + lambda I,
    F = Re(Obar^H (E - Ebar))/B, with R = d Re log psi / d theta, I = d Im log psi / d theta [B_local, P] and the local
    energies e [B_local] of this rank; `distributed` sums the means, F and S (or every CG product) over the ranks."""
    import torch
    from .trainer import allreduce_sum_
    red = allreduce_sum_ if distributed else (lambda t: t)
    B = int(red(torch.tensor([float(R.shape[0])], dtype=torch.float64, device=R.device)).item())
    e = torch.as_tensor(e).to(device=R.device, dtype=torch.complex128)
    e = e - red(e.sum().reshape(1)) / B
    X = torch.cat([R - red(R.sum(dim=0)) / B, I - red(I.sum(dim=0)) / B], dim=0)
    ep = torch.cat([e.real, e.imag]).to(X.dtype)
    F = red(X.T @ ep) / B
    if iterative_solver:
        def apply(v):
            return red(X.T @ (X @ v)) / B + diag_shift * v
        x, it, res = conjugate_gradient(apply, F, conjugate_gradient_tol, max_iterations)
        if stats is not None:
            stats['iterations'], stats['residual_norm'] = it, res
        return x
    if gram is None:
        gram = lambda Y: Y.T @ Y   # noqa: E731
    S = red(gram(X)) / B
    S = S + diag_shift * torch.eye(S.shape[0], dtype=S.dtype, device=S.device)
    L = torch.linalg.cholesky(S)
    return torch.cholesky_solve(F.reshape(-1, 1).to(S.dtype), L).reshape(-1)


def distributed_cholesky_solve(T, b, block=1024):
    """w = T^-1 b (fp64) for a symmetric positive-definite `T` that is REPLICATED on every rank of torch.distributed, with
    the factorisation shared between the ranks instead of repeated on each of them (the replicated 2B x 2B Cholesky is
    the part of the sharded SR step that does not shrink with the number of GPUs).

    Right-looking blocked Cholesky, block columns dealt round-robin to the ranks: the owner of block column k factorises
    its diagonal block and solves the panel below it (`L_ik = A_ik L_kk^-T`), broadcasts the panel (the only
    communication: n^2/2 numbers in total), and every rank applies the rank-`block` update to the block columns it owns.
    Because every panel is broadcast, every rank ends up holding the whole factor L and finishes with the same two
    triangular solves -- the result is identical on all ranks.  Only the lower triangle of T is read."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    n = T.shape[0]
    nblk = (n + block - 1) // block
    start = [k * block for k in range(nblk)] + [n]
    # my block columns, rows from the diagonal block down, promoted to fp64 one column at a time
    mine = {k: T[start[k]:, start[k]:start[k + 1]].to(dtype=torch.float64, copy=True) for k in range(nblk) if k % world == rank}
    L = torch.zeros((n, n), dtype=torch.float64, device=T.device)
    for k in range(nblk):
        owner = k % world
        width = start[k + 1] - start[k]
        panel = L[start[k]:, start[k]:start[k + 1]]           # view into the factor: [(n - start_k), width]
        if rank == owner:
            col = mine.pop(k)
            Lkk = torch.linalg.cholesky(col[:width])
            col[:width] = Lkk
            if col.shape[0] > width:                           # A_ik L_kk^-T  ==  solve X L_kk^T = A_ik
                col[width:] = torch.linalg.solve_triangular(Lkk.T, col[width:], upper=True, left=False)
            panel.copy_(col)
        if world > 1:
            buf = panel.contiguous()
            dist.broadcast(buf, src=owner)
            if rank != owner:
                panel.copy_(buf)
        for j in mine:                                          # trailing update of the block columns this rank still owns
            if j > k:
                off = start[j] - start[k]
                mine[j] -= panel[off:] @ panel[off:off + (start[j + 1] - start[j])].T
    return torch.cholesky_solve(b.double().reshape(-1, 1), L).reshape(-1)


def sample_space_sr_delta(X, ep, diag_shift=0.05, distributed=False, gram=None, low_precision=None, timings=None,
                          shared_cholesky=False):
    """Sample-space SR, optionally with the batch sharded over the ranks of torch.distributed.

    `X` = [Re Obar_r ; Im Obar_r], the rows of *this rank* ([2 B_r, P], already centred with the global means), `ep` =
    [Re(E - Ebar) ; Im(E - Ebar)] of the same rows.  Returns delta = X_all^T (X_all X_all^T / B + lambda I)^-1 ep_all / B
    for the concatenation over the ranks, identical on every rank (push-through form of optimizer.py:33-124).

    The 2B x 2B Gram contracts over the parameter axis, which every rank holds in full for its own rows only.  One
    all-to-all re-shards X from sample-major to parameter-major (rank g receives columns [g P/G, (g+1) P/G) of every
    rank's rows -- as many bytes as it sends), each rank forms the partial Gram of its column slice on its tensor cores,
    one allreduce sums the partial Grams (fp32, (2B)^2), every rank factorises the same fp64 system, and the update
    delta = sum_r X_r^T w_r is one more allreduce of P numbers.  Shards may be ragged.

    shared_cholesky: factorise the replicated system with distributed_cholesky_solve (block columns dealt to the ranks)
    instead of once per rank.
    gram: callable Y [n, k] -> Y Y^T in fp32 (default: torch matmul in the rows' dtype, at least fp32); low_precision: torch dtype the rows are
    rounded to before the all-to-all and the Gram (bf16 halves the exchange and runs the GEMM at tensor-core rate)."""
    import torch
    import torch.distributed as dist
    from .trainer import allreduce_sum_
    if gram is None:
        def gram(Y):
            Y = Y if Y.dtype in (torch.float32, torch.float64) else Y.float()
            return Y @ Y.T
    rows, P = X.shape
    world = dist.get_world_size() if distributed and dist.is_initialized() else 1
    dev = X.device
    mark = (lambda: torch.cuda.Event(enable_timing=True)) if (timings is not None and X.is_cuda) else None
    events = []

    def stamp():
        if mark is not None:
            ev = mark()
            ev.record()
            events.append(ev)
    stamp()
    if world == 1:
        Y = X if low_precision is None else X.to(low_precision)
        rows_of = [rows]
        ep_all = ep
    else:
        rank = dist.get_rank()
        counts = torch.zeros(world, dtype=torch.int64, device=dev)
        counts[rank] = rows
        rows_of = [int(c) for c in allreduce_sum_(counts).tolist()]
        align = 64 * world                       # every column slice starts 16-byte aligned and is a multiple of 64 wide
        P_pad = (P + align - 1) // align * align
        slice_cols = P_pad // world
        dt = X.dtype if low_precision is None else low_precision
        send = torch.zeros((world, rows, slice_cols), dtype=dt, device=dev)        # [destination][row][column of the slice]
        for g in range(world):
            lo, hi = g * slice_cols, min(P, (g + 1) * slice_cols)
            if hi > lo:
                send[g, :, :hi - lo].copy_(X[:, lo:hi])
        Y = torch.empty((sum(rows_of), slice_cols), dtype=dt, device=dev)
        dist.all_to_all_single(Y, send.reshape(world * rows, slice_cols), output_split_sizes=rows_of,
                               input_split_sizes=[rows] * world)
        del send
        ep_all = torch.zeros(sum(rows_of), dtype=ep.dtype, device=dev)      # ragged gather = sum of disjoint slices
        first = sum(rows_of[:rank])
        ep_all[first:first + rows] = ep
        allreduce_sum_(ep_all)
    stamp()
    T = gram(Y)
    if world > 1:
        allreduce_sum_(T)
    del Y
    stamp()
    B = sum(rows_of) // 2
    T = T / B
    T.diagonal().add_(diag_shift)
    if shared_cholesky and world > 1:
        w = distributed_cholesky_solve(T, ep_all.double() / B).to(X.dtype)      # factorisation shared between the ranks
    else:
        L = torch.linalg.cholesky(T.double())          # fp64: the Gram can be badly conditioned
        w = torch.cholesky_solve((ep_all.double() / B).reshape(-1, 1), L).reshape(-1).to(X.dtype)
    stamp()
    if world == 1:
        delta = X.T @ w
    else:
        delta = allreduce_sum_(X.T @ w[first:first + rows])
    stamp()
    if events:
        torch.cuda.synchronize()
        names = ('exchange', 'gram', 'cholesky', 'update')
        timings.update({n: events[i].elapsed_time(events[i + 1]) for i, n in enumerate(names)})
    return delta


class _SRBase(object):
    def __init__(self, model, lr=0.01, diag_shift=0.05, iterative_solver=True, conjugate_gradient_tol=1e-3,
                 iterative_solver_max_iterations=200, use_cholesky=True, distributed=False):
        self.distributed = distributed      # samples sharded over the ranks: reductions of the SR system go over NCCL
        self.y_over_global_batch = True     # y_true as produced by DistributedVariationalMonteCarlo.next_batch
        self.model, self.machine = model, model.machine
        self.lr, self.diag_shift = lr, diag_shift
        self.iterative_solver = iterative_solver
        self.conjugate_gradient_tol = conjugate_gradient_tol
        self.iterative_solver_max_iterations = iterative_solver_max_iterations
        self.use_cholesky = use_cholesky
        self.conjugate_gradient_iterations = 0
        self.conjugate_gradient_residual_norm = 0.0

    def _gram_engine(self):
        """tcgen05 Gram (fp16 hi/lo operands, fp32 accumulation) when the model runs the tensor-core engine"""
        from .. import _lib
        return getattr(self.model, 'engine', _lib.FK_ENGINE_FP32)

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
            G = sr_gram(torch.cat([R, I], dim=1), transpose_a=True, engine=self._gram_engine())      # [2P, 2P]
            P = R.shape[1]
            S = torch.complex(G[:P, :P] + G[P:, P:], G[:P, P:] - G[P:, :P]) / B
        else:
            S = sr_gram(O_bar.float(), transpose_a=True, engine=self._gram_engine()) / B
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
        y = torch.as_tensor(np.asarray(y_true, np.complex64)).to(O.device)
        if self.distributed:
            from .._device import sr_gram
            stats = {}
            delta = sr_delta(O, y, self.diag_shift, self.iterative_solver, self.conjugate_gradient_tol,
                             self.iterative_solver_max_iterations, distributed=True,
                             gram=lambda X: sr_gram(X.float(), transpose_a=True, engine=self._gram_engine()), stats=stats,
                             y_over_global_batch=self.y_over_global_batch)
            self.conjugate_gradient_iterations = stats.get('iterations', 0)
            self.conjugate_gradient_residual_norm = stats.get('residual_norm', 0.0)
            return delta
        O_bar = O - O.mean(dim=0, keepdim=True)
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
    """Real-parameter SR (ConvNetAutoregressive2D, SimpleConvNetAutoregressive1D).

    With X = [Re Obar ; Im Obar] (2B x P) the system is (X^T X / B + lambda I) delta = X^T e' / B, e' = [Re(E - Ebar) ;
    Im(E - Ebar)].  Three solvers, all the same delta:
      * direct        P x P Gram through fk_sr_gram + Cholesky               (small machines, reference default shape)
      * iterative     conjugate gradient on v -> X^T (X v) / B + lambda v    (linear_equations.py:34-137)
      * sample space  delta = X^T (X X^T / B + lambda I)^-1 e' / B -- the push-through identity, exact, with a
                      2B x 2B Gram whose K dimension is the parameter axis.  This is the form that fits the 850 k
                      parameter machine of the headline configuration (P x P would be 2.9 TB): the Gram is one
                      tensor-core GEMM (hand-written tcgen05 kernel behind fk_sr_gram_xxt, bf16 operands / fp32 accumulation;
                      sample_space_sr.py) + a Cholesky of size 2B.
    `sample_space=None` picks the sample-space form when P > 2B."""

    def __init__(self, model, sample_space=None, gram_dtype='bf16', jacobian_chunk=1024, shared_cholesky=False,
                 device_pipeline=True, read_timings=True, solver='mixed', split_solve=True, **kwargs):
        super(StochasticReconfiguration, self).__init__(model, **kwargs)
        self.split_solve = split_solve           # sharded device pipeline, local energies passed as a function: one rank factors while the others evaluate them
        self.solver = solver                     # device pipeline: 'mixed' (fp32 Cholesky + fp64 refinement) or 'fp64'
        self.device_pipeline = device_pipeline   # False: the torch route below (fp32 rows, torch.mm Gram) -- kept as a cross-check
        self.read_timings = read_timings         # False: no synchronise after the update (timings / potrf status unread)
        self.shared_cholesky = shared_cholesky     # distributed sample-space solve: share the Cholesky between the ranks
        self.sample_space = sample_space
        self.gram_dtype = gram_dtype
        self.jacobian_chunk = jacobian_chunk
        self.last_timings_ms = {}

    def _jacobian_engine(self, net):
        """tensor-core Jacobians when the model asks for the tensor-core engine and the machine is supported"""
        from .. import _lib
        if getattr(self.model, 'engine', _lib.FK_ENGINE_FP32) in (_lib.FK_ENGINE_TC, _lib.FK_ENGINE_TC_EXACT) and \
                net.lib.fk_grad_per_sample_tc_workspace_bytes(net.handle, 1) >= 0:
            return _lib.FK_ENGINE_TC
        return _lib.FK_ENGINE_FP32

    def stacked_jacobian(self, sigma):
        """X = [Re O ; Im O] (uncentred) as one [2B, P] fp32 device tensor, built in chunks of samples."""
        import torch
        net = self.machine.device_net()
        sig = net.to_sigma(sigma)
        B, P = sig.shape[0], net.num_params
        X = torch.empty((2 * B, P), dtype=torch.float32, device=sig.device)
        for b0 in range(0, B, self.jacobian_chunk):
            b1 = min(B, b0 + self.jacobian_chunk)
            O_re, O_im = net.grad_per_sample(sig[b0:b1], imag=True, engine=self._jacobian_engine(net))
            X[b0:b1] = O_re
            X[B + b0:B + b1] = O_im
            del O_re, O_im
        return X

    def _device_pipeline(self):
        """The device-resident sample-space pipeline (sample_space_sr.py: bf16 Jacobian rows -> hand-written tcgen05 Gram ->
        fp64 solve -> X^T w, all behind the C ABI) when the machine has tensor-core Jacobians and the defaults are kept."""
        from .sample_space_sr import DeviceSampleSpaceSR
        from .. import _lib
        if self.gram_dtype != 'bf16' or self.shared_cholesky or not self.device_pipeline:
            return None
        net = self.machine.device_net()
        if self._jacobian_engine(net) != _lib.FK_ENGINE_TC or not DeviceSampleSpaceSR.supported(net):
            return None
        if getattr(self, '_pipeline', None) is None or self._pipeline.net is not net:
            self._pipeline = DeviceSampleSpaceSR(net, self.diag_shift, solver=self.solver)
        self._pipeline.solver = self.solver
        self._pipeline.split_solve = bool(self.split_solve)
        self._pipeline.diag_shift = float(self.diag_shift)
        return self._pipeline

    def compute_update(self, sigma, local_energy):
        import torch
        net0 = self.machine.device_net()
        sample_space = self.sample_space
        if sample_space is None:
            world = self._world_size() if self.distributed else 1
            sample_space = net0.num_params > 2 * int(sigma.shape[0]) * world   # P > 2 B_global
        pipe = self._device_pipeline() if sample_space else None
        use_pipe = pipe is not None and (not self.distributed or (2 * int(sigma.shape[0])) % 128 == 0)
        self._last_local_energy = None
        if callable(local_energy) and not use_pipe:      # only the device pipeline can defer / re-deal the local energies
            from .sample_space_sr import PerSampleLocalEnergy
            local_energy = local_energy(net0.to_sigma(sigma)) if isinstance(local_energy, PerSampleLocalEnergy) else local_energy()
            self._last_local_energy = local_energy
        if use_pipe:
            delta = pipe.delta(sigma, local_energy, distributed=self.distributed)
            self._last_local_energy = pipe.last_local_energy
            if self.read_timings:
                torch.cuda.synchronize()
                self.last_timings_ms = dict(pipe.read_timings())
                if pipe.needs_fp64_solve:
                    delta = pipe.refine_with_fp64()
                if pipe.potrf_info != 0:
                    raise RuntimeError('sample-space SR: the Cholesky factorisation failed (potrf info %d)' % pipe.potrf_info)
            return delta
        if self.distributed:
            if sample_space:
                return self._compute_update_sample_space_distributed(sigma, local_energy)
            from .._device import sr_gram
            net = self.machine.device_net()
            R, I = net.grad_per_sample(net.to_sigma(sigma), imag=True, engine=self._jacobian_engine(net))
            e = local_energy if torch.is_tensor(local_energy) else torch.as_tensor(np.asarray(local_energy, np.complex128))
            stats = {}
            delta = real_sr_delta(R, I, e, self.diag_shift, self.iterative_solver, self.conjugate_gradient_tol,
                                  self.iterative_solver_max_iterations, distributed=True,
                                  gram=lambda Y: sr_gram(Y.float(), transpose_a=True, engine=self._gram_engine()), stats=stats)
            self.conjugate_gradient_iterations = stats.get('iterations', 0)
            self.conjugate_gradient_residual_norm = stats.get('residual_norm', 0.0)
            return delta
        t = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        t[0].record()
        X = self.stacked_jacobian(sigma)
        B = X.shape[0] // 2
        if torch.is_tensor(local_energy):
            e = local_energy.to(device=X.device, dtype=torch.complex128)   # device-resident E_loc: no host round trip
        else:
            e = torch.as_tensor(np.asarray(local_energy, np.complex128)).to(X.device)
        e = e - e.mean()
        X[:B] -= X[:B].mean(dim=0, keepdim=True)
        X[B:] -= X[B:].mean(dim=0, keepdim=True)
        ep = torch.cat([e.real, e.imag]).float()
        t[1].record()
        sample_space = self.sample_space if self.sample_space is not None else X.shape[1] > X.shape[0]
        if sample_space:
            delta = self._solve_sample_space(X, ep, B)
        elif self.iterative_solver:
            delta = self._solve_real(X, X.T @ ep / B, B)
        else:
            delta = self.solve(X, X.T @ ep / B)
        t[2].record()
        torch.cuda.synchronize()
        self.last_timings_ms = {'jacobian': t[0].elapsed_time(t[1]), 'solve': t[1].elapsed_time(t[2])}
        ev = getattr(self, '_solve_events', None)
        if sample_space and ev is not None and self.gram_dtype != 'fp32':
            self.last_timings_ms.update({'convert': ev[0].elapsed_time(ev[1]), 'gram': ev[1].elapsed_time(ev[2]),
                                         'cholesky': ev[2].elapsed_time(ev[3])})
        return delta

    @staticmethod
    def _world_size():
        import torch.distributed as dist
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def _compute_update_sample_space_distributed(self, sigma, local_energy):
        """Sample-space SR with the batch sharded over the ranks: per-rank Jacobians, global centring (two allreduces of
        P numbers), then sample_space_sr_delta (all-to-all re-shard, partial Gram per rank, allreduce, replicated solve)."""
        import torch
        from .trainer import allreduce_sum_
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        X = self.stacked_jacobian(sigma)
        B_local = X.shape[0] // 2
        B = int(allreduce_sum_(torch.tensor([float(B_local)], dtype=torch.float64, device=X.device)).item())
        if torch.is_tensor(local_energy):
            e = local_energy.to(device=X.device, dtype=torch.complex128)
        else:
            e = torch.as_tensor(np.asarray(local_energy, np.complex128)).to(X.device)
        e = e - torch.view_as_complex(allreduce_sum_(torch.view_as_real(e.sum().reshape(1)).clone())) / B
        X[:B_local] -= allreduce_sum_(X[:B_local].sum(dim=0)) / B
        X[B_local:] -= allreduce_sum_(X[B_local:].sum(dim=0)) / B
        # rows of this rank stay [Re ; Im]: the Gram and its right-hand side only need a consistent row order
        ep = torch.cat([e.real, e.imag]).float()
        t1.record()
        if self.gram_dtype not in ('bf16', 'fp32'):
            raise NotImplementedError('distributed sample-space SR exchanges bf16 (default) or fp32 rows')
        low = torch.bfloat16 if self.gram_dtype == 'bf16' else None
        gram = (lambda Y: self._symmetric_gram(Y)) if low is not None else None
        timings = {}
        delta = sample_space_sr_delta(X, ep, self.diag_shift, distributed=True, gram=gram, low_precision=low,
                                      timings=timings, shared_cholesky=self.shared_cholesky)
        torch.cuda.synchronize()
        self.last_timings_ms = dict(timings, jacobian=t0.elapsed_time(t1))
        return delta

    def _solve_sample_space(self, X, ep, B):
        import torch
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        if self.gram_dtype == 'fp32':
            T = X @ X.T
        else:
            dt = {'bf16': torch.bfloat16, 'fp16': torch.float16}[self.gram_dtype]
            P = X.shape[1]
            P_pad = (P + 127) // 128 * 128     # 16-byte aligned rows: cuBLAS takes its slow path otherwise (4.7x here)
            Xl = torch.empty((X.shape[0], P_pad), dtype=dt, device=X.device)
            Xl[:, P:] = 0
            if dt == torch.bfloat16:
                scale = torch.ones((), dtype=torch.float32, device=X.device)   # fp32 range: no scaling, one converting copy
                Xl[:, :P].copy_(X)
            else:
                # fp16: scale to O(1) before rounding, undo on the fp32 result; row chunks so that no second fp32 copy
                # of X exists (56 GB at B = 8192 on the 850 k parameter machine)
                rows = 1024
                scale = torch.stack([X[r:r + rows].abs().amax() for r in range(0, X.shape[0], rows)]).amax().clamp_min(1e-30)
                for r in range(0, X.shape[0], rows):
                    Xl[r:r + rows, :P] = X[r:r + rows] / scale
            ev[1].record()
            T = self._symmetric_gram(Xl) * (scale * scale)
            del Xl
        ev[2].record()
        T = T / B
        T.diagonal().add_(self.diag_shift)
        L = torch.linalg.cholesky(T.double())          # 2B x 2B, fp64: the Gram can be badly conditioned
        w = torch.cholesky_solve((ep.double() / B).reshape(-1, 1), L).reshape(-1).float()
        ev[3].record()
        delta = X.T @ w
        self._solve_events = ev
        return delta

    @staticmethod
    def _symmetric_gram(Xl, block=2048):
        """Xl Xl^T in fp32 from low-precision rows, computing only the block upper triangle (the Gram is symmetric)"""
        import torch
        n = Xl.shape[0]
        if n <= 2 * block:
            return torch.mm(Xl, Xl.T, out_dtype=torch.float32)
        T = torch.empty((n, n), dtype=torch.float32, device=Xl.device)
        for i in range(0, n, block):
            # one GEMM per block row: columns i.. only
            blk = torch.mm(Xl[i:i + block], Xl[i:].T, out_dtype=torch.float32)
            T[i:i + block, i:] = blk
            if i + block < n:
                T[i + block:, i:i + block] = blk[:, block:].T
        return T

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
        S = sr_gram(stacked.float(), transpose_a=True, engine=self._gram_engine()) / B
        S = S + self.diag_shift * torch.eye(S.shape[0], dtype=S.dtype, device=S.device)
        L = torch.linalg.cholesky(S)
        return torch.cholesky_solve(rhs.reshape(-1, 1), L).reshape(-1)

    def step(self, sigma, local_energy):
        delta = self.compute_update(sigma, local_energy)
        params = self.machine.flat_params_device()
        params.add_(delta.float(), alpha=-self.lr)
        self.machine.params_updated()
        return delta

    @property
    def last_local_energy(self):
        """this rank's local energies of the last device-pipeline update (they are evaluated inside the update when a
        callable / PerSampleLocalEnergy was passed)"""
        return getattr(self, '_last_local_energy', None)

    def step_generator(self, generator):
        """One SR update driven from a VariationalMonteCarlo generator with the local energies evaluated INSIDE the update
        (split solve of the sharded step): samples from the generator, local energies dealt over the ranks by the pipeline,
        this rank's values handed back to the generator for its energy statistics."""
        x = generator.next_samples()
        delta = self.step(x, generator.local_energy_function())
        e = self.last_local_energy
        if e is None or not hasattr(e, 'is_cuda'):     # the routes that evaluated the function up front do not keep the values
            e = generator.local_energy_function()(self.machine.device_net().to_sigma(x))
        generator.set_local_energy(e)
        return delta
