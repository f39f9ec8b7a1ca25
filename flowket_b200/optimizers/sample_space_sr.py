"""Sample-space stochastic reconfiguration of the headline machine, device-resident end to end.

The reference's SR (flowket/optimizers/stochastic_reconfiguration/optimizer.py:33-124) forms S = Obar^H Obar / B + lambda I
over the parameters; for the 854 k-parameter 10x10 machine the same update is computed in sample space,

    delta = X^T C (C X X^T C / B + lambda I)^-1 e' / B,      X = [Re O ; Im O]  (2B x P),  e' = [Re(E - Ebar) ; Im(E - Ebar)],

with C the per-half centring projector (the "O - mean(O)" of optimizer.py:79-81 applied inside the Gram matrix).  Every
stage is a call into the C ABI:

    fk_jacobian_rows_tc   per-sample Jacobian rows on the tensor cores, weight-norm transform fused, written as bf16 in the
                          panel-major layout the Gram kernel's TMA boxes want (no fp32 copy of X ever exists)
    fk_sr_gram_xxt        hand-written cta_group::2 tcgen05 GEMM, upper block triangle + mirror
    fk_sr_centre_shift    S = C G C / B + lambda I in fp64
    fk_sr_solve_mixed     fp32 Cholesky factor + fp64 iterative refinement (cuSOLVER potrf/potrs behind the ABI; residuals
                          and updates are this library's kernels); `solver='fp64'` selects the plain fp64 fk_sr_solve
    fk_sr_xt_w            delta = X^T (C w), one pass over X at HBM speed

Sharded over the ranks of torch.distributed: each rank produces the rows of its own samples; ONE all-to-all
re-shards X from sample-major to parameter-major (in the panel-major layout a parameter slice is a contiguous range of
panels, so the exchange needs no packing; the received per-rank blocks are consumed in place through the Gram kernel's
row-block tensor map); the partial Gram matrices are summed with one fp32 allreduce, every rank solves the same system, and
the slices of delta are gathered.

Split solve (`PerSampleLocalEnergy`, world > 1): the matrix of the system does not depend on the local energies, and its
factorisation is the one stage of the step that does not shrink with the number of ranks.  When the caller hands over the
local-energy FUNCTION instead of its values, the ranks gather the (tiny) samples after the Gram allreduce, ONE rank factors
S while the others already evaluate local energies, and the samples are dealt so that everybody finishes together: with
rho = T_factor / T_eloc(whole batch) the solver rank takes (1 - (N - 1) rho) / N of the batch (measured with CUDA events on
the previous step, identical on every rank because the measurements are all-gathered).  The right-hand side is assembled
from one allreduce of the local energies, the solver rank runs the triangular solves + refinement and broadcasts w."""
import ctypes

import numpy as np

from .. import _lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class PerSampleLocalEnergy(object):
    """The local energies as a function of the samples, `fn(sigma int8 CUDA [n, sites]) -> complex128 CUDA [n]`, for the split
    solve of the sharded step (the pipeline decides which rank evaluates which samples).  Anywhere else it is simply
    evaluated on the rank's own samples."""

    def __init__(self, fn):
        self.fn = fn

    def __call__(self, sigma):
        return self.fn(sigma)


def split_solve_shares(world, batch, rho, solver_rank=0, kappa=1.0):
    """Samples per rank for the local energies of a global batch when `solver_rank` also factors the matrix.  With T_E the
    time of the whole batch's local energies at the other ranks' per-sample rate, rho = T_factor / T_E and kappa = (per-sample
    time on the solver rank) / (per-sample time on the others; the solver rank's clocks differ: it comes out of the
    factorisation, the others out of the power-capped Gram):  s0 kappa T_E + T_factor = (1 - s0) T_E / (N - 1)
    ->  s0 = (1 / (N - 1) - rho) / (kappa + 1 / (N - 1)), clipped at 0 (kappa = 1: (1 - (N - 1) rho) / N); the other ranks
    share the rest evenly (the first ones take the remainder).  -> list of `world` ints summing to `batch`."""
    if world == 1:
        return [int(batch)]
    inv = 1.0 / (world - 1)
    s0 = max(0.0, (inv - float(rho)) / (max(float(kappa), 1e-3) + inv))
    n0 = min(int(batch), int(round(s0 * batch)))
    rest, others = int(batch) - n0, world - 1
    counts, k = [], 0
    for r in range(world):
        if r == solver_rank:
            counts.append(n0)
        else:
            counts.append(rest // others + (1 if k < rest % others else 0))
            k += 1
    return counts


def deal_local_energies(sigma, fn, counts, rank, world, events=None, after_gather=None):
    """Local energies of the GLOBAL batch with the samples dealt over the ranks: all-gather the ranks' samples (equal
    counts per rank, rank-major = global sample order), evaluate `fn` on samples [sum(counts[:rank]), + counts[rank]) of the
    global batch, assemble with one allreduce.  -> complex128 [B_global], identical on every rank.  `events`: two CUDA events
    recorded around the evaluation (split-solve bookkeeping); `after_gather`: called once the sample gather is enqueued and
    before the evaluation (the solver rank factors there: a collective enqueued behind the factorisation would make every
    other rank wait for it).  Backend-agnostic (NCCL on the device, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    Bl = sigma.shape[0]
    sig_all = torch.empty((Bl * world,) + tuple(sigma.shape[1:]), dtype=sigma.dtype, device=sigma.device)
    dist.all_gather_into_tensor(sig_all, sigma.contiguous())
    if after_gather is not None:
        after_gather()
    lo = sum(counts[:rank])
    hi = lo + counts[rank]
    e_all = torch.zeros(Bl * world, dtype=torch.complex128, device=sigma.device)
    if hi > lo:
        if events is not None:
            events[0].record()
        e_all[lo:hi] = fn(sig_all[lo:hi]).to(torch.complex128)
        if events is not None:
            events[1].record()
    dist.all_reduce(torch.view_as_real(e_all))
    return e_all


class DeviceSampleSpaceSR(object):
    def __init__(self, net, diag_shift, solver='mixed', refinements=3, split_solve=True, solver_rank=0):
        import torch
        self.torch = torch
        self.net = net
        self.lib = net.lib
        self.diag_shift = float(diag_shift)
        self.solver = solver                 # 'mixed': fp32 factor + fp64 refinement;  'fp64': fp64 factor
        self.refinements = int(refinements)   # measured at n = 16384: 2.6e-6 -> 8e-12 -> 3e-15 relative residual at random init, 1.2e-3 -> 3e-6 -> 1e-8 -> ... 60 SR updates into a training run
        self.refinement_tol = 1e-6           # |S x - rhs| / |rhs| the refined solution must reach (checked in read_timings;
                                             # above it -- or when the fp32 factorisation fails -- the solve is repeated in fp64)
        self.split_solve = bool(split_solve)   # world > 1 and a PerSampleLocalEnergy: one rank factors, the others take its samples
        self.solver_rank = int(solver_rank)
        self.split_rho = 0.1                   # T_factor / T_eloc(whole batch): prior, replaced by the measurement of the last step
        self.split_kappa = 1.0                 # per-sample local-energy time on the solver rank relative to the other ranks
        self._split_events = None
        self._solver = None
        self.timings_ms = {}

    @staticmethod
    def supported(net):
        return net.lib.fk_jacobian_rows_tc_workspace_bytes(net.handle, 1) >= 0

    def __del__(self):
        try:
            if self._solver is not None:
                self.lib.fk_sr_solver_destroy(self._solver)
                self._solver = None
        except Exception:
            pass

    def _solver_handle(self):
        if self._solver is None:
            h = ctypes.c_void_p()
            _lib.check(self.lib.fk_sr_solver_create(ctypes.byref(h)))
            self._solver = h
        return self._solver

    def _world(self, distributed):
        import torch.distributed as dist
        if distributed and dist.is_available() and dist.is_initialized():
            return dist.get_world_size(), dist.get_rank()
        return 1, 0

    def _split_measurement(self):
        """[eloc ms, eloc samples, factor ms] of this rank in the last split step (zeros when there is none); the caller has
        synchronised the stream"""
        ev = self._split_events
        out = [0.0, 0.0, 0.0]
        if ev is not None:
            if ev.get('eloc') is not None:
                out[0], out[1] = ev['eloc'][0].elapsed_time(ev['eloc'][1]), float(ev['n'])
            if ev.get('factor') is not None:
                out[2] = ev['factor'][0].elapsed_time(ev['factor'][1])
        return out

    def delta(self, sigma, local_energy, distributed=False):
        """-> delta [P] fp32 (identical on every rank).  sigma: int8 CUDA tensor [B_local, sites] (this rank's samples),
        local_energy: complex128 CUDA tensor [B_local], a zero-argument callable returning it (evaluated between the Gram and
        the solve), or a PerSampleLocalEnergy (split solve when world > 1, see the module docstring)."""
        torch, net, lib = self.torch, self.net, self.lib
        import torch.distributed as dist
        world, rank = self._world(distributed)
        dev = net.device
        sig = net.to_sigma(sigma)
        Bl = sig.shape[0]
        P = net.num_params
        nkb = (P + 63) // 64
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
        stream = _lib.stream_ptr()
        per_sample = isinstance(local_energy, PerSampleLocalEnergy)
        split = per_sample and world > 1 and self.split_solve and self.solver == 'mixed'
        if per_sample and not split:
            local_energy = local_energy(sig)
        if world > 1:
            if split:
                torch.cuda.current_stream().synchronize()      # (the size check below synchronises anyway)
            sizes = torch.tensor([float(Bl)] + (self._split_measurement() if split else [0.0, 0.0, 0.0]), dtype=torch.float64,
                                 device=dev)
            gathered = torch.zeros((world, 4), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(gathered, sizes)
            gathered = gathered.cpu().numpy()
            if any(int(g) != Bl for g in gathered[:, 0]):
                raise ValueError('the device sample-space SR needs the same number of samples on every rank')
            if (2 * Bl) % 128 != 0:
                raise ValueError('the sharded device sample-space SR needs a multiple of 64 samples per rank')
            if split and gathered[self.solver_rank, 3] > 0:
                other = [r for r in range(world) if r != self.solver_rank and gathered[r, 2] > 0]
                if other:      # per-sample time of the other ranks, of the solver rank relative to it, factorisation relative to both
                    tau = gathered[other, 1].sum() / gathered[other, 2].sum()
                    self.split_rho = float(gathered[self.solver_rank, 3] / (tau * Bl * world))
                    if gathered[self.solver_rank, 2] > 0:
                        self.split_kappa = float(gathered[self.solver_rank, 1] / gathered[self.solver_rank, 2] / tau)
        self._split_events = None
        B = Bl * world
        R, Rl = 2 * B, 2 * Bl
        nkb_r = (nkb + world - 1) // world          # panels per rank after the exchange
        nkb_pad = nkb_r * world
        panel_bytes = Rl * 128
        ev[0].record()
        # ---- X rows of this rank: [nkb_pad][Rl][64] bf16 (padding panels zero)
        X = net.workspace('sr_x', nkb_pad * panel_bytes)
        if nkb_pad > nkb:
            X[nkb * panel_bytes:nkb_pad * panel_bytes].zero_()
        wsb = lib.fk_jacobian_rows_tc_workspace_bytes(net.handle, Bl)
        ws = net.workspace('sr_rows', wsb)
        with torch.cuda.device(dev):
            _lib.check(lib.fk_jacobian_rows_tc(net.handle, _ptr(sig), Bl, _ptr(X), Rl, 0, Bl, _ptr(ws), ws.numel(), stream))
        ev[1].record()
        # ---- re-shard: rank g receives panels [g nkb_r, (g + 1) nkb_r) of every rank as one row block each
        if world > 1:
            Xg = net.workspace('sr_xg', nkb_pad * panel_bytes)
            dist.all_to_all_single(Xg[:nkb_pad * panel_bytes], X[:nkb_pad * panel_bytes])
            Kg = min(P, (rank + 1) * nkb_r * 64) - rank * nkb_r * 64
            Kg = max(Kg, 0)
            block_stride = nkb_r * panel_bytes
        else:
            Xg, Kg, block_stride = X, P, 0
        ev[2].record()
        # ---- Gram of this rank's parameter slice over all samples
        G = net.workspace('sr_g', R * R * 4).view(torch.float32)[:R * R]
        gws_b = lib.fk_sr_gram_xxt_workspace_bytes(R)
        gws = net.workspace('sr_gram_ws', gws_b)
        with torch.cuda.device(dev):
            if Kg > 0:
                _lib.check(lib.fk_sr_gram_xxt(_ptr(Xg), R, Kg, Rl, world, block_stride, 1.0, _ptr(G), R, _ptr(gws), gws.numel(),
                                              stream))
            else:
                G.zero_()
        if world > 1:
            dist.all_reduce(G)
        ev[3].record()
        # ---- S = C G C / B + lambda I, right-hand side, solve
        S = net.workspace('sr_s', R * R * 8).view(torch.float64)[:R * R]
        cws_b = lib.fk_sr_centre_shift_workspace_bytes(R)
        cws = net.workspace('sr_centre_ws', cws_b)
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        solver = self._solver_handle()
        is_solver = (not split) or rank == self.solver_rank

        def centre():
            with torch.cuda.device(dev):
                _lib.check(lib.fk_sr_centre_shift(_ptr(G), R, R, world, self.diag_shift, _ptr(S), _ptr(cws), cws.numel(), stream))

        def solve_workspace(mixed):
            sws_b = lib.fk_sr_solve_mixed_workspace_bytes(solver, R) if mixed else lib.fk_sr_solve_workspace_bytes(solver, R)
            if sws_b < 0:
                raise _lib.FlowketB200Error('fk_sr_solve_workspace_bytes failed')
            return net.workspace('sr_solve_ws', sws_b)

        self._eloc_events = None
        if split:
            # the samples of every rank (B x sites bytes), then: the solver rank centres and factors S, everybody evaluates
            # its share of the local energies of the GLOBAL batch, one allreduce assembles them
            counts = split_solve_shares(world, B, self.split_rho, self.solver_rank, self.split_kappa)
            events = {'n': counts[rank], 'eloc': None, 'factor': None}

            def factor():
                fe = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                centre()
                fe[0].record()
                sws = solve_workspace(True)
                with torch.cuda.device(dev):
                    _lib.check(lib.fk_sr_factor_mixed(solver, _ptr(S), R, _ptr(info), _ptr(sws), sws.numel(), stream))
                fe[1].record()
                events['factor'] = fe

            ee = [torch.cuda.Event(enable_timing=True) for _ in range(2)] if counts[rank] > 0 else None
            e_all = deal_local_energies(sig, local_energy, counts, rank, world, ee, after_gather=factor if is_solver else None)
            if ee is not None:
                events['eloc'] = ee
                self._eloc_events = ee
            self._split_events = events
            self.split_counts = counts
            self.last_local_energy = e_all[rank * Bl:(rank + 1) * Bl]
            e = (e_all - e_all.mean()).view(world, Bl)
            rhs = torch.stack([e.real, e.imag], dim=1).reshape(-1) / B      # row order of the blocks: [Re ; Im] per rank
        else:
            # `local_energy` may be a callable: the local energies do not depend on the Jacobian or the Gram matrix, so they can
            # be evaluated HERE, between the Gram and the solve.  The Gram runs the GPU into its power cap (SM clock ~1.2 GHz) and
            # whatever follows it inherits the throttled clock for a while: measured at B = 8192, the factorisation drops from
            # 57-70 ms to 43 ms (standalone: 41) when the local-energy kernel sits in between -- and the local-energy kernel pays
            # 20-30 ms instead, so the step time is the same within the box-to-box spread.  Kept as an option.
            if callable(local_energy):
                ee = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                ee[0].record()
                local_energy = local_energy()
                ee[1].record()
                self._eloc_events = ee
            self.last_local_energy = local_energy
            e = local_energy.to(device=dev, dtype=torch.complex128) if torch.is_tensor(local_energy) else \
                torch.as_tensor(np.asarray(local_energy, np.complex128)).to(dev)
            esum = torch.view_as_real(e.sum().reshape(1)).clone()
            if world > 1:
                dist.all_reduce(esum)
            e = e - torch.view_as_complex(esum) / B
            if world > 1:
                parts = [torch.empty_like(e) for _ in range(world)]
                dist.all_gather(parts, e)
                rhs = torch.cat([torch.cat([p.real, p.imag]) for p in parts]) / B      # row order of the blocks: [Re ; Im] per rank
            else:
                rhs = torch.cat([e.real, e.imag]) / B
            centre()
        rhs0 = rhs.contiguous()
        ev[4].record()
        nres = self.refinements + 2

        def solve_local(kind, factored):
            """S w = rhs0 on this rank -> (w fp64 [R], residual history or None)"""
            rhs = rhs0.clone()
            mixed = kind == 'mixed'
            sws = solve_workspace(mixed)
            resid = torch.zeros(nres, dtype=torch.float64, device=dev) if mixed else None
            with torch.cuda.device(dev):
                if mixed and factored:
                    _lib.check(lib.fk_sr_solve_factored(solver, _ptr(S), _ptr(rhs), R, self.refinements, _ptr(resid), _ptr(sws),
                                                        sws.numel(), stream))
                elif mixed:
                    _lib.check(lib.fk_sr_solve_mixed(solver, _ptr(S), _ptr(rhs), R, self.refinements, _ptr(info), _ptr(resid),
                                                     _ptr(sws), sws.numel(), stream))
                else:   # (overwrites S with its factor)
                    _lib.check(lib.fk_sr_solve(solver, _ptr(S), _ptr(rhs), R, _ptr(info), _ptr(sws), sws.numel(), stream))
            return rhs, resid

        def finish(kind):
            """solve S w = rhs with the given solver, then delta = X^T (C w) (gathered over the ranks)"""
            mixed = kind == 'mixed'
            if split:
                # the solver rank solves; w, the residual history and the potrf status travel in one broadcast
                buf = torch.zeros(R + nres + 1, dtype=torch.float64, device=dev)
                if is_solver:
                    w64, resid = solve_local(kind, factored=True)
                    buf[:R] = w64
                    if resid is not None:
                        buf[R:R + nres] = resid
                    buf[R + nres] = info[0].to(torch.float64)
                dist.broadcast(buf, src=self.solver_rank)
                rhs = buf[:R]
                resid = buf[R:R + nres] if mixed else None
                info.copy_(buf[R + nres:].round().to(torch.int32))
            else:
                rhs, resid = solve_local(kind, factored=False)
            if kind == self.solver:
                ev[5].record()
            # centre w per half, one pass over the parameter slice, gather the slices
            w = rhs.view(world, 2, Bl)
            w = (w - w.mean(dim=(0, 2), keepdim=True)).reshape(-1).float().contiguous()
            Kpad = nkb_r * 64
            d_loc = torch.zeros(Kpad, dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                if Kg > 0:
                    _lib.check(lib.fk_sr_xt_w(_ptr(Xg), R, Kg, Rl, world, block_stride, _ptr(w), _ptr(d_loc), stream))
            if world > 1:
                parts = [torch.empty_like(d_loc) for _ in range(world)]
                dist.all_gather(parts, d_loc)
                out = torch.cat(parts)[:P]
            else:
                out = d_loc[:P]
            self._resid = resid
            return out

        delta = finish(self.solver)
        self._refinish = lambda: finish('fp64')      # (S is still intact after the mixed-precision solve)
        ev[6].record()
        self._events = ev
        self._info = info
        return delta

    def read_timings(self):
        """after a synchronise: phase times of the last delta() in ms (+ the potrf status)"""
        ev = self._events
        names = ['jacobian', 'exchange', 'gram', 'centre', 'cholesky', 'update']
        self.timings_ms = {n: ev[i].elapsed_time(ev[i + 1]) for i, n in enumerate(names)}
        if self._eloc_events is not None:      # the local energies were evaluated inside the 'centre' interval
            self.timings_ms['eloc'] = self._eloc_events[0].elapsed_time(self._eloc_events[1])
            self.timings_ms['centre'] -= self.timings_ms['eloc']
        if self._split_events is not None:     # split solve: this rank's share of the local energies, the factorisation (solver
            fe = self._split_events['factor']  # rank), and in 'centre' what is left: sample gather, centring, waiting for the others
            self.timings_ms.setdefault('eloc', 0.0)
            self.timings_ms['factor'] = fe[0].elapsed_time(fe[1]) if fe is not None else 0.0
            self.timings_ms['centre'] -= self.timings_ms['factor']
            self.timings_ms['eloc_samples'] = self._split_events['n']
        self.timings_ms['solve'] = sum(self.timings_ms[n] for n in names[1:])
        self.potrf_info = int(self._info.item())
        self.needs_fp64_solve = False
        if self._resid is not None:
            r2 = self._resid.cpu().numpy()
            self.refinement_residuals = list(np.sqrt(r2[1:] / r2[0])) if r2[0] > 0 else [0.0]
            # the fp32 factorisation failed (matrix not positive definite in fp32) or the refinement did not contract far enough:
            # the caller re-solves in fp64 (refine_with_fp64), identically on every rank (S and the residuals are replicated)
            if self.potrf_info != 0 or not self.refinement_residuals[-1] < self.refinement_tol:
                self.needs_fp64_solve = True
        return self.timings_ms

    def refine_with_fp64(self):
        """delta again with the fp64 Cholesky (S, the right-hand side and X of the last delta() are still in place)"""
        delta = self._refinish()
        self.torch.cuda.synchronize()
        self.potrf_info = int(self._info.item())
        self.fp64_fallbacks = getattr(self, 'fp64_fallbacks', 0) + 1
        return delta
