#!/usr/bin/env python
"""Headline benchmark: one VMC step of the north-star target -- exact autoregressive sampling -> local energy ->
stochastic-reconfiguration update -- of Heisenberg 2-D 10x10 OBC, ConvNetAutoregressive2D(depth 20, 32 channels),
global batch 8192 sharded over the GPUs (strong scaling), synthetic random-init weights.

  python bench.py --gpus N --steps K --warmup W            # this repository (CUDA, one process per GPU)
  python bench.py --impl reference --gpus N ...            # reference-equivalent CPU path (oracle port) on the host cores

Prints ONE JSON line (rank 0).  `value` = whole-job samples/s with inputs resident in HBM; `e2e` = the same step
through the public FlowKet-shaped API with host buffers (H2D/D2H inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W, DEPTH, CHANNELS = 10, 10, 20, 32
MAC_PER_SITE = 844288            # k=3, C=32, depth=20
F_FWD = 2.0 * MAC_PER_SITE * H * W   # 168.86 MFLOP per full forward


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p['hbm_gbs'], bf16=p['bf16_tflops'], bf16_sustained=p.get('bf16_tflops_sustained'), source='measured')
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source='fallback')


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu_index, self.lines, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu_index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (reference-equivalent CPU path; TensorFlow is not installable)
# ----------------------------------------------------------------------------------------------------------
def cpu_step(sample_batch, seed=0, diag_shift=0.05, weights=None):
    """One north-star step on the host cores: incremental sampling + E_loc over all connections + stochastic
    reconfiguration (per-sample Jacobians by autograd, sample-space solve in fp64).  Returns the timings and the
    samples / local energies (the GPU arm checks its engines against them)."""
    import torch
    from oracle import nets, operators as oops, local_energy as oeloc, sampler as osampler
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    spec = nets.Conv2DSpec(H, W, DEPTH, CHANNELS)
    if weights is None:
        params = nets.init_params(spec, seed=0, dtype=torch.float32)
    else:                                                           # the GPU arm's machine, same bytes
        params = [torch.from_numpy(np.asarray(w, np.float32)) for w in weights]
    op = oops.OracleOperator('heisenberg', (H, W), pbc=False)
    inc = osampler.IncrementalSampler2D(spec, params)
    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    sigma, _ = inc.sample(rng.random((sample_batch, H, W)))
    t1 = time.perf_counter()
    lv = oeloc.local_values(op, lambda c: nets.log_psi_numpy(spec, params, c, batch_size=256), sigma.astype(np.float64))
    t2 = time.perf_counter()
    O_re = nets.per_sample_gradients(spec, params, sigma, 'real').double().numpy()
    O_im = nets.per_sample_gradients(spec, params, sigma, 'imag').double().numpy()
    X = np.concatenate([O_re - O_re.mean(0), O_im - O_im.mean(0)])
    e = np.asarray(lv).reshape(-1) - np.mean(lv)
    ep = np.concatenate([e.real, e.imag])
    T = X @ X.T / sample_batch + diag_shift * np.eye(2 * sample_batch)
    delta = X.T @ np.linalg.solve(T, ep / sample_batch)
    t3 = time.perf_counter()
    return {'sample_s': t1 - t0, 'eloc_s': t2 - t1, 'sr_s': t3 - t2, 'total_s': t3 - t0, 'cores': cores,
            'sigma': sigma, 'local_values': np.asarray(lv).reshape(-1), 'delta_norm': float(np.linalg.norm(delta))}


def cpu_step_rate(sample_batch, repeats, weights=None):
    runs = [cpu_step(sample_batch, seed=i, weights=weights) for i in range(repeats)]
    best = min(runs, key=lambda r: r['total_s'])
    return {'value': sample_batch / best['total_s'], 'sampling_samples_per_s': sample_batch / min(r['sample_s'] for r in runs),
            'eloc_evals_per_s': sample_batch / min(r['eloc_s'] for r in runs),
            'sr_samples_per_s': sample_batch / min(r['sr_s'] for r in runs), 'cores': best['cores'],
            'ms_per_step': best['total_s'] * 1e3, 'last': runs[-1]}


CPU_SAMPLE = ('%d samples per step of the same workload (the step is linear in the batch up to the 2B x 2B solve): incremental '
              'sampling + E_loc over all connections + per-sample Jacobians and sample-space SR solve, torch-CPU fp32 oracle port '
              '(fp64 solve), all host threads')


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sample_batch = args.cpu_batch
    t_all, res = [], None
    for i in range(args.warmup + args.steps):
        res = cpu_step(sample_batch, seed=i)
        if i >= args.warmup:
            t_all.append(res['total_s'] * 1e3)
    ms = float(np.mean(t_all))
    value = sample_batch / (ms * 1e-3)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'samples/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, 1),
        'cpu_baseline': {'value': value, 'unit': 'samples/s', 'cores': res['cores'], 'kind': 'port',
                         'sample': CPU_SAMPLE % sample_batch},
        'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    emit(line)


METRIC = ('vmc_sr_step_samples_per_sec (exact autoregressive sampling + local energy + stochastic-reconfiguration update), '
          'Heisenberg 2D 10x10 OBC ConvNetAutoregressive2D d20 c32, global batch 8192')


SPLIT_SOLVE_DEFAULT = True      # validated against the plain sharded step on 2 and 4 GPUs (tests/tools_split_solve.py, profiles/r02_split_solve_probe_2gpu.txt)


def workload_config(args, world):
    return {'workload': 'Heisenberg 2-D 10x10 OBC, ConvNetAutoregressive2D depth 20 / 32 channels, fast sampling, global '
                        'batch %d sharded over the GPUs (BASELINE.json configs[2]: "batch 8192 on 8xB200")' % args.global_batch,
            'lattice': [H, W], 'depth': DEPTH, 'channels': CHANNELS, 'global_batch': args.global_batch,
            'batch_per_gpu': args.global_batch // world,
            'step': 'sample + E_loc + SR update (diag_shift 0.05, lr 0.01, sample-space form of optimizer.py:55-108)',
            'engines': {'sampler': 'tcgen05 fp16 operands / fp32 accumulate (fk_sample_tc)',
                        'local_energy': 'value: tc-exact (fp16 hi+lo split operands, 22 bits, fp32 accumulate); value_fast: fp16 operands',
                        'jacobian': 'tcgen05 fp16 operands / fp32 accumulate, rows stored in bf16',
                        'gram': 'hand-written cta_group::2 tcgen05 GEMM, bf16 operands, fp32 accumulate',
                        'solve': 'fp32 Cholesky factor (cuSOLVER behind fk_sr_factor_mixed) + fp64 iterative refinement; automatic fp64 re-solve'},
            'parallelism': 'samples sharded over %d GPU(s); all-to-all of the bf16 Jacobian rows, fp32 allreduce of the partial '
                           'Gram matrices, allgather of the update slices' % world +
                           ('' if world == 1 or args.no_split_solve else
                            '; split solve: rank 0 factors the SR matrix while the other ranks evaluate local energies (the samples '
                            'are gathered and dealt so that all ranks finish together), one allreduce of the local energies, one '
                            'broadcast of the solution'),
            'l2_policy': 'per-step working set (28 GB of Jacobian rows, 1 GB Gram) is much larger than the 126 MB L2'}


# ----------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from flowket_b200 import Input, Model, _lib, FK_ENGINE_FP32, FK_ENGINE_TC, FK_ENGINE_TC_EXACT
    from flowket_b200.machines import ConvNetAutoregressive2D
    from flowket_b200.operators import Heisenberg
    from flowket_b200.samplers import FastAutoregressiveSampler
    from flowket_b200.optimization import VariationalMonteCarlo, DistributedVariationalMonteCarlo
    from flowket_b200.observables.monte_carlo import Observable
    from flowket_b200.optimizers import Adam, StochasticReconfiguration

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    lib = _lib.require_cuda()
    GB = args.global_batch
    assert GB % world == 0, 'the global batch must divide over the ranks'
    B = GB // world

    inp = Input(shape=(H, W), dtype='int8')
    machine = ConvNetAutoregressive2D(inp, depth=DEPTH, num_of_channels=CHANNELS, seed=0)
    model = Model(inputs=inp, outputs=machine.predictions)
    cond = Model(inputs=inp, outputs=machine.conditional_log_probs)
    cond.engine = FK_ENGINE_TC          # the sampler follows the engine of the conditional-log-probs model
    net = machine.device_net()
    params0 = machine.flat_params_device().clone()
    if world > 1:   # rank-0 broadcast of the initial variables (BroadcastGlobalVariablesCallback(0))
        dist.broadcast(params0, src=0)
    sampler = FastAutoregressiveSampler(cond, B, seed=1234, sample_offset=rank * B)
    operator = Heisenberg(hilbert_state_shape=[H, W], pbc=False)
    obs = Observable(operator)
    obs.count_connections = False        # nothing on the timed path reads a count back from the device
    sr = StochasticReconfiguration(model, lr=0.01, diag_shift=0.05, sample_space=True, distributed=world > 1,
                                   read_timings=False)

    def reset():
        machine.flat_params_device().copy_(params0)
        machine.params_updated()
        machine.device_net()

    # world > 1: split solve -- the optimizer gets the local-energy FUNCTION; after the Gram allreduce one rank factors the
    # (local-energy independent) matrix while the others evaluate local energies, the samples dealt so that all finish together
    # (flowket_b200/optimizers/sample_space_sr.py); --no-split-solve: every rank evaluates its own samples and factors
    split = world > 1 and not args.no_split_solve

    def sr_step(engine):
        """sample -> local energy -> SR update; inputs and outputs stay in HBM"""
        model.engine = engine
        sigma = sampler.next_device()
        eloc = obs.per_sample_device(model) if split else obs.local_values_device(model, sigma)
        sr.step(sigma, eloc)
        machine.device_net()              # re-derive the effective (weight-normalised) kernels, repack the operand images

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(steps):
            fn()
        t1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    warm = max(args.warmup, 3)
    # ---- headline: tc-exact local energy
    reset()
    for _ in range(warm):
        sr_step(FK_ENGINE_TC_EXACT)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = lib.fk_launch_count()
    total_ms = timed(lambda: sr_step(FK_ENGINE_TC_EXACT), args.steps)
    launches = lib.fk_launch_count() - launches0
    clk = clocks.stop() if rank == 0 else None
    # ---- the same step with the fp16 local-energy engine
    reset()
    for _ in range(2):
        sr_step(FK_ENGINE_TC)
    fast_ms = timed(lambda: sr_step(FK_ENGINE_TC), args.steps)
    # ---- per-phase device times (separate pass: the per-phase synchronisation must not pollute `value`)
    reset()
    sr.read_timings = True
    obs.count_connections = True
    phases = {}
    for engine, tag in ((FK_ENGINE_TC_EXACT, 'exact'), (FK_ENGINE_TC, 'fast')):
        model.engine = engine
        best = None
        for _ in range(2):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            sigma = sampler.next_device()
            ev[1].record()
            eloc = obs.per_sample_device(model) if split else obs.local_values_device(model, sigma)
            ev[2].record()
            sr.step(sigma, eloc)
            machine.device_net()
            ev[3].record()
            torch.cuda.synchronize()
            cur = {'sample': ev[0].elapsed_time(ev[1]), 'eloc': ev[1].elapsed_time(ev[2]), 'sr_total': ev[2].elapsed_time(ev[3])}
            tm = dict(sr.last_timings_ms)
            if split:      # rank 0's view: its share of the local energies and the factorisation sit inside the update
                cur['eloc'] = tm.pop('eloc', 0.0)
                cur['sr_total'] -= cur['eloc']
                cur['eloc_samples_rank0'] = tm.pop('eloc_samples', 0)
            cur.update({'sr_' + k: v for k, v in tm.items() if k != 'solve'})
            if best is None or cur['eloc'] + cur['sr_total'] < best['eloc'] + best['sr_total']:
                best = cur
        phases[tag] = best
    n_conn = obs.last_num_connections           # sum_b (1 + n_conn_b) on this rank
    sr.read_timings = False
    obs.count_connections = False

    # ---- end to end through the public FlowKet-shaped API with host buffers
    reset()
    model.engine = FK_ENGINE_TC_EXACT
    vmc_cls = DistributedVariationalMonteCarlo if world > 1 else VariationalMonteCarlo
    vmc = vmc_cls(model, operator, sampler)
    h2d, d2h = [0], [0]

    def e2e_step():
        if split:
            sr.step_generator(vmc)                       # D2H: sigma; H2D: sigma; D2H: this rank's E_loc + the energy
            x = vmc.current_batch
            h2d[0] = x.nbytes
        else:
            x, y = vmc.next_batch()                      # D2H: sigma (int8 host ndarray) + E_loc (complex128 host ndarray)
            sr.step(x, vmc.current_local_energy)          # H2D: sigma + local energies as host ndarrays
            h2d[0] = x.nbytes + vmc.current_local_energy.nbytes
        machine.device_net()
        e = complex(vmc.current_energy)               # the step's result on the host
        d2h[0] = x.nbytes + vmc.current_local_energy.nbytes + 16
        return e

    for _ in range(2):
        e2e_step()
    e2e_ms = timed(e2e_step, args.steps)

    # ---- secondary: the Adam step of round 1 (sample + E_loc + weighted gradient + Adam), weak scaling, fp16 engines
    adam = None
    if not args.no_adam:
        reset()
        model.engine = FK_ENGINE_TC
        Bw = args.batch_per_gpu
        sampler_w = FastAutoregressiveSampler(cond, Bw, seed=99, sample_offset=rank * Bw)
        opt = Adam(lr=1e-3, beta_1=0.9, beta_2=0.9)

        def adam_step():
            sigma = sampler_w.next_device()
            eloc = obs.local_values_device(model, sigma)
            stats = obs.last_stats.clone()
            if world > 1:
                dist.all_reduce(stats)
            mean = torch.complex(stats[0], stats[1]) / stats[3]
            y = (torch.conj(eloc - mean) / (Bw * world)).to(torch.complex64)
            grad = net.grad_weighted(net.to_sigma(sigma), y, engine=FK_ENGINE_TC) / float(Bw)
            if world > 1:
                dist.all_reduce(grad)
            opt.step(machine.flat_params_device(), grad)
            machine.params_updated()
            machine.device_net()

        for _ in range(2):
            adam_step()
        adam_ms = timed(adam_step, 2) / 2
        adam = {'ms_per_step': adam_ms, 'samples_per_s': Bw * world / (adam_ms * 1e-3), 'batch_per_gpu': Bw, 'scaling': 'weak',
                'engines': 'fp16 tensor-core sampler, local energy and gradient; Adam(beta 0.9, 0.9)'}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    peak_tf = pk['bf16_sustained'] or pk['bf16']
    ms_per_step = total_ms / args.steps
    value = GB / (ms_per_step * 1e-3)
    fast_step = fast_ms / args.steps
    eloc_exact_ms, eloc_fast_ms = phases['exact']['eloc'], phases['fast']['eloc']
    # samples whose local energies rank 0 evaluated in the phase pass (split solve: its share of the global batch, which may
    # be empty when the factorisation alone takes as long as the other ranks' local energies)
    n_exact = phases['exact'].get('eloc_samples_rank0', B)
    n_fast = phases['fast'].get('eloc_samples_rank0', B)
    conn_per_sample = n_conn / float(n_fast) if n_fast else None        # (the last phase pass was the fp16 engine's)
    if conn_per_sample is None:
        conn_per_sample = 85.3                                            # measured at N = 1 on this workload
    flops_eloc = conn_per_sample * n_exact * F_FWD     # algorithmic: (1 + n_conn) * F_fwd per sample, SURVEY 8(d), this rank's share

    def rate(x, ms):
        return x / (ms * 1e-3) if ms and ms > 0 else None

    tf_exact = (rate(flops_eloc, eloc_exact_ms) or 0.0) / 1e12
    tf_fast = (rate(conn_per_sample * n_fast * F_FWD, eloc_fast_ms) or 0.0) / 1e12
    P = net.num_params
    Kg = (P + world - 1) // world
    nt = (2 * GB + 255) // 256
    gram_flops = 2.0 * (nt * (nt + 1) / 2) * 65536 * Kg          # computed upper block triangle, this rank's parameter slice
    gram_ms = phases['exact'].get('sr_gram')
    line = {
        'metric': METRIC, 'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps, 'warmup': warm,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f16 hi+lo split (22-bit) forward, f16 Jacobian, bf16 Gram, f64 solve', 'data': 'synthetic',
        'config': workload_config(args, world),
        'value_fast': GB / (fast_step * 1e-3), 'ms_per_step_fast': fast_step,
        'phases_ms': phases,
        'sampling_samples_per_s': GB / (phases['exact']['sample'] * 1e-3),
        # (rank 0's rate x the number of GPUs)
        'eloc_evals_per_s': rate(n_exact * world, eloc_exact_ms), 'eloc_evals_per_s_fast': rate(n_fast * world, eloc_fast_ms),
        'psi_evals_per_s': rate(conn_per_sample * n_exact * world, eloc_exact_ms),
        'psi_evals_per_s_fast': rate(conn_per_sample * n_fast * world, eloc_fast_ms),
        'connections_per_sample': conn_per_sample,
        'split_solve': bool(split),
        'e2e': {'value': GB / (e2e_ms / args.steps * 1e-3), 'unit': 'samples/s', 'h2d_bytes_per_step': int(h2d[0]),
                'd2h_bytes_per_step': int(d2h[0])},
        'gpu_launches': int(launches),
        'clocks': clk,
        # dominant kernel of the step: the tc-exact wave-function evaluations of the local energy (tcx_forward_kernel)
        'roofline': {'bound': 'tensor', 'achieved': tf_exact, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': tf_exact / peak_tf,
                     # dram__bytes_read + dram__bytes_write of the work-list launch of this kernel at this batch, from the ncu
                     # capture profiles/r02_ncu_traffic_B8192.txt (spins, work list, matrix elements, weights once; activations never
                     # leave shared memory / TMEM).  Only quoted for the configuration the capture was taken on.
                     # dram__bytes_read + dram__bytes_write of the tile-pass launch of this kernel at this batch (halo rows read back from
                     # the samples' activation cache), from the ncu capture of the same launch; only quoted for that configuration
                     'traffic': (156.40e9 + 3.87e9) if (world == 1 and GB == 8192) else None,
                     'traffic_source': 'profiles/r02_ncu_traffic_prefix_B8192.txt (ncu capture of the same launches; without prefix reuse: 13 MB)',
                     'note': 'algorithmic FLOPs = one full forward per connected configuration (what the reference evaluates); with prefix '
                             'reuse the kernel issues the MMAs of 62 tiles per sample instead of 85 full evaluations, each product as '
                             'three fp16 tensor-core passes',
                     'kernel': 'tcx_forward_kernel (tile pass of the prefix reuse): local-energy wave-function evaluations, 3 tensor-core '
                               'products per MAC (hi*hi, hi*lo, lo*hi) counted as ONE algorithmic MAC',
                     'peak_source': pk['source'] + ' bf16 sustained (kernel timed inside a long step)',
                     'flops_per_launch': flops_eloc, 'launch_ms': eloc_exact_ms,
                     'fast_engine': {'achieved': tf_fast, 'frac': tf_fast / peak_tf, 'launch_ms': eloc_fast_ms,
                                     'kernel': 'tc_forward_kernel (fp16 operands, prefix reuse)'},
                     'gram': ({'achieved': gram_flops / (gram_ms * 1e-3) / 1e12, 'frac': gram_flops / (gram_ms * 1e-3) / 1e12 / peak_tf,
                               'launch_ms': gram_ms, 'flops_per_launch': gram_flops,
                               'kernel': 'gram2_kernel (cta_group::2 tcgen05, bf16), upper block triangle of the 2B x 2B Gram'}
                              if gram_ms else None)},
    }
    if adam is not None:
        line['adam_step_weak'] = adam
    if world == 1 and not args.no_cpu_baseline:
        reset()
        cb = cpu_step_rate(args.cpu_batch, args.cpu_repeats, weights=machine.get_weights())
        line['cpu_baseline'] = {'value': cb['value'], 'unit': 'samples/s', 'cores': cb['cores'], 'kind': 'port',
                                'sampling_samples_per_s': cb['sampling_samples_per_s'], 'eloc_evals_per_s': cb['eloc_evals_per_s'],
                                'sr_samples_per_s': cb['sr_samples_per_s'],
                                'sample': (CPU_SAMPLE % args.cpu_batch) + ', best of %d' % args.cpu_repeats}
        # accuracy of the device engines on the CPU leg's own samples: E_loc against the fp32 oracle's local values
        try:
            reset()
            last = cb['last']
            sg = net.to_sigma(last['sigma'])
            want = torch.as_tensor(last['local_values']).to(sg.device)
            acc = {}
            for engine, tag in ((FK_ENGINE_FP32, 'fp32_engine'), (FK_ENGINE_TC_EXACT, 'tc_exact'), (FK_ENGINE_TC, 'tc_fp16')):
                got, _, _ = net.local_energy(operator.device_desc(), sg, engine=engine, count=False)
                acc[tag] = float(((got - want).abs().max() / want.abs().max()).item())
            line['accuracy'] = {'what': 'max_b |E_loc(b) - oracle| / max_b |E_loc| on the %d samples of the CPU leg '
                                        '(oracle = torch-CPU fp32 restatement)' % args.cpu_batch, 'eloc': acc}
        except Exception as exc:
            line['accuracy'] = {'error': '%s: %s' % (type(exc).__name__, exc)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def claim_stdout():
    """stdout must carry exactly ONE JSON line, but native libraries write there too (NCCL prints its version banner on the
    first communicator).  Keep a private duplicate of fd 1 for the JSON line and point fd 1 at stderr for everything else."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--global-batch', type=int, default=8192, help='samples per step, sharded over the GPUs (strong scaling)')
    ap.add_argument('--batch-per-gpu', type=int, default=8192, help='batch of the secondary weak-scaling Adam step')
    ap.add_argument('--cpu-batch', type=int, default=64)
    ap.add_argument('--cpu-repeats', type=int, default=3)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-adam', action='store_true', help='skip the secondary weak-scaling Adam step')
    ap.add_argument('--no-split-solve', action='store_true', default=not SPLIT_SOLVE_DEFAULT,
                    help='N > 1: every rank evaluates the local energies of its own samples and factors the SR matrix itself')
    ap.add_argument('--split-solve', dest='no_split_solve', action='store_false',
                    help='N > 1: one rank factors the SR matrix while the others evaluate local energies')
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if not (world == 1 and args.gpus > 1 and args.impl != 'reference'):
        claim_stdout()      # (the torchrun re-launch below lets its children own the real stdout)
    if args.impl == 'reference':
        return run_reference(args)
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun (the driver launches torchrun itself)
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(args.gpus),
               '--master-addr', '127.0.0.1', '--master-port', '29531'] + sys.argv
        sys.exit(subprocess.call(cmd))
    run_gpu(args)


if __name__ == '__main__':
    main()
