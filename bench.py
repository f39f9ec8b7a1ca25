#!/usr/bin/env python
"""Headline benchmark: one VMC step (exact autoregressive sampling -> local energy -> weighted gradient + Adam)
of Heisenberg 2-D 10x10 OBC, ConvNetAutoregressive2D(depth 20, 32 channels), synthetic random-init weights.

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
def cpu_step_rate(sample_batch, repeats, seed=0):
    """samples/s of one VMC step (incremental sampling + E_loc + weighted gradient) on the host cores."""
    import torch
    from oracle import nets, operators as oops, local_energy as oeloc, sampler as osampler
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    spec = nets.Conv2DSpec(H, W, DEPTH, CHANNELS)
    params = nets.init_params(spec, seed=seed, dtype=torch.float32)
    op = oops.OracleOperator('heisenberg', (H, W), pbc=False)
    inc = osampler.IncrementalSampler2D(spec, params)
    rng = np.random.default_rng(seed)
    times, n_conn = [], 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        sigma, _ = inc.sample(rng.random((sample_batch, H, W)))
        t1 = time.perf_counter()
        lv = oeloc.local_values(op, lambda c: nets.log_psi_numpy(spec, params, c, batch_size=256), sigma.astype(np.float64))
        t2 = time.perf_counter()
        y = oeloc.loss_coefficients(lv, lv.mean(), sample_batch)
        nets.weighted_gradient(spec, params, sigma, y)
        t3 = time.perf_counter()
        times.append((t1 - t0, t2 - t1, t3 - t2))
    t = np.array(times)
    best = t.sum(axis=1).min()
    return {'value': sample_batch / best, 'sampling_samples_per_s': sample_batch / t[:, 0].min(),
            'eloc_evals_per_s': sample_batch / t[:, 1].min(), 'grad_samples_per_s': sample_batch / t[:, 2].min(),
            'cores': cores, 'ms_per_step': best * 1e3}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sample_batch = args.cpu_batch
    t_all = []
    res = None
    for i in range(args.warmup + args.steps):
        res = cpu_step_rate(sample_batch, 1, seed=i)
        if i >= args.warmup:
            t_all.append(res['ms_per_step'])
    ms = float(np.mean(t_all))
    value = sample_batch / (ms * 1e-3)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'samples/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, 1),
        'cpu_baseline': {'value': value, 'unit': 'samples/s', 'cores': res['cores'], 'kind': 'port',
                         'sample': '%d samples per step (of the %d-sample workload): incremental sampling + E_loc over '
                                   'all connections + weighted gradient, torch-CPU fp32 oracle' % (sample_batch, args.batch_per_gpu)},
        'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    emit(line)


METRIC = 'vmc_step_samples_per_sec (sample + E_loc + gradient), Heisenberg 2D 10x10 OBC ConvNetAutoregressive2D d20 c32'


def workload_config(args, world):
    return {'workload': 'Heisenberg 2-D 10x10 OBC, ConvNetAutoregressive2D depth 20 / 32 channels, fast sampling, '
                        'batch %d per GPU (BASELINE.json configs[2])' % args.batch_per_gpu,
            'lattice': [H, W], 'depth': DEPTH, 'channels': CHANNELS, 'global_batch': args.batch_per_gpu * world,
            'batch_per_gpu': args.batch_per_gpu, 'engine': args.engine,
            'precision': ('tcgen05: fp16 operands, fp32 accumulation (sampling, E_loc, gradient with power-of-two loss scaling)'
                          if args.engine == 'tc' else 'fp32 CUDA cores'),
            'parallelism': 'samples sharded over %d GPU(s); allreduce of energy statistics and flat gradient' % world,
            'l2_policy': 'per-step working set (activation workspaces, several GB) is much larger than the 126 MB L2'}


def sharded_sr_leg(args, world, rank, model, cond, machine, operator, obs, timed):
    """N > 1: the north-star step with stochastic reconfiguration -- sample + local energy + sample-space SR update -- on a
    GLOBAL batch of `--batch-per-gpu` samples sharded over the ranks (strong scaling for this leg: the 2B x 2B system is
    a function of the global batch; 8192 per GPU x 8 would be a 131 072^2 Gram).  Exchange per update: all-to-all of the
    bf16 Jacobian rows, allreduce of the partial Grams, two allreduces of P floats.  Every rank
    runs it; the time is the max over ranks.  Failures are reported in the JSON instead of taking the headline down."""
    import torch
    from flowket_b200.samplers import FastAutoregressiveSampler
    from flowket_b200.optimizers import StochasticReconfiguration
    try:
        torch.cuda.empty_cache()
        B_sr = max(1, args.batch_per_gpu // world)
        sampler_sr = FastAutoregressiveSampler(cond, B_sr, seed=4321, sample_offset=rank * B_sr)
        sr = StochasticReconfiguration(model, lr=0.01, diag_shift=0.05, sample_space=True, gram_dtype='bf16',
                                       jacobian_chunk=512, distributed=True)

        def step():
            sigma = sampler_sr.next_device()
            eloc = obs.local_values_device(model, sigma)
            sr.step(sigma, eloc)
            machine.device_net()

        step()                                   # warm-up (allocations, NCCL channels for the all-to-all)
        steps = 2
        ms = timed(step, steps, record=None) / steps
        torch.cuda.synchronize()
        res = {'ms_per_step': ms, 'samples_per_s': B_sr * world / (ms * 1e-3), 'global_batch': B_sr * world,
               'batch_per_gpu': B_sr, 'scaling': 'strong', 'sr_update_ms': dict(sr.last_timings_ms),
               'solver': 'sample space, batch sharded over %d ranks: all-to-all re-shard of the bf16 Jacobian rows to '
                         'parameter-major, partial 2B x 2B Gram per rank, fp32 allreduce, replicated fp64 Cholesky' % world,
               'peak_mem_gb': torch.cuda.max_memory_allocated() / 1e9}
        del sr
        torch.cuda.empty_cache()
        return res
    except Exception as exc:
        return {'error': '%s: %s' % (type(exc).__name__, exc)}


# ----------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from flowket_b200 import Input, Model, _lib, FK_ENGINE_FP32, FK_ENGINE_TC
    from flowket_b200.machines import ConvNetAutoregressive2D
    from flowket_b200.operators import Heisenberg
    from flowket_b200.samplers import FastAutoregressiveSampler
    from flowket_b200.optimization import VariationalMonteCarlo, DistributedVariationalMonteCarlo
    from flowket_b200.observables.monte_carlo import Observable
    from flowket_b200.optimizers import Adam, Trainer, allreduce_sum_

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    lib = _lib.require_cuda()
    B = args.batch_per_gpu
    engine = FK_ENGINE_TC if args.engine == 'tc' else FK_ENGINE_FP32

    inp = Input(shape=(H, W), dtype='int8')
    machine = ConvNetAutoregressive2D(inp, depth=DEPTH, num_of_channels=CHANNELS, seed=0)
    model = Model(inputs=inp, outputs=machine.predictions)
    model.engine = engine
    cond = Model(inputs=inp, outputs=machine.conditional_log_probs)
    cond.engine = engine          # the sampler follows the engine of the conditional-log-probs model
    net = machine.device_net()
    if world > 1:   # rank-0 broadcast of the initial variables (BroadcastGlobalVariablesCallback(0))
        dist.broadcast(machine.flat_params_device(), src=0)
        machine.params_updated()
    sampler = FastAutoregressiveSampler(cond, B, seed=1234, sample_offset=rank * B)
    operator = Heisenberg(hilbert_state_shape=[H, W], pbc=False)
    obs = Observable(operator)
    opt = Adam(lr=1e-3, beta_1=0.9, beta_2=0.9)
    phase_ms = {'sample': [], 'eloc': [], 'grad': []}
    conn_count = [0]

    def device_step(record):
        """inputs and outputs stay in HBM"""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        sigma = sampler.next_device()
        ev[1].record()
        eloc = obs.local_values_device(model, sigma)
        ev[2].record()
        stats = obs.last_stats.clone()
        if world > 1:
            dist.all_reduce(stats)
        mean = torch.complex(stats[0], stats[1]) / stats[3]
        y = (torch.conj(eloc - mean) / (B * world)).to(torch.complex64)
        grad = net.grad_weighted(net.to_sigma(sigma), y, engine=engine) / float(B)
        if world > 1:
            dist.all_reduce(grad)
        opt.step(machine.flat_params_device(), grad)
        machine.params_updated()
        machine.device_net()          # re-derive the effective (weight-normalised) kernels
        ev[3].record()
        if record:
            torch.cuda.synchronize()
            phase_ms['sample'].append(ev[0].elapsed_time(ev[1]))
            phase_ms['eloc'].append(ev[1].elapsed_time(ev[2]))
            phase_ms['grad'].append(ev[2].elapsed_time(ev[3]))
            conn_count[0] = obs.last_num_connections
        return mean

    vmc_cls = DistributedVariationalMonteCarlo if world > 1 else VariationalMonteCarlo
    vmc = vmc_cls(model, operator, sampler)
    trainer = Trainer(model, vmc, opt, distributed=world > 1)
    h2d = [0]
    d2h = [0]

    def e2e_step():
        """public API, host buffers: next_batch() returns host ndarrays, train_on_batch takes host ndarrays"""
        x, y = next(vmc)            # D2H: sigma (int8) + E_loc (complex128)
        g = trainer.gradient(x, y)  # H2D: sigma (int8) + y (complex64)
        if world > 1:
            allreduce_sum_(g)
        opt.step(machine.flat_params_device(), g)
        machine.params_updated()
        machine.device_net()
        e = complex(vmc.current_energy)   # D2H read of the step's result
        h2d[0] = x.nbytes + np.asarray(y, np.complex64).nbytes
        d2h[0] = x.nbytes + vmc.current_local_energy.nbytes
        return e

    def timed(fn, steps, record=False):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(steps):
            fn(record) if record is not None else fn()
        t1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        device_step(False)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = lib.fk_launch_count()
    total_ms = timed(device_step, args.steps, record=False)
    launches = lib.fk_launch_count() - launches0
    clk = clocks.stop() if rank == 0 else None
    # per-phase device times (separate pass so the per-phase synchronisation does not pollute `value`)
    for _ in range(min(args.steps, 2)):
        device_step(True)
    # end-to-end through the public API with host buffers
    for _ in range(2):
        e2e_step()
    e2e_ms = timed(lambda: e2e_step(), args.steps, record=None)

    sr_sharded = None
    if world > 1 and not args.no_sr:
        sr_sharded = sharded_sr_leg(args, world, rank, model, cond, machine, operator, obs, timed)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    ms_per_step = total_ms / args.steps
    value = B * world / (ms_per_step * 1e-3)
    n_conn = conn_count[0]                       # sum_b (1 + n_conn_b) on this rank
    eloc_ms = float(np.min(phase_ms['eloc']))
    flops_eloc = n_conn * F_FWD                   # algorithmic: (1 + n_conn) * F_fwd per sample, SURVEY 8(d)
    achieved_tf = flops_eloc / (eloc_ms * 1e-3) / 1e12
    line = {
        'metric': METRIC, 'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f16' if args.engine == 'tc' else 'f32', 'data': 'synthetic',
        'config': workload_config(args, world),
        'phases_ms': {k: float(np.min(v)) for k, v in phase_ms.items()},
        'sampling_samples_per_s': B * world / (float(np.min(phase_ms['sample'])) * 1e-3),
        'eloc_evals_per_s': B * world / (eloc_ms * 1e-3),
        'psi_evals_per_s': n_conn * world / (eloc_ms * 1e-3),
        'connections_per_sample': n_conn / float(B),
        'e2e': {'value': B * world / (e2e_ms / args.steps * 1e-3), 'unit': 'samples/s', 'h2d_bytes_per_step': int(h2d[0]),
                'd2h_bytes_per_step': int(d2h[0])},
        'gpu_launches': int(launches),
        'clocks': clk,
        'roofline': {'bound': 'tensor', 'achieved': achieved_tf, 'peak': pk['bf16_sustained'] or pk['bf16'], 'unit': 'TFLOP/s',
                     'frac': achieved_tf / (pk['bf16_sustained'] or pk['bf16']),
                     # dram__bytes_read + dram__bytes_write of one tc_forward_kernel launch (65 536 configurations), ncu capture
                     # profiles/r01_ncu_tc_forward_v10.txt: the kernel lives in shared memory / TMEM, HBM sees the weights once
                     'traffic': 8863232 if args.engine == 'tc' else None,
                     'kernel': 'local-energy wave-function evaluations (%s engine)' % args.engine,
                     'peak_source': pk['source'] + ' bf16 sustained (kernel timed inside a long step)',
                     'flops_per_launch': flops_eloc,
                     # measured: with 32 output channels per MMA the forward kernel is bounded by the
                     # shared-memory operand stream (128 B/cycle/SM), not by tensor math; ~300 KB per configuration and block
                     'smem_roofline': {'bytes_per_cfg_block': 300e3, 'peak_bytes_per_cycle_per_sm': 128,
                                       'frac': ((n_conn / (eloc_ms * 1e-3)) * 38 * 300e3 / (148 * 128 * 1.965e9)) if args.engine == 'tc' else None}},
    }
    if world == 1 and not args.no_sr:
        # The same step with stochastic reconfiguration instead of Adam (north_star: sample + local energy + SR): per-sample
        # Jacobians of the 854 k parameters, sample-space Gram (2B x 2B, K = P) as one tensor-core GEMM, Cholesky, update.
        try:
            from flowket_b200.optimizers import StochasticReconfiguration
            sr = StochasticReconfiguration(model, lr=0.01, diag_shift=0.05, sample_space=True, gram_dtype='bf16',
                                           jacobian_chunk=512)
            times = []
            for i in range(3):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                sigma = sampler.next_device()
                eloc = obs.local_values_device(model, sigma)
                sr.step(sigma, eloc)
                machine.device_net()
                e1.record()
                torch.cuda.synchronize()
                if i > 0:
                    times.append(e0.elapsed_time(e1))
            ms = float(np.mean(times))
            line['sr_step'] = {'ms_per_step': ms, 'samples_per_s': B / (ms * 1e-3), 'sr_update_ms': dict(sr.last_timings_ms),
                               'solver': 'sample space: delta = X^T (X X^T / B + lambda)^-1 e / B, X = [Re O; Im O] (2B x P), '
                                         'bf16 Gram / fp32 accumulate, fp64 Cholesky',
                               'peak_mem_gb': torch.cuda.max_memory_allocated() / 1e9}
            del sr
            torch.cuda.empty_cache()
        except Exception as exc:  # the SR leg must never take the headline line down with it
            line['sr_step'] = {'error': '%s: %s' % (type(exc).__name__, exc)}
    if sr_sharded is not None:
        line['sr_step'] = sr_sharded
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_step_rate(args.cpu_batch, 2)
        line['cpu_baseline'] = {'value': cb['value'], 'unit': 'samples/s', 'cores': cb['cores'], 'kind': 'port',
                                'sampling_samples_per_s': cb['sampling_samples_per_s'], 'eloc_evals_per_s': cb['eloc_evals_per_s'],
                                'sample': '%d samples of the same workload, best of 2: incremental sampling + E_loc over all '
                                          'connections + weighted gradient, torch-CPU fp32 oracle port' % args.cpu_batch}
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
    ap.add_argument('--engine', default=os.environ.get('FK_BENCH_ENGINE', 'tc'), choices=['fp32', 'tc'],
                    help='tc: tcgen05 engines for sampling, E_loc and the gradient (fp16 operands / fp32 accumulate); '
                         'fp32: CUDA-core exact engines everywhere')
    ap.add_argument('--batch-per-gpu', type=int, default=8192)
    ap.add_argument('--cpu-batch', type=int, default=8)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-sr', action='store_true', help='skip the stochastic-reconfiguration variant of the step')
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
