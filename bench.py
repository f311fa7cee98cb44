#!/usr/bin/env python
"""Headline benchmark: batched NMPC steps/sec on the CSTR (nx=4, nu=2, N=50, Radau-3), B=4096 instances per GPU.

Contract: ``python bench.py --gpus N --steps K --warmup W`` (rank 0 prints ONE JSON line); under torchrun every rank
runs its own shard of the batch (weak scaling, no data-path collective — instances are independent), timing is
bracketed by barrier + synchronize and reduced with MAX over ranks.  ``--impl reference`` times the CPU restatement
of the reference's path (the real reference cannot run here: CasADi/IPOPT are not installable) on the
host cores, on a bounded sample of the same workload.

One "step" = one warm-started ``optimize()`` of every instance in the batch to IPOPT tolerance (1e-8), closed loop:
the next measured state is the predicted next state plus a seeded disturbance.
"""
from __future__ import annotations

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

N_HORIZON = 50
BATCH_PER_GPU = 4096
NX, NU = 4, 2
KERNEL_NAMES = "init,eval,riccati,expand,prep,trial,decide,apply,commit,count,step".split(',')
DIST_SCALE = np.array([.02, .02, .5, .5])  # closed-loop disturbance added to the predicted next state


def _peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d['hbm_gbs']), 'measured'
    return 6650.0, 'fallback'


class ClockSampler:
    """Samples SM clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).  NVML is queried
    in-process from a background thread (the ctypes solve call releases the GIL); an `nvidia-smi -lms` child process is
    the fallback.  `nvidia-smi -lms` next to the timed loop cost ~3 % of the step time on the shared boxes, which is why the in-process path is the default."""

    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, index=0, mode='nvml', interval_ms=50):
        self.index, self.mode, self.interval = index, mode, interval_ms * 1e-3
        self.proc = self.thread = None
        self.lines, self.sm, self.smax, self.reasons = [], [], [], set()
        self._stop = threading.Event()

    def start(self):
        if self.mode == 'none':
            return
        if self.mode == 'nvml':
            try:
                import pynvml
                pynvml.nvmlInit()
                phys = self.index
                vis = os.environ.get('CUDA_VISIBLE_DEVICES')
                if vis:
                    try:
                        phys = int(vis.split(',')[self.index])
                    except (ValueError, IndexError):
                        phys = self.index
                self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
                self._nv = pynvml
                self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
                self.thread.start()
                return
            except Exception:  # noqa: BLE001 - any NVML problem falls back to the CLI
                self.mode = 'smi'
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits',
                                          '-i', str(self.index), '-lms', '200'], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll_nvml(self):
        nv = self._nv
        bits = {'hw_slowdown': getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8),
                'hw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40),
                'sw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20),
                'sw_power_cap': getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4)}
        get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
            getattr(nv, 'nvmlDeviceGetCurrentClocksThrottleReasons')
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                self.smax.append(float(nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)))
                r = int(get_reasons(self._h))
                for nm, bit in bits.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(self.interval)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.mode == 'none':
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['clock sampling disabled'], 'samples': 0}
        if self.mode == 'nvml':
            self._stop.set()
            if self.thread is not None:
                self.thread.join(timeout=2)
            return {'sm_mhz': float(np.median(self.sm)) if self.sm else None,
                    'sm_max_mhz': max(self.smax) if self.smax else None, 'reasons': sorted(self.reasons),
                    'samples': len(self.sm), 'source': 'nvml'}
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(self.NAMES, parts[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'reasons': sorted(reasons), 'samples': len(sm), 'source': 'nvidia-smi'}


def cpu_port_sample(n_instances: int, n_procs: int, seed: int = 0):
    """Time the oracle (NumPy/SciPy restatement of the reference's CasADi+IPOPT path) on ``n_instances`` instances of
    the same workload with ``n_procs`` worker processes. Returns (steps_per_s, seconds)."""
    from hilo_mpc_b200.examples import cstr_initial_states
    x0 = cstr_initial_states(n_instances, seed=seed)
    if n_procs <= 1:
        t0 = time.perf_counter()
        _cpu_worker((x0, 0))
        el = time.perf_counter() - t0
    else:
        import multiprocessing as mp
        chunks = [(c, i) for i, c in enumerate(np.array_split(x0, n_procs)) if len(c)]
        with mp.get_context('fork').Pool(len(chunks)) as pool:
            pool.map(_cpu_warm, range(len(chunks)))  # build the symbolic stage functions once per worker (setup)
            t0 = time.perf_counter()
            pool.map(_cpu_worker, chunks)
            el = time.perf_counter() - t0
    return n_instances / el, el


_CPU_NLP = None


def _cpu_nlp():
    global _CPU_NLP
    if _CPU_NLP is None:
        from hilo_mpc_b200.examples import cstr_nmpc
        from oracle.nmpc import FlatNlp
        nmpc = cstr_nmpc(N=N_HORIZON, setup=False)
        nmpc._build_module = False  # the CPU arm needs only the problem description, not the CUDA module
        nmpc.setup()
        spec = nmpc.spec
        nlp = FlatNlp(spec, exact_hessian=True)
        xs = np.asarray(nmpc._xs_full)
        # trigger symbolic differentiation + NumPy codegen outside the timed region (the reference's setup() cost)
        nlp.solve(spec.x_guess.copy(), dt=nmpc.sampling_interval, max_iter=1)
        _CPU_NLP = (nlp, xs, nmpc.sampling_interval)
    return _CPU_NLP


def _cpu_warm(_):
    _cpu_nlp()
    return 0


def _cpu_worker(arg):
    x0s, _ = arg
    nlp, xs, dt = _cpu_nlp()
    ok = 0
    for x0 in x0s:
        res, _, _ = nlp.solve(x0 / xs, dt=dt)
        ok += int(res.status == 1)
    return ok


def gp_predict_line(dev, n: int = 512, d: int = 6, m: int = 65536, reps: int = 5):
    """Second BASELINE metric: GP predict microseconds per query point (SE kernel, n=512 training points, d=6 features,
    BASELINE config 3 sizes), posterior mean + variance and mean only, through ``GaussianProcess.predict`` with device
    tensors.  The variance is one (n x n) x (n x m) fp64 product on the DMMA pipe; flops = n^2 m (+ Gram n m (3d+4))."""
    import torch
    from hilo_mpc_b200.gp import GaussianProcess, Kernel
    rng = np.random.default_rng(1)
    X = rng.uniform(-1., 1., size=(d, n))
    y = np.sin(X.sum(0, keepdims=True)) + 1e-2 * rng.standard_normal((1, n))
    gp = GaussianProcess([f'x{i}' for i in range(d)], 'y', kernel=Kernel.squared_exponential(length_scales=.8),
                         noise_variance=1e-4)
    gp.set_training_data(X, y)
    t0 = time.perf_counter()
    gp.setup()
    torch.cuda.synchronize()
    fit_ms = (time.perf_counter() - t0) * 1e3
    Xq = torch.tensor(rng.uniform(-1., 1., size=(d, m)), dtype=torch.float64, device=dev)
    out = {'n_train': n, 'features': d, 'queries': m, 'fit_ms_incl_alloc': fit_ms}
    for key, mean_only in (('mean_var', False), ('mean_only', True)):
        for _ in range(2):
            gp.predict(Xq, mean_only=mean_only)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            gp.predict(Xq, mean_only=mean_only)
        e1.record()
        torch.cuda.synchronize()
        us_pt = e0.elapsed_time(e1) * 1e3 / (reps * m)
        flops_pt = n * (3 * d + 4) + 2 * n + (0 if mean_only else 2 * n * n)
        out[key] = {'us_per_pt': us_pt, 'gflops': flops_pt / us_pt * 1e-3}
    out['fp64_peak_note'] = 'B200 nominal fp64 (vector = tensor) 37-40 TFLOP/s; no measured fp64 peak in MEASURED_PEAKS.json'
    return out


def run_reference(args):
    """``--impl reference``: CPU arm, rank 0 only."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = max(cores, 8)
    _cpu_nlp()
    for _ in range(args.warmup):
        cpu_port_sample(min(per_step, cores), cores, seed=1)
    times = []
    for k in range(args.steps):
        _, el = cpu_port_sample(per_step, cores, seed=100 + k)
        times.append(el)
    total = float(np.sum(times))
    value = per_step * args.steps / total
    line = {
        'metric': 'batched NMPC steps/sec (CSTR, N=50)', 'value': value, 'unit': 'steps/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'impl': 'reference',
        'config': {'workload': 'NMPC CSTR reference tracking nx=4 nu=2 N=50 Radau-3 collocation, cold start per '
                               'instance', 'sample_instances_per_step': per_step},
        'cpu_baseline': {'value': value, 'unit': 'steps/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{per_step} instances per step x {args.steps} steps, {cores} processes; NumPy/'
                                   f'SciPy restatement of the CasADi+IPOPT path (reference not installable: no casadi)'},
        'e2e': {'value': value, 'unit': 'steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=BATCH_PER_GPU, help='instances per GPU')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-gp', action='store_true', help='skip the GP-predict side metric')
    ap.add_argument('--debug-steps', action='store_true')
    ap.add_argument('--prewarm-ms', type=float, default=600.)
    ap.add_argument('--clock-sampler', default='nvml', choices=['nvml', 'smi', 'none'])
    ap.add_argument('--cpu-sample', type=int, default=16)
    ap.add_argument('--hessian', default='exact', choices=['exact', 'gauss-newton'])
    ap.add_argument('--riccati-variant', type=int, default=None, help='override b200.riccati_variant (tuning runs)')
    ap.add_argument('--no-compaction', action='store_true', help='b200.compaction=0: every launch over the whole batch')
    ap.add_argument('--warm-start', type=int, default=1, help='1: reuse multipliers + small initial barrier parameter '
                                                              'between closed-loop steps; 0: IPOPT-default restart')
    ap.add_argument('--warm-mu', type=float, default=1e-8,
                    help='initial barrier parameter of warm-started solves (library default 1e-4; 1e-8 = one decade '
                         'above the final one, tol/10: all instances of this closed loop still converge to tol)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from hilo_mpc_b200.examples import cstr_initial_states, cstr_nmpc

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (the B200 engine has no CPU fallback)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'   # the version banner goes to stdout, which carries exactly one JSON line
        dist.init_process_group('nccl', device_id=dev)

    B = args.batch
    opts = {'b200.hessian': args.hessian, 'b200.warm_start': args.warm_start, 'b200.warm_start_mu': args.warm_mu}
    if args.riccati_variant is not None:
        opts['b200.riccati_variant'] = args.riccati_variant
    if args.no_compaction:
        opts['b200.compaction'] = 0
    nmpc = cstr_nmpc(N=N_HORIZON, solver_options=opts)
    eng = nmpc.engine
    xs = torch.tensor(nmpc._xs_full, dtype=torch.float64, device=dev)
    us = torch.tensor(nmpc._us_full, dtype=torch.float64, device=dev)
    x0 = torch.tensor(cstr_initial_states(B, seed=rank), dtype=torch.float64, device=dev)
    n_total = args.warmup + args.steps
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    noise = (torch.rand(2 * n_total + 2, B, NX, generator=gen, device=dev, dtype=torch.float64) * 2 - 1) * \
        torch.tensor(DIST_SCALE, dtype=torch.float64, device=dev)
    num = nmpc._solver_numbers()
    dt = nmpc.sampling_interval

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ device-resident arm (`value`)
    kernel_ms = np.zeros(len(KERNEL_NAMES))
    kernel_count = np.zeros(len(KERNEL_NAMES), dtype=np.int64)
    iters_hist = []

    def step_device(k, x_cur, timing):
        out = eng.solve(x_cur / xs, dt=dt, time0=0., kernel_ms=kernel_ms if timing else None,
                        kernel_count=kernel_count if timing else None, **num)
        x_next = out['x'][:, 1, :] * xs + noise[k]
        return x_next, out

    # GPU ramp-up: the timed region is only K x ~6 ms long, shorter than the time the part needs to settle after idling
    # (the same loop measured 15 % slower right after start-up than half a second later).  Solve the initial states
    # repeatedly for --prewarm-ms before the W warm-up steps; the closed loop then starts cold from x0 as usual.
    t_pre = time.perf_counter()
    while (time.perf_counter() - t_pre) * 1e3 < args.prewarm_ms:
        eng.reset_guess()
        eng.solve(x0 / xs, dt=dt, time0=0., **num)
        torch.cuda.synchronize()
    eng.reset_guess()
    x_cur = x0
    for k in range(args.warmup):
        x_cur, out = step_device(k, x_cur, False)
    barrier()
    sampler = ClockSampler(local_rank, mode=args.clock_sampler)
    if rank == 0:
        sampler.start()
    launches0 = getattr(eng, 'total_kernel_launches', 0)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    outs, launched_hist, launched_e2e, dbg = [], [], [], []
    for k in range(args.steps):
        if args.debug_steps:
            dbg.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
            dbg[-1][0].record()
        x_cur, out = step_device(args.warmup + k, x_cur, False)
        if args.debug_steps:
            dbg[-1][1].record()
        launched_hist.append(eng.last_launched)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = getattr(eng, 'total_kernel_launches', 0) - launches0
    iters_hist = [int(out['iters'].max().item())]   # last step; per-step launch counts are in launched_hist
    if args.debug_steps:
        print('value-arm step ms', [round(a.elapsed_time(b), 2) for a, b in dbg], file=sys.stderr)
    # per-kernel pass: the same closed loop continued for K more steps with a CUDA-event pair around every launch (on
    # the launching stream, inside the C-ABI call).  Kept out of the `value` region because ~330 event records and
    # elapsed-time reads per step cost about 3 ms of device idle time per step.
    x_k = x_cur
    for k in range(args.steps):
        x_k, _ = step_device(args.warmup + k, x_k, True)
    torch.cuda.synchronize()
    status_ok = float((out['status'] == 1).double().mean().item())
    iters_mean = float(out['iters'].double().mean().item())
    clocks = sampler.stop() if rank == 0 else None
    t_ms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = t_ms.item() / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # ------------------------------------------------------------------ end-to-end arm through the public API (host buffers)
    # The same closed loop as above restarted from the same initial states and disturbance sequence, so that both arms
    # see the same solves (the loop settles towards the set point and later steps need fewer iterations).
    x_host = torch.empty(B, NX, dtype=torch.float64).pin_memory()
    u_host = torch.empty(B, NU, dtype=torch.float64).pin_memory()
    x_host.copy_(x0.cpu())
    noise_h = noise[:n_total].cpu()
    eng.reset_guess()

    def step_e2e(k):
        u = nmpc.optimize(x_host.to(dev, non_blocking=True))  # H2D of this step's measured states
        u_host.copy_(u, non_blocking=False)                    # D2H of the applied inputs
        x_pred, _, _ = nmpc.return_prediction()
        x_host.copy_(x_pred[:, 1, :NX])                        # plant model: predicted next state (+ disturbance)
        x_host.add_(noise_h[k])

    for k in range(args.warmup):
        step_e2e(k)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        step_e2e(args.warmup + k)
        launched_e2e.append(eng.last_launched)
    e1.record()
    barrier()
    t_e = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * B / (t_e.item() / args.steps * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ roofline of the dominant kernel
    peak, peak_kind = _peaks()
    d = eng.dims
    sd, td = d.stage_doubles, d.elim_doubles
    # algorithmic bytes per launch: per instance
    nx, nu, nc, N = d.nx, d.nu, d.nc, d.N
    bytes_riccati = B * (N * (sd + 2 * nx + nu) + (nx * nx + nx) + 2 * nx) * 8
    bytes_eval = B * (N * ((nx + nu + nc) + sd + td) + (nx + nx * nx + nx)) * 8
    alg = {'riccati': bytes_riccati, 'eval': bytes_eval}
    per_kernel = {}
    for i, nm in enumerate(KERNEL_NAMES):
        if kernel_count[i]:
            per_kernel[nm] = {'ms_total': float(kernel_ms[i]), 'launches': int(kernel_count[i]),
                              'avg_us': float(1e3 * kernel_ms[i] / kernel_count[i])}
    dominant = max(per_kernel, key=lambda n: per_kernel[n]['ms_total']) if per_kernel else 'riccati'
    roof_name = dominant if dominant in alg else 'riccati'
    avg_s = per_kernel[roof_name]['avg_us'] * 1e-6 if roof_name in per_kernel else float('nan')
    achieved = alg[roof_name] / avg_s / 1e9
    roofline = {'kernel': roof_name, 'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': None, 'peak_kind': peak_kind,
                'timing': 'CUDA event pair around every launch on the launching stream; pass of K closed-loop steps '
                          'right after the value region (events inside it would add ~3 ms/step of idle time)',
                'algorithmic_bytes_per_launch': alg[roof_name]}
    ric = per_kernel.get('riccati')
    if ric:
        roofline['riccati_sweep'] = {'achieved': bytes_riccati / (ric['avg_us'] * 1e-6) / 1e9,
                                     'frac': bytes_riccati / (ric['avg_us'] * 1e-6) / 1e9 / peak,
                                     'avg_us': ric['avg_us']}

    # measured DRAM traffic per launch (ncu --set full, profiles/ncu_traffic.json; refreshed with every kernel change)
    traffic_path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if os.path.exists(traffic_path):
        with open(traffic_path) as f:
            tr = json.load(f)
        if roof_name in tr:
            roofline['traffic'] = tr[roof_name]['dram_bytes_per_launch']
            roofline['traffic_source'] = tr[roof_name].get('source')
        if 'riccati' in tr and 'riccati_sweep' in roofline:
            roofline['riccati_sweep']['traffic'] = tr['riccati']['dram_bytes_per_launch']

    extras = {}
    if not args.no_gp:
        extras['gp_predict'] = gp_predict_line(dev)

    cpu_baseline = None
    if not args.no_cpu_baseline:
        v, el = cpu_port_sample(args.cpu_sample, 1)
        cpu_baseline = {'value': v, 'unit': 'steps/s', 'cores': 1, 'kind': 'port',
                        'sample': f'{args.cpu_sample} cold-start instances of the same workload, {el:.1f} s, single '
                                  f'process; NumPy/SciPy restatement of the CasADi+IPOPT path'}

    work_mb = eng._t['work'].numel() * 8 / 1e6
    line = {
        'metric': 'batched NMPC steps/sec (CSTR, N=50)', 'value': value, 'unit': 'steps/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'NMPC CSTR reference tracking nx=4 nu=2 N=50 Radau-3 collocation, batch={B}/GPU, '
                               f'warm-started closed loop with seeded disturbances, tol=1e-8',
                   'batch_per_gpu': B, 'horizon': N_HORIZON, 'parallelism': f'batch-shard x{world}',
                   'l2': f'working set {work_mb:.0f} MB per step > 126 MB L2 (no explicit flush)',
                   'hessian': args.hessian, 'warm_start': bool(args.warm_start), 'warm_start_mu': args.warm_mu,
                   'ipm_iterations_mean': iters_mean, 'ipm_iterations_max_last_step': iters_hist[0],
                   'ipm_iterations_launched_per_step': launched_hist, 'ipm_iterations_launched_per_step_e2e': launched_e2e,
                   'status_optimal_frac': status_ok},
        'e2e': {'value': e2e_value, 'unit': 'steps/s', 'h2d_bytes_per_step': B * NX * 8,
                'd2h_bytes_per_step': B * NU * 8 + B * (N_HORIZON + 1) * 0 + B * NX * 8},
        'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roofline, 'kernels': per_kernel,
        'cpu_baseline': cpu_baseline, 'extras': extras,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
