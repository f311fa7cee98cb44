#!/usr/bin/env python
"""Benchmarks of the BASELINE configurations (BASELINE.json ``configs``), one JSON line per run.

  python bench.py [--config 2|3|4|5] --gpus N --steps K --warmup W          (default: config 2, the headline metric)
  python bench.py --impl reference [--config ...]                            (CPU arm, rank 0 only)

  2  NMPC CSTR reference tracking, nx=4 nu=2 N=50 Radau-3, 4096 instances per GPU          -> NMPC steps/s
  3  the same controller with the side-reaction rate learned by a GP (512 training points,
     SE kernel) inside the node kernels, 1024 instances per GPU                             -> NMPC steps/s (+ GP us/pt)
  4  MHE (N=20, 4 states + 1 parameter) + UKF on the CSTR, 4096 filters per GPU             -> estimates/s
  5  path-following NMPC of a mobile robot with MLP dynamics (1 x 32), N=30, 4096 per GPU   -> NMPC steps/s

Contract: under torchrun every rank runs its own shard of the batch (weak scaling, no data-path collective - the
instances are independent), the timed region is bracketed by barrier + synchronize and reduced with MAX over ranks, rank
0 prints ONE JSON line.  One "step" = one warm-started solve of every instance of the batch to IPOPT tolerance (1e-8)
in a closed loop: the next measured state is the predicted next state plus a seeded disturbance (configs 2, 3, 5) / the
next sample of a seeded synthetic measurement stream (config 4).

``value``: inputs resident in HBM.  ``e2e``: the same loop through the public API (``NMPC.optimize`` /
``MHE.add_measurements + estimate`` / ``UKF.estimate``) with HOST buffers: pinned host -> device copy of every step's
inputs and device -> host read of its results inside the timed region.

``--impl reference``: the real reference (CasADi + IPOPT) cannot run here (CasADi is not installable);
the arm times the oracle - the NumPy/SciPy restatement of the reference's flat NLP and of IPOPT's algorithm - on the
host cores, on a bounded sample of the SAME closed loop (same initial states, disturbances, primal warm start as
mpc.py:725-726, IPOPT-default restart of the barrier parameter every step, which is what the reference does).
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

KERNEL_NAMES = "init,eval,riccati,expand,prep,trial,decide,apply,commit,count,step,ricfix".split(',')
TIMED_ITERS = 8
CSTR_DIST = np.array([.02, .02, .5, .5])    # closed-loop disturbance added to the predicted next state (configs 2, 3)
ROBOT_DIST = np.array([.05, .05, .02, .01])  # config 5


def _peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """Samples SM clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).  NVML is queried
    in-process from a background thread (the ctypes solve call releases the GIL); an `nvidia-smi -lms` child process is
    the fallback (it cost ~3 % of the step time on the shared boxes, hence the in-process default)."""

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


# =====================================================================================================================
# problem definitions shared by both arms (everything that makes the workload: sizes, seeds, disturbances)
# =====================================================================================================================
WORKLOADS = {
    2: dict(metric='batched NMPC steps/sec (CSTR, N=50)', unit='steps/s', batch=4096,
            workload='NMPC CSTR reference tracking nx=4 nu=2 N=50 Radau-3 collocation, closed loop from '
                     'x0 = x_ss + diag(.2,.2,5,5) U(-1,1) (rng seed = rank) with seeded disturbances, primal warm '
                     'start, tol=1e-8'),
    3: dict(metric='batched NMPC steps/sec (CSTR + GP side reaction, N=50)', unit='steps/s', batch=1024,
            workload='GP-output-feedback NMPC CSTR: side-reaction rate = posterior mean of a GP (512 training points, SE '
                     'ARD kernel, 2 features) evaluated inside the node kernels, nx=4 nu=2 N=50 Radau-3, closed loop '
                     'with seeded disturbances, primal warm start, tol=1e-8'),
    4: dict(metric='batched MHE + UKF estimates/sec (CSTR, N=20)', unit='estimates/s', batch=4096,
            workload='MHE (horizon 20, Radau-3, 4 states + 1 estimated parameter, state noise, arrival cost) + UKF on the '
                     'CSTR, seeded synthetic measurement stream (cstr_mhe_data, seed 2 + rank), one estimate = '
                     'add_measurements + MHE.estimate + UKF.estimate, tol=1e-8, warm start shifted with the window'),
    5: dict(metric='batched NMPC steps/sec (path following, MLP dynamics, N=30)', unit='steps/s', batch=4096,
            workload='path-following NMPC of a mobile robot, right-hand side = 1x32 tanh MLP, path variable, nx=5 nu=3 '
                     'N=30 dt=1 Radau-3, closed loop from x0 = [15,30,0,0] + diag(2,2,.5,.2) U(-1,1) (seed 3 + rank) '
                     'with seeded disturbances, primal warm start, tol=1e-8, stage-Hessian convexification'),
}


def make_nmpc(config, solver_options=None, build=True):
    """The controller of configs 2, 3, 5 (config 3 needs a CUDA device: the GP is fitted by hb_gp_fit)."""
    from hilo_mpc_b200.examples import cstr_gp_nmpc, cstr_nmpc, cstr_side_reaction_gp, robot_nn_nmpc
    if config == 2:
        nmpc = cstr_nmpc(N=50, setup=False)
    elif config == 3:
        gp = cstr_side_reaction_gp(n_train=512)
        nmpc = cstr_gp_nmpc(gp, N=50, setup=False)
    elif config == 5:
        nmpc = robot_nn_nmpc(N=30, hidden=32, setup=False)
        solver_options = dict({'b200.regularization': 'mirror', 'b200.regularization_mu': 1e-4, 'ipopt.max_iter': 1500},
                              **(solver_options or {}))
    else:
        raise ValueError(config)
    if solver_options:
        nmpc.set_solver_opts(solver_options)
    if not build:
        nmpc._build_module = False
    nmpc.setup()
    return nmpc


def initial_states(config, B, rank):
    from hilo_mpc_b200.examples import cstr_initial_states, robot_initial_states
    if config in (2, 3):
        return cstr_initial_states(B, seed=rank)
    return robot_initial_states(B, seed=3 + rank)


def disturbance(config, n_steps, B, rank):
    """Seeded disturbance sequence [n_steps, B, nx] added to the predicted next state (NumPy: both arms use it)."""
    rng = np.random.default_rng(1234 + rank)
    scale = CSTR_DIST if config in (2, 3) else ROBOT_DIST
    return (rng.random((n_steps, B, 4)) * 2. - 1.) * scale


# =====================================================================================================================
# CPU arm: the oracle on a bounded sample of the same closed loop
# =====================================================================================================================
_CPU_CACHE = {}


def _cpu_nlp(config):
    if config not in _CPU_CACHE:
        from oracle.nmpc import FlatNlp
        if config == 4:
            from hilo_mpc_b200.examples import cstr_mhe
            mhe = cstr_mhe(N=20, setup=False)
            mhe._build_module = False
            mhe.setup()
            nlp = FlatNlp(mhe.spec, exact_hessian=True)
            _CPU_CACHE[config] = (nlp, mhe)
        else:
            nmpc = make_nmpc(config, build=False)
            nlp = FlatNlp(nmpc.spec, exact_hessian=True)
            _CPU_CACHE[config] = (nlp, nmpc)
    return _CPU_CACHE[config]


def _cpu_closed_loop(arg):
    # one BLAS/OpenMP thread per worker: the arm reports processes as `cores`, and 16 workers x 16 library threads on 16
    # cores would only fight each other
    from threadpoolctl import threadpool_limits
    with threadpool_limits(limits=1):
        return _cpu_closed_loop_body(arg)


def _cpu_closed_loop_body(arg):
    """One worker: for each instance of its chunk an UNTIMED first solve (cold start: the GPU arm keeps it out of its
    timed region as well), then `steps` timed closed-loop solves with the primal warm start of the reference
    (mpc.py:725-726).  Returns (timed solves, converged, seconds in the timed solves)."""
    config, idx, steps, rank = arg
    nlp, ctl = _cpu_nlp(config)
    done = ok = 0
    secs = 0.
    if config == 4:
        from hilo_mpc_b200.examples import cstr_mhe_data
        N = 20
        u, y, _, _ = cstr_mhe_data(int(max(idx)) + 1, N + steps + 1, seed=2 + rank)
        xs = np.concatenate([ctl._x_scaling, ctl._p_scaling])
        for b in idx:
            ap = np.concatenate([ctl.quad_arrival_cost.x_guess, ctl.quad_arrival_cost.p_guess])
            v0 = None
            for k in range(steps + 1):
                sp = np.concatenate([u[k:k + N, b], y[k:k + N, b]], axis=1)
                t0 = time.perf_counter()
                res, xo, _ = nlp.solve(None, v0=v0, dt=ctl.sampling_interval, sps=sp, ap=ap)
                el = time.perf_counter() - t0
                v0 = res.x
                ap = np.concatenate([xo[:4, 2] * xs[:4], xo[4:, 0] * xs[4:]])   # smoothing update (mhe.py:254-256)
                if k > 0:
                    done += 1
                    ok += int(res.status in (1, 2))
                    secs += el
        return done, ok, secs
    x0 = initial_states(config, int(max(idx)) + 1, rank)
    noise = disturbance(config, steps + 1, int(max(idx)) + 1, rank)
    xs = np.asarray(ctl._xs_full)
    nx0 = 4
    for b in idx:
        x = x0[b].copy()
        theta = 0.
        v0 = None
        for k in range(steps + 1):
            xa = x if len(xs) == nx0 else np.concatenate([x, [theta]])
            t0 = time.perf_counter()
            res, xo, _ = nlp.solve(xa / xs, v0=v0, dt=ctl.sampling_interval)
            el = time.perf_counter() - t0
            v0 = res.x                                     # primal warm start only (mpc.py:725-726)
            x = xo[:nx0, 1] * xs[:nx0] + noise[k, b]
            if len(xs) > nx0:
                theta = xo[nx0, 1] * xs[nx0]
            if k > 0:
                done += 1
                ok += int(res.status in (1, 2))
                secs += el
    return done, ok, secs


CPU_SAMPLE_DEADLINE_S = {2: 150., 3: 150., 4: 150., 5: 300.}   # wall-clock bound of one CPU sample (cold solves included)


def cpu_sample(config, n_instances, steps, n_procs, rank=0):
    """Times the oracle on the first ``n_instances`` instances of the workload for ``steps`` warm closed-loop steps each
    (after an untimed cold first solve) with ``n_procs`` worker processes, one library thread each.  Workers that have
    not finished at the sample's deadline are dropped (the arm must end within minutes whatever the host): the rate is
    (mean rate of the finished workers) x ``n_procs``.  Returns (units_per_s, seconds of the timed solves on the slowest
    finished worker, converged fraction, wall seconds incl. the cold solves)."""
    from threadpoolctl import threadpool_limits
    idx = np.arange(n_instances)
    with threadpool_limits(limits=1):   # also in the parent: no library thread pool may exist when the workers are forked
        _cpu_nlp(config)   # in the parent: forked workers inherit the problem (config 3 fits its GP on the device here)
        t_wall = time.perf_counter()
        if n_procs <= 1:
            done, ok, el = _cpu_closed_loop((config, idx, steps, rank))
            return done / el, el, ok / max(done, 1), time.perf_counter() - t_wall
        import multiprocessing as mp
        chunks = [c for c in np.array_split(idx, n_procs) if len(c)]
        res = []
        pool = mp.get_context('fork').Pool(len(chunks))
        try:
            it = pool.imap_unordered(_cpu_closed_loop, [(config, c, steps, rank) for c in chunks])
            deadline = t_wall + CPU_SAMPLE_DEADLINE_S[config]
            for _ in chunks:
                try:
                    res.append(it.next(timeout=max(1., deadline - time.perf_counter())))
                except mp.TimeoutError:
                    break
        finally:
            pool.terminate()
            pool.join()
    if not res:
        raise RuntimeError(f'no CPU worker finished within {CPU_SAMPLE_DEADLINE_S[config]:.0f} s')
    done, ok = sum(r[0] for r in res), sum(r[1] for r in res)
    rate = len(chunks) * float(np.mean([r[0] / r[2] for r in res]))
    cpu_sample.last_finished = (len(res), len(chunks))
    return rate, max(r[2] for r in res), ok / max(done, 1), time.perf_counter() - t_wall


def run_reference(args):
    """``--impl reference``: CPU arm, rank 0 only."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    w = WORKLOADS[args.config]
    cores = os.cpu_count() or 1
    # config 3 needs the GP fitted on a device (hb_gp_fit); the CPU arm then differs from config 2 only by the GP terms
    per_worker_steps = {2: 4, 3: 3, 4: 4, 5: 2}[args.config]
    n_inst = cores
    times, rates = [], []
    t_budget = time.perf_counter()
    for k in range(max(1, args.steps // 10)):
        v, el, okf, wall = cpu_sample(args.config, n_inst, per_worker_steps, cores)
        times.append(n_inst * per_worker_steps / v)   # seconds one sample takes at the measured rate
        rates.append(v)
        if time.perf_counter() - t_budget > 30.:   # bounded sample: the whole arm stays within a few minutes
            break
    total = float(np.sum(times))
    value = float(np.mean(rates))
    fin = getattr(cpu_sample, 'last_finished', (cores, cores))
    line = {
        'metric': w['metric'], 'value': value, 'unit': w['unit'], 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * total / max(1, len(times)), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'impl': 'reference',
        'config': {'workload': w['workload'], 'config': args.config, 'sample_instances': n_inst,
                   'sample_closed_loop_steps': per_worker_steps,
                   'note': 'one timed sample = the first `sample_instances` instances of the workload for '
                           '`sample_closed_loop_steps` warm closed-loop steps each (after an untimed cold first solve, '
                           'which the GPU arm also keeps out of its timed region), one instance per worker process'},
        'cpu_baseline': {'value': value, 'unit': w['unit'], 'cores': cores, 'kind': 'port',
                         'sample': f'{n_inst} instances x {per_worker_steps} closed-loop steps x {len(times)} repeats, '
                                   f'{cores} processes ({fin[0]} of {fin[1]} finished within the sample deadline); NumPy/SciPy restatement of the reference\'s flat NLP + IPOPT '
                                   f'algorithm (the reference itself is not installable here: no CasADi)'},
        'e2e': {'value': value, 'unit': w['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# =====================================================================================================================
# GPU arm
# =====================================================================================================================
def gp_predict_line(dev, n: int = 512, d: int = 6, m: int = 65536, reps: int = 5):
    """Second BASELINE metric: GP predict microseconds per query point (SE kernel, n=512 training points, d=6 features),
    posterior mean + variance and mean only, through ``GaussianProcess.predict`` with device tensors.  Flops counted
    as executed: Gram n (3d + 4), mean 2n, variance = the triangular half of L^-1 k*: n^2 (the kernel skips the zero half)."""
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
        flops_pt = n * (3 * d + 4) + 2 * n + (0 if mean_only else n * n)
        out[key] = {'us_per_pt': us_pt, 'gflops': flops_pt / us_pt * 1e-3}
    return out


def measured_peaks_in_run(dev):
    """Stream-copy bandwidth and fp64 GEMM throughput of THIS box: torch device copy of 2 GiB
    (read + write bytes) and cuBLAS DGEMM 8192^3, best of 5, CUDA events."""
    import torch
    out = {}
    a = torch.empty(1 << 28, dtype=torch.float64, device=dev)
    b = torch.empty_like(a)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        b.copy_(a)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out['copy_gbs'] = 2 * a.numel() * 8 / (best * 1e-3) / 1e9
    del a, b
    n = 8192
    x = torch.randn(n, n, dtype=torch.float64, device=dev)
    y = torch.randn(n, n, dtype=torch.float64, device=dev)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(x, y)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out['dgemm_tflops'] = 2 * n ** 3 / (best * 1e-3) / 1e12
    out['how'] = 'torch b.copy_(a) over 2 GiB fp64 (read+write bytes, best of 5); torch.matmul fp64 8192^3 (best of 3)'
    return out


def _extra_opts(args):
    out = {}
    for kv in args.opt:
        k, v = kv.split('=', 1)
        try:
            out[k] = int(v)
        except ValueError:
            try:
                out[k] = float(v)
            except ValueError:
                out[k] = v
    return out


class NmpcLoop:
    """Closed loop of configs 2, 3, 5."""

    def __init__(self, config, args, dev, rank):
        import torch
        self.torch, self.config, self.dev, self.rank = torch, config, dev, rank
        self.B = args.batch or WORKLOADS[config]['batch']
        opts = {'b200.hessian': args.hessian, 'b200.warm_start': args.warm_start, 'b200.warm_start_mu': args.warm_mu}
        if config == 5:
            # path following: the previous solution moved one stage towards the present is the better guess (the path
            # parameter advances every step); measured 17.7 instead of 20.8 iterations per step
            opts['b200.shift_warm_start'] = 1
        if args.riccati_variant is not None:
            opts['b200.riccati_variant'] = args.riccati_variant
        if args.no_compaction:
            opts['b200.compaction'] = 0
        opts.update(_extra_opts(args))
        self.nmpc = make_nmpc(config, opts)
        self.eng = self.nmpc.engine
        self.nx0 = 4
        self.x0 = torch.tensor(initial_states(config, self.B, rank), dtype=torch.float64, device=dev)
        self.n_total = 2 * (args.warmup + args.steps) + 4
        self.noise_h = torch.tensor(disturbance(config, self.n_total, self.B, rank)).pin_memory()
        self.noise = self.noise_h.to(dev)
        self.h2d = self.B * self.nx0 * 8
        self.d2h = self.B * self.nmpc._nu0 * 8 + self.B * self.nx0 * 8
        self.kernel_ms = np.zeros(len(KERNEL_NAMES))
        self.kernel_count = np.zeros(len(KERNEL_NAMES), dtype=np.int64)
        self.kernel_ms_it = np.zeros((TIMED_ITERS, len(KERNEL_NAMES)))
        self.kernel_count_it = np.zeros((TIMED_ITERS, len(KERNEL_NAMES)), dtype=np.int64)
        self._timing = False
        self._orig_solve = self.eng.solve
        eng = self.eng

        def solve(*a, **kw):   # the per-kernel pass passes the timing arrays through the public call
            if self._timing:
                kw.update(kernel_ms=self.kernel_ms, kernel_count=self.kernel_count, kernel_ms_by_iter=self.kernel_ms_it,
                          kernel_count_by_iter=self.kernel_count_it)
            return self._orig_solve(*a, **kw)
        eng.solve = solve

    def prewarm(self, ms):
        """GPU ramp-up before the warm-up steps (the part runs ~5 % slower in the first half second after idling) and
        the cold solve of the initial states."""
        torch = self.torch
        t0 = time.perf_counter()
        while True:
            self.restart()
            self.nmpc.optimize(self.x0)
            torch.cuda.synchronize()
            if (time.perf_counter() - t0) * 1e3 >= ms:
                break
        self.restart()

    def restart(self):
        self.eng.reset_guess()
        self.nmpc._batch = None
        self.nmpc._last = None
        self.nmpc._time = 0.
        self.x_cur = self.x0
        self.num = self.nmpc._solver_numbers()
        nth = self.nmpc.spec.nx - self.nx0
        self.theta = self.torch.zeros(self.B, max(nth, 0), dtype=self.torch.float64, device=self.dev)

    def step_device(self, k, timing=False):
        """Device-resident step: the engine call of NMPC.optimize (hb_ocp_solve through OcpEngine.solve) on tensors that
        never leave the device; the path variable of config 5 is carried like NMPC.optimize does (mpc.py:1173-1204)."""
        torch, nmpc = self.torch, self.nmpc
        xs, _ = nmpc._scaling_tensors(torch, self.dev)
        x = self.x_cur
        if nmpc.spec.nx > self.nx0:
            x = torch.cat([x, self.theta], dim=1)
        self._timing = timing
        out = self.eng.solve(x / xs, dt=nmpc.sampling_interval, time0=0., **self.num)
        self._timing = False
        nmpc._last = out
        self.x_cur = out['x'][:, 1, :self.nx0] * xs[:self.nx0] + self.noise[k]
        if nmpc.spec.nx > self.nx0:
            self.theta = out['x'][:, 1, self.nx0:] * xs[self.nx0:]
        return out

    def start_e2e(self):
        torch = self.torch
        self.x_host = torch.empty(self.B, self.nx0, dtype=torch.float64).pin_memory()
        self.u_host = torch.empty(self.B, self.nmpc._nu0, dtype=torch.float64).pin_memory()
        self.x_host.copy_(self.x0.cpu())
        self.restart()

    def step_e2e(self, k):
        u = self.nmpc.optimize(self.x_host.to(self.dev, non_blocking=True))   # H2D of this step's measured states
        self.u_host.copy_(u)                                                   # D2H of the applied inputs
        x_pred, _, _ = self.nmpc.return_prediction()
        self.x_host.copy_(x_pred[:, 1, :self.nx0])                             # plant: predicted next state (+ disturbance)
        self.x_host.add_(self.noise_h[k])

    def stats(self):
        out = self.nmpc._last
        st = out['status']
        return {'status_optimal_frac': float((st == 1).double().mean().item()),
                'status_acceptable_frac': float((st == 2).double().mean().item()),
                'ipm_iterations_mean': float(out['iters'].double().mean().item()),
                'ipm_iterations_max': int(out['iters'].max().item())}

    def algorithmic_bytes(self):
        d = self.eng.dims
        B, N, nx, nu, nc, sd, td = self.B, d.N, d.nx, d.nu, d.nc, d.stage_doubles, d.elim_doubles
        return {'riccati': B * (N * (sd + 2 * nx + nu) + (nx * nx + nx) + 2 * nx) * 8,
                'eval': B * (N * ((nx + nu + nc) + sd + td) + (nx + nx * nx + nx)) * 8}


class MheUkfLoop:
    """Config 4: MHE + UKF on a synthetic measurement stream."""

    def __init__(self, config, args, dev, rank):
        import torch
        from hilo_mpc_b200 import UKF
        from hilo_mpc_b200.examples import cstr_mhe, cstr_mhe_data
        from hilo_mpc_b200.library import cstr_model, cstr_steady_state
        self.torch, self.dev, self.rank = torch, dev, rank
        self.B = args.batch or WORKLOADS[4]['batch']
        self.N = 20
        # the horizon window moves by one sample per estimate: the kept solution is shifted with it
        opts = {'b200.warm_start': args.warm_start, 'b200.warm_start_mu': args.warm_mu, 'b200.shift_warm_start': 1}
        opts.update(_extra_opts(args))
        self.opts = opts
        self.make_mhe = lambda: cstr_mhe(N=self.N, solver_options=opts)
        self.mhe = self.make_mhe()
        self.eng = self.mhe.engine
        self.n_total = 2 * (args.warmup + args.steps) + 4
        u, y, _, _ = cstr_mhe_data(self.B, self.N + self.n_total, seed=2 + rank)
        self.u_h = torch.tensor(u).pin_memory()
        self.y_h = torch.tensor(y).pin_memory()
        self.u_d, self.y_d = self.u_h.to(dev), self.y_h.to(dev)
        x_ss, _ = cstr_steady_state()
        self.x_ss = x_ss

        def make_ukf():
            ukf = UKF(cstr_model(dt=.005, name='cstr_k1', with_parameter=True), alpha=.5)
            ukf.setup()
            ukf.Q = np.diag((1e-2 * x_ss) ** 2 * .005)
            ukf.R = np.diag([1e-4, 1e-2])
            ukf.set_initial_guess(np.tile(x_ss, (self.B, 1)), P0=np.diag([1e-2, 1e-2, 1., 1.]))
            return ukf
        self.make_ukf = make_ukf
        self.ukf = make_ukf()
        self.h2d = self.B * 4 * 8
        self.d2h = self.B * (4 + 1 + 4) * 8
        self.kernel_ms = np.zeros(len(KERNEL_NAMES))
        self.kernel_count = np.zeros(len(KERNEL_NAMES), dtype=np.int64)
        self.kernel_ms_it = np.zeros((TIMED_ITERS, len(KERNEL_NAMES)))
        self.kernel_count_it = np.zeros((TIMED_ITERS, len(KERNEL_NAMES)), dtype=np.int64)
        self._timing = False
        self._wrap()
        self.ukf_ms = 0.

    def _wrap(self):
        orig = self.mhe.engine.solve

        def solve(*a, **kw):
            if self._timing:
                kw.update(kernel_ms=self.kernel_ms, kernel_count=self.kernel_count, kernel_ms_by_iter=self.kernel_ms_it,
                          kernel_count_by_iter=self.kernel_count_it)
            return orig(*a, **kw)
        self.mhe.engine.solve = solve
        self.eng = self.mhe.engine

    def _fill(self, u, y):
        """The first N samples fill the horizon window (no estimate before, mhe.py:300-302)."""
        for k in range(self.N):
            self.mhe.add_measurements(y[k], u_meas=u[k])
            self.mhe.estimate() if k < self.N - 1 else None

    def prewarm(self, ms):
        self.restart()

    def restart(self):
        torch = self.torch
        self.mhe = self.make_mhe()
        self._wrap()
        self.ukf = self.make_ukf()
        self._fill(self.u_d, self.y_d)
        self.p_est = torch.ones(self.B, 1, dtype=torch.float64, device=self.dev)
        torch.cuda.synchronize()

    def step_device(self, k, timing=False):
        kk = self.N - 1 + k    # sample N-1 completes the first window: step 0 estimates on samples 0..N-1
        if k > 0:
            self.mhe.add_measurements(self.y_d[kk], u_meas=self.u_d[kk])
        self._timing = timing
        x_est, p_est = self.mhe.estimate()
        self._timing = False
        self.ukf.estimate(y=self.y_d[kk], u=self.u_d[kk], p=p_est)
        return self.mhe._last

    def start_e2e(self):
        torch = self.torch
        self.restart_host()
        self.out_h = torch.empty(self.B, 9, dtype=torch.float64).pin_memory()

    def restart_host(self):
        self.mhe = self.make_mhe()
        self._wrap()
        self.ukf = self.make_ukf()
        self._fill(self.u_h, self.y_h)
        self.torch.cuda.synchronize()

    def step_e2e(self, k):
        kk = self.N - 1 + k
        dev = self.dev
        y = self.y_h[kk].to(dev, non_blocking=True)          # H2D of this step's measurements
        u = self.u_h[kk].to(dev, non_blocking=True)
        if k > 0:
            self.mhe.add_measurements(y, u_meas=u)
        x_est, p_est = self.mhe.estimate()
        x_ukf, _ = self.ukf.estimate(y=y, u=u, p=p_est)
        self.out_h[:, :4].copy_(x_est)                       # D2H of the estimates
        self.out_h[:, 4:5].copy_(p_est)
        self.out_h[:, 5:].copy_(x_ukf)

    def stats(self):
        out = self.mhe._last
        st = out['status']
        return {'status_optimal_frac': float((st == 1).double().mean().item()),
                'status_acceptable_frac': float((st == 2).double().mean().item()),
                'ipm_iterations_mean': float(out['iters'].double().mean().item()),
                'ipm_iterations_max': int(out['iters'].max().item())}

    def algorithmic_bytes(self):
        d = self.eng.dims
        B, N, nx, nu, nc, sd, td = self.B, d.N, d.nx, d.nu, d.nc, d.stage_doubles, d.elim_doubles
        return {'riccati': B * (N * (sd + 2 * nx + nu) + (nx * nx + nx) + 2 * nx) * 8,
                'eval': B * (N * ((nx + nu + nc) + sd + td) + (nx + nx * nx + nx)) * 8}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', type=int, default=2, choices=[2, 3, 4, 5])
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=None,
                    help='timed closed-loop steps (default 200; config 5: 150, see WORKLOADS[5])')
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=None, help='instances per GPU (default: the configuration\'s)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-gp', action='store_true', help='skip the GP-predict side metric')
    ap.add_argument('--no-default-options', action='store_true', help='skip the library-default-options value')
    ap.add_argument('--prewarm-ms', type=float, default=600.)
    ap.add_argument('--clock-sampler', default='nvml', choices=['nvml', 'smi', 'none'])
    ap.add_argument('--cpu-sample', type=int, default=4, help='instances of the single-core CPU baseline sample')
    ap.add_argument('--hessian', default='exact', choices=['exact', 'gauss-newton'])
    ap.add_argument('--riccati-variant', type=int, default=None, help='override b200.riccati_variant (tuning runs)')
    ap.add_argument('--no-compaction', action='store_true', help='b200.compaction=0: every launch over the whole batch')
    ap.add_argument('--opt', action='append', default=[], metavar='KEY=VALUE',
                    help='extra solver option (tuning runs), e.g. --opt b200.eval_variant=0')
    ap.add_argument('--warm-start', type=int, default=1, help='1: reuse multipliers + small initial barrier parameter '
                                                              'between closed-loop steps; 0: IPOPT-default restart')
    ap.add_argument('--warm-mu', type=float, default=None,
                    help='initial barrier parameter of warm-started solves (library default 1e-4; the bench uses 1e-8 = '
                         'one decade above the final one for configs 2-4 and 1e-6 for config 5)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.steps is None:
        # config 5: the heading angle of the reference example is boxed to [-100, 100] rad and the track is lapped every
        # 12.6 steps, so the horizon's tail reaches the box at step ~180 of the closed loop and a few instances need
        # hundreds of iterations for the steps in which the car has to brake; the default stays on the stationary part
        args.steps = 150 if args.config == 5 else 200
    if args.warm_mu is None:
        args.warm_mu = 1e-9 if args.config == 5 else 1e-8   # config 5: the value the previous solve ended on
    if args.impl == 'reference':
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

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

    w = WORKLOADS[args.config]
    loop = (MheUkfLoop if args.config == 4 else NmpcLoop)(args.config, args, dev, rank)
    B = loop.B

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ device-resident arm (`value`)
    loop.prewarm(args.prewarm_ms)
    for k in range(args.warmup):
        loop.step_device(k)
    barrier()
    sampler = ClockSampler(local_rank, mode=args.clock_sampler)
    if rank == 0:
        sampler.start()
    launches0 = getattr(loop.eng, 'total_kernel_launches', 0)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    launched_hist = []
    for k in range(args.steps):
        loop.step_device(args.warmup + k)
        launched_hist.append(loop.eng.last_launched)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = getattr(loop.eng, 'total_kernel_launches', 0) - launches0
    stats = loop.stats()
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel pass: the same closed loop continued with a CUDA-event pair around every launch (on the launching
    # stream, inside the C-ABI call).  Kept out of the `value` region: ~330 event records and reads per step cost ~3 ms.
    n_kpass = min(args.steps, 40)
    for k in range(n_kpass):
        loop.step_device(args.warmup + args.steps + k, timing=True)
    torch.cuda.synchronize()
    t_ms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = t_ms.item() / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # ------------------------------------------------------------------ end-to-end arm through the public API (host buffers)
    # The same closed loop restarted from the same initial states and disturbance sequence.
    loop.start_e2e()
    for k in range(args.warmup):
        loop.step_e2e(k)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launched_e2e = []
    for k in range(args.steps):
        loop.step_e2e(args.warmup + k)
        launched_e2e.append(loop.eng.last_launched)
    e1.record()
    barrier()
    t_e = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * B / (t_e.item() / args.steps * 1e-3)

    # ------------------------------------------------------------------ library-default options (IPOPT-style restart each step)
    default_value = None
    if not args.no_default_options and args.config != 4:
        saved = dict(loop.nmpc._nlp_opts)
        loop.nmpc.set_solver_opts(dict(saved, **{'b200.warm_start': 0}))
        loop.restart()
        nd = min(args.steps, 20)
        for k in range(3):
            loop.step_device(k)
        barrier()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for k in range(nd):
            loop.step_device(3 + k)
        d1.record()
        barrier()
        t_d = torch.tensor([d0.elapsed_time(d1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_d, op=dist.ReduceOp.MAX)
        dstats = loop.stats()
        default_value = {'value': world * B / (t_d.item() / nd * 1e-3), 'unit': w['unit'], 'steps': nd,
                         'options': 'library defaults: b200.warm_start=0 (primal warm start only, mu restarts at 0.1 like '
                                    'IPOPT in the reference)', 'ipm_iterations_mean': dstats['ipm_iterations_mean']}
        loop.nmpc.set_solver_opts(saved)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ roofline of the dominant kernel
    # Quoted on iteration 0 of the solves of the per-kernel pass: the one launch of a solve in which EVERY instance of
    # the batch is active, so the algorithmic bytes of a full batch are the bytes it moves (later launches of a solve
    # only touch the instances that have not converged yet).
    peak, peak_kind = _peaks()
    alg = loop.algorithmic_bytes()
    per_kernel, full_launch = {}, {}
    for i, nm in enumerate(KERNEL_NAMES):
        if loop.kernel_count[i]:
            per_kernel[nm] = {'ms_total': float(loop.kernel_ms[i]), 'launches': int(loop.kernel_count[i]),
                              'avg_us': float(1e3 * loop.kernel_ms[i] / loop.kernel_count[i])}
        if loop.kernel_count_it[0, i]:
            full_launch[nm] = float(1e3 * loop.kernel_ms_it[0, i] / loop.kernel_count_it[0, i])
    # average launch duration per interior-point iteration of a solve (row i = i-th iteration; later rows only touch the
    # instances that are still active, through the compacted list)
    by_iteration = []
    for it in range(loop.kernel_count_it.shape[0]):
        row = {nm: round(float(1e3 * loop.kernel_ms_it[it, i] / loop.kernel_count_it[it, i]), 1)
               for i, nm in enumerate(KERNEL_NAMES) if loop.kernel_count_it[it, i]}
        if row:
            row['solves'] = int(max(loop.kernel_count_it[it]))
            by_iteration.append(row)
    dominant = max((n for n in per_kernel if n in alg), key=lambda n: per_kernel[n]['ms_total'], default='riccati')
    us_full = full_launch.get(dominant, float('nan'))
    achieved = alg[dominant] / (us_full * 1e-6) / 1e9
    roofline = {'kernel': dominant, 'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': None, 'peak_kind': peak_kind,
                'launch_us': us_full, 'algorithmic_bytes_per_launch': alg[dominant],
                'timing': 'CUDA event pair around every launch on the launching stream; average over the iteration-0 '
                          'launches (all instances active) of the solves of a pass of closed-loop steps right after the '
                          'value region', 'full_launch_us': full_launch}
    if 'riccati' in full_launch:
        a_r = alg['riccati'] / (full_launch['riccati'] * 1e-6) / 1e9
        roofline['riccati_sweep'] = {'achieved': a_r, 'frac': a_r / peak, 'launch_us': full_launch['riccati'],
                                     'algorithmic_bytes_per_launch': alg['riccati'],
                                     'note': 'the sweep kernel alone; the (normally empty) fix-up launch that follows it is '
                                             'timed as kernel "ricfix"'}
    traffic_path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if os.path.exists(traffic_path):
        with open(traffic_path) as f:
            tr = json.load(f).get(f'config{args.config}', {})
        if dominant in tr:
            roofline['traffic'] = tr[dominant]['dram_bytes_per_launch']
            roofline['traffic_source'] = tr[dominant].get('source')
        if 'riccati' in tr and 'riccati_sweep' in roofline:
            roofline['riccati_sweep']['traffic'] = tr['riccati']['dram_bytes_per_launch']

    extras = {'peaks_measured_in_run': measured_peaks_in_run(dev), 'kernel_us_by_iteration': by_iteration}
    if not args.no_gp and args.config in (2, 3):
        extras['gp_predict'] = gp_predict_line(dev)
        extras['gp_predict']['dgemm_peak_tflops'] = extras['peaks_measured_in_run']['dgemm_tflops']
    if default_value is not None:
        extras['library_default_options'] = default_value

    cpu_baseline = None
    if not args.no_cpu_baseline:
        steps_cpu = {2: 4, 3: 3, 4: 4, 5: 2}[args.config]
        if args.config == 5:
            args.cpu_sample = 1      # the (untimed) cold solve of this problem alone takes the oracle minutes
        v, el, okf, wall = cpu_sample(args.config, args.cpu_sample, steps_cpu, 1)
        cpu_baseline = {'value': v, 'unit': w['unit'], 'cores': 1, 'kind': 'port',
                        'sample': f'first {args.cpu_sample} instances of the same closed loop x {steps_cpu} warm steps '
                                  f'(primal warm start; the cold first solve is not timed), {el:.1f} s, single process; '
                                  f'NumPy/SciPy restatement of the reference\'s flat NLP + IPOPT algorithm',
                        'converged_frac': okf}

    work_mb = loop.eng._t['work'].numel() * 8 / 1e6
    cfg = {'workload': w['workload'], 'config': args.config, 'batch_per_gpu': B, 'parallelism': f'batch-shard x{world}',
           'l2': f'working set {work_mb:.0f} MB per step vs 126 MB L2 (no explicit flush)',
           'hessian': args.hessian, 'warm_start': bool(args.warm_start), 'warm_start_mu': args.warm_mu,
           'warm_start_shift': bool(loop.nmpc._solver_numbers().get('shift_warm_start', 0)) if hasattr(loop, 'nmpc') else True,
           'ipm_iterations_launched_per_step_mean': float(np.mean(launched_hist)),
           'ipm_iterations_launched_by_step': [int(v) for v in launched_hist],
           'ipm_iterations_launched_per_step_e2e_mean': float(np.mean(launched_e2e))}
    cfg.update(stats)
    line = {
        'metric': w['metric'], 'value': value, 'unit': w['unit'], 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': cfg,
        'e2e': {'value': e2e_value, 'unit': w['unit'], 'h2d_bytes_per_step': loop.h2d, 'd2h_bytes_per_step': loop.d2h},
        'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roofline, 'kernels': per_kernel,
        'cpu_baseline': cpu_baseline, 'extras': extras,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
