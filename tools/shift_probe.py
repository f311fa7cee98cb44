"""Probe: interior-point iterations per closed-loop step with the reference's un-shifted warm start (mpc.py:725-726) vs
a warm start shifted by one stage (primal only / primal + bound multipliers).  python tools/shift_probe.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from hilo_mpc_b200.examples import cstr_initial_states, cstr_nmpc
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device('cuda', 0)
N, NX, NU, NC = 50, 4, 2, 12


def run(mode):
    nmpc = cstr_nmpc(N=N, solver_options={'b200.warm_start': 1, 'b200.warm_start_mu': 1e-8})
    eng = nmpc.engine
    xs = torch.tensor(nmpc._xs_full, dtype=torch.float64, device=dev)
    x0 = torch.tensor(cstr_initial_states(B, seed=0), dtype=torch.float64, device=dev)
    gen = torch.Generator(device=dev); gen.manual_seed(1234)
    noise = (torch.rand(64, B, NX, generator=gen, device=dev, dtype=torch.float64) * 2 - 1) * \
        torch.tensor(bench.DIST_SCALE, dtype=torch.float64, device=dev)
    num = nmpc._solver_numbers()
    x_cur = x0
    hist = []
    for k in range(14):
        if k > 0 and mode > 0:
            t = eng._t
            for nm in ('x', 'u', 'xc', 'lam'):
                a = t[nm]
                a[:, :-1] = a[:, 1:].clone()
            if mode > 1:
                w = t['work']
                off = 0
                def take(cnt):
                    global _o
                    return None
                sizes = [B * (N + 1) * NX] * 2 + [B * N * NU] * 2 + [B * N * NC] * 2
                shapes = [(B, N + 1, NX)] * 2 + [(B, N, NU)] * 2 + [(B, N, NC)] * 2
                base = ((128 - (w.data_ptr() & 127)) & 127) // 8
                off = base
                for sz, sh in zip(sizes, shapes):
                    z = w[off:off + sz].view(*sh)
                    z[:, :-1] = z[:, 1:].clone()
                    off += (sz + 15) & ~15
        out = eng.solve(x_cur / xs, dt=nmpc.sampling_interval, time0=0., **num)
        x_cur = out['x'][:, 1, :] * xs + noise[k]
        it = out['iters'].cpu().numpy()
        hist.append((eng.last_launched, float(it.mean()), int(it.max()), np.bincount(it, minlength=8)[:8].tolist(),
                     float((out['status'] == 1).double().mean().item())))
    return hist


for mode, name in ((0, 'un-shifted (reference)'), (1, 'primal shifted'), (2, 'primal + multipliers shifted')):
    h = run(mode)
    print(name)
    for k, r in enumerate(h):
        if k >= 4:
            print('  step %2d launched %d mean %.2f max %d hist %s ok %.3f' % ((k,) + r))
