import sys, time, torch, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from hilo_mpc_b200.examples import cstr_initial_states, cstr_nmpc
dev = torch.device('cuda', 0)
B = 4096
nmpc = cstr_nmpc(N=50, solver_options={'b200.warm_start': 1})
eng = nmpc.engine
xs = torch.tensor(nmpc._xs_full, dtype=torch.float64, device=dev)
x0 = torch.tensor(cstr_initial_states(B, seed=0), dtype=torch.float64, device=dev)
num = nmpc._solver_numbers(); dt = nmpc.sampling_interval
gen = torch.Generator(device=dev); gen.manual_seed(1234)
noise = (torch.rand(40, B, 4, generator=gen, device=dev, dtype=torch.float64) * 2 - 1) * torch.tensor(bench.DIST_SCALE, dtype=torch.float64, device=dev)
for rep in range(3):
    eng.reset_guess()
    x = x0
    gpu, wall = [], []
    for k in range(14):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = eng.solve(x / xs, dt=dt, time0=0., **num)
        e1.record()
        x = out['x'][:, 1, :] * xs + noise[k]
        if rep == 1:
            torch.cuda.synchronize()
        wall.append(time.perf_counter() - t0)
        gpu.append((e0, e1))
    torch.cuda.synchronize()
    print('rep', rep, 'sync' if rep == 1 else 'nosync', 'gpu ms', [round(a.elapsed_time(b), 2) for a, b in gpu][3:], 'wall ms', [round(w * 1e3, 2) for w in wall][3:])
