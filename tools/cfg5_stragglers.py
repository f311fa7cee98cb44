"""Diagnostic: closed loop of config 5 on the device with a synchronise per step; per step the launched iterations, the
distribution of the per-instance iteration counts and the state of the slowest instance.  Dumps the inputs of the worst
solve (instance state + its warm start) to gpurun_out/cfg5_worst.npz for a replay on the CPU harness.
python tools/cfg5_stragglers.py [steps] [batch]"""
import os, sys, json, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device('cuda', 0)
args = types.SimpleNamespace(batch=B, hessian='exact', warm_start=1, warm_mu=1e-9, riccati_variant=None, no_compaction=False,
                             opt=[], warmup=3, steps=steps)
loop = bench.NmpcLoop(5, args, dev, 0)
loop.prewarm(1.)
names = "MU NU ALPHA_P ALPHA_D AMAX_P AMAX_D PHI0 DPHI C1 EDUAL EPRIM ECOMP ACTIVE OBJ NACC THMAX THMIN FILT_N FILT_MU LSP DW".split()
worst = None
rows = []
eng = loop.eng
for k in range(steps):
    t = eng._t
    snap = None
    if t is not None and 'x' in t:
        snap = {nm: t[nm].clone() for nm in ('x', 'u', 'xc', 'lam', 'work') if nm in t}
    x_in = torch.cat([loop.x_cur, loop.theta], dim=1).clone()
    out = loop.step_device(k)
    torch.cuda.synchronize()
    it = out['iters'].cpu().numpy()
    st = out['status'].cpu().numpy()
    b = int(it.argmax())
    stats = out['stats'][b].cpu().numpy() if 'stats' in out else None
    row = {'step': k, 'launched': int(eng.last_launched), 'mean': float(it.mean()), 'p99': float(np.percentile(it, 99)),
           'max': int(it.max()), 'argmax': b, 'status_max': int(st[b]), 'n_gt_1.5mean': int((it > 1.5 * it.mean()).sum()),
           'not_optimal': int((st != 1).sum())}
    if stats is not None:
        row['worst'] = {n: float(stats[i]) for i, n in enumerate(names) if n in ('MU', 'EDUAL', 'EPRIM', 'ECOMP', 'DW', 'FILT_N', 'OBJ')}
    rows.append(row)
    if snap is not None and (worst is None or it.max() > worst[0]):
        worst = (int(it.max()), k, b, x_in[b].cpu().numpy(), {nm: (v[b].cpu().numpy() if nm != 'work' else None) for nm, v in snap.items()})
    if k % 10 == 0 or it.max() > 2 * it.mean():
        print(json.dumps(row), flush=True)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(rows, open('gpurun_out/cfg5_stragglers.json', 'w'))
if worst is not None:
    np.savez('gpurun_out/cfg5_worst.npz', iters=worst[0], step=worst[1], instance=worst[2], x_in=worst[3],
             **{nm: v for nm, v in worst[4].items() if v is not None})
    print('worst', worst[0], 'at step', worst[1], 'instance', worst[2], 'x_in', worst[3])
