"""Run under torchrun on N GPUs (N >= 2): python -m torch.distributed.run --nproc-per-node N tools/multigpu_check.py
1. GP training shards are all-gathered over NCCL (the one collective of the path) and every rank's fit /
   prediction equals the fit on the full training set.
2. A batch of NMPC problems sharded over the ranks (no data-path collective) gives, gathered on rank 0, exactly the
   inputs a single rank computes for the whole batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from hilo_mpc_b200 import distributed as hd
from hilo_mpc_b200.gp import GaussianProcess, Kernel
from hilo_mpc_b200.examples import cstr_initial_states, cstr_nmpc


def main():
    rank, world, local = hd.init_process_group('nccl')
    dev = torch.device('cuda', local)
    # ---- 1. GP
    rng = np.random.default_rng(7)
    n, d = 515, 3
    X = rng.uniform(-1, 1, size=(d, n)); y = np.sin(X.sum(0, keepdims=True))
    lo, hi = hd.shard_bounds(n, world, rank)
    gp = GaussianProcess(['a', 'b', 'c'], 'y', kernel=Kernel.squared_exponential(length_scales=.7), noise_variance=1e-3)
    gp.set_training_shards(torch.tensor(X[:, lo:hi], device=dev), torch.tensor(y[:, lo:hi], device=dev))
    gp.setup()
    ref = GaussianProcess(['a', 'b', 'c'], 'y', kernel=Kernel.squared_exponential(length_scales=.7), noise_variance=1e-3)
    ref.set_training_data(X, y); ref.setup()
    Xq = rng.uniform(-1, 1, size=(d, 64))
    m1, v1 = gp.predict(Xq); m2, v2 = ref.predict(Xq)
    assert np.array_equal(m1, m2) and np.array_equal(v1, v2), 'sharded GP differs from the full fit'
    assert abs(gp.log_marginal_likelihood() - ref.log_marginal_likelihood()) == 0.0
    # ---- 2. sharded NMPC batch
    Bt = 64 * world + 3
    x0 = cstr_initial_states(Bt, seed=11)
    nmpc = cstr_nmpc(N=10)
    lo, hi = hd.shard_bounds(Bt, world, rank)
    u_mine = nmpc.optimize(torch.tensor(x0[lo:hi], device=dev))
    assert (nmpc.solver_status_code == 1).all()
    u_all = hd.gather_to_rank0(u_mine)
    if rank == 0:
        nm2 = cstr_nmpc(N=10)
        u_full = nm2.optimize(torch.tensor(x0, device=dev))
        assert torch.equal(u_all.to(dev), u_full), 'sharded solve differs from the single-rank solve'
        print(f'[multigpu_check] world={world}: GP all-gather fit identical, sharded NMPC batch ({Bt}) identical')
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
