"""world_size-2 gloo tests (CPU) of the N>1 host logic: batch sharding, GP training-shard all-gather, result gather
and the max-over-ranks timing reduction used by bench.py."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from hilo_mpc_b200 import distributed as hd


def test_shard_bounds_cover_the_batch_exactly():
    for total in (0, 1, 7, 4096, 16385):
        for world in (1, 2, 3, 8):
            pieces = [hd.shard_bounds(total, world, r) for r in range(world)]
            assert pieces[0][0] == 0 and pieces[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
            sizes = [hi - lo for lo, hi in pieces]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        hd.shard_bounds(10, 2, 2)


def _worker(rank, world, port, tmp):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    r, w, _ = hd.init_process_group('gloo')
    assert (r, w) == (rank, world)
    rng = np.random.default_rng(0)
    X = rng.normal(size=(3, 101))
    y = rng.normal(size=(1, 101))
    lo, hi = hd.shard_bounds(101, world, rank)
    Xf, yf = hd.all_gather_training_shards(X[:, lo:hi], y[:, lo:hi])
    np.testing.assert_array_equal(Xf.numpy(), X)
    np.testing.assert_array_equal(yf.numpy(), y)
    batch = torch.arange(13 * 2, dtype=torch.float64).reshape(13, 2)
    mine = hd.shard(batch)
    full = hd.gather_to_rank0(mine * 2)
    if rank == 0:
        assert torch.equal(full, batch * 2)
    else:
        assert full is None
    assert hd.max_over_ranks(float(rank + 1)) == float(world)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
    open(os.path.join(tmp, f'ok{rank}'), 'w').write('ok')


def test_two_rank_gloo(tmp_path):
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / 'ok0').exists() and (tmp_path / 'ok1').exists()
