"""CPU tests of the round-2 decomposition rules (tools/dd_model.py): home bricks, halo import, pair ownership, int64 halo
force return.  One pure partition test for 2/4/8 bricks and one real 2-process gloo run."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import dd_model as dd  # noqa: E402

CUTOFF, REACH = 0.9, 0.99


def _single():
    pos, box, q, sig, eps = dd.model_system()
    return dd.forces_of_rank(pos, box, q, sig, eps, CUTOFF, (1, 1, 1), 0, REACH), (pos, box, q, sig, eps)


def test_every_pair_is_owned_exactly_once_for_2_4_8_bricks():
    ref, (pos, box, q, sig, eps) = _single()
    assert np.abs(ref).max() > 0
    for grid in ((2, 1, 1), (2, 2, 1), (2, 2, 2)):
        world = int(np.prod(grid))
        total = np.zeros_like(ref)
        homes = 0
        for r in range(world):
            home, halo, _ = dd.local_sets(pos, box, grid, r, REACH)
            assert len(np.intersect1d(home, halo)) == 0
            homes += len(home)
            total += dd.forces_of_rank(pos, box, q, sig, eps, CUTOFF, grid, r, REACH)
        assert homes == len(pos)                              # the bricks partition the atoms
        assert np.array_equal(total, ref)                     # bit for bit, whatever the decomposition
        assert np.abs(total.sum(axis=0)).max() == 0           # Newton's third law survives the cut


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pos, box, q, sig, eps = dd.model_system()
    f = torch.from_numpy(dd.forces_of_rank(pos, box, q, sig, eps, CUTOFF, (world, 1, 1), rank, REACH))
    dist.all_reduce(f)                                        # stands in for the halo-force return (p2p in the product)
    home, _, _ = dd.local_sets(pos, box, (world, 1, 1), rank, REACH)
    mine = torch.zeros(len(pos), dtype=torch.int64)
    mine[torch.from_numpy(home)] = 1
    dist.all_reduce(mine)
    if rank == 0:
        np.save(out, np.concatenate([f.numpy().ravel(), mine.numpy()]))
    dist.destroy_process_group()


def test_two_ranks_reproduce_the_single_rank_forces_bit_for_bit(tmp_path):
    out = str(tmp_path/"dd.npy")
    mp.spawn(_worker, args=(2, 33500 + (os.getpid() % 2000), out), nprocs=2, join=True)
    got = np.load(out)
    ref, (pos, *_rest) = _single()
    n = len(pos)
    assert np.array_equal(got[:3*n].reshape(n, 3), ref)
    assert (got[3*n:] == 1).all()                             # every atom has exactly one home rank
