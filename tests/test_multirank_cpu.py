"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU force decomposition.

What runs on the GPUs (kernels, NCCL) cannot run here; what CAN be checked without a GPU is the decomposition itself:
 * the term sharding rule the kernels use (work item g belongs to rank g % world) partitions every list exactly once,
 * summing the ranks' partial forces as 2^32 fixed-point int64 (the engine's all-reduce payload) reproduces the
   unsharded result bit-for-bit, independent of the rank count -- the property that lets every rank integrate the same
   trajectory,
 * the bootstrap plumbing bench.py uses (rank 0 creates a 128-byte id, broadcast to all ranks)."""
import os
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

SCALE = 4294967296.0


def _shard(desc, rank, world):
    import copy
    d = copy.copy(desc)
    sel = slice(rank, None, world)
    for names in (("exc_i", "exc_j", "exc_qq", "exc_sigma", "exc_eps"), ("bond_i", "bond_j", "bond_r0", "bond_k"),
                  ("angle_i", "angle_j", "angle_k", "angle_t0", "angle_kk")):
        for n in names:
            setattr(d, n, getattr(desc, n)[sel])
    return d


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from openmm_b200 import systems
    from oracle import port as orc
    desc = systems.water_box(4, cutoff=0.6, rigid=False).rounded()
    L = orc.lib()
    n = desc.natoms
    # bootstrap plumbing: 128-byte id from rank 0
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.arange(128, dtype=torch.uint8)
    dist.broadcast(uid, 0)
    assert uid.tolist() == list(range(128))
    # sharded bonded + exception terms (rank r owns work items r, r+world, ...)
    d = _shard(desc, rank, world)
    f = np.zeros((n, 3))
    pos = orc._d(desc.positions)
    e = 0.0
    e += L.orc_bonds(len(d.bond_i), orc._ip(orc._i(d.bond_i)), orc._ip(orc._i(d.bond_j)), orc._dp(orc._d(d.bond_r0)), orc._dp(orc._d(d.bond_k)), orc._dp(pos), orc._dp(f))
    e += L.orc_angles(len(d.angle_i), orc._ip(orc._i(d.angle_i)), orc._ip(orc._i(d.angle_j)), orc._ip(orc._i(d.angle_k)), orc._dp(orc._d(d.angle_t0)),
                      orc._dp(orc._d(d.angle_kk)), orc._dp(pos), orc._dp(f))
    alpha = desc.pme_parameters()[0]
    box = orc._d(desc.box).reshape(9)
    e += L.orc_exclusion_correction(len(d.exc_i), orc._ip(orc._i(d.exc_i)), orc._ip(orc._i(d.exc_j)), orc._dp(pos), orc._dp(orc._d(desc.charges)),
                                    orc._dp(box), 0, alpha, orc._dp(f))
    fixed = torch.from_numpy(np.rint(f*SCALE).astype(np.int64))
    dist.all_reduce(fixed)                       # int64 sum: exact, order independent
    et = torch.tensor([e], dtype=torch.float64)
    dist.all_reduce(et)
    if rank == 0:
        np.save(out, np.concatenate([fixed.numpy().astype(np.float64).ravel()/SCALE, et.numpy()]))
    dist.destroy_process_group()


def test_force_decomposition_world2(tmp_path):
    from openmm_b200 import systems
    from oracle import port as orc
    out = str(tmp_path/"r.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    desc = systems.water_box(4, cutoff=0.6, rigid=False).rounded()
    L = orc.lib()
    n = desc.natoms
    f = np.zeros((n, 3))
    pos = orc._d(desc.positions)
    e = L.orc_bonds(len(desc.bond_i), orc._ip(orc._i(desc.bond_i)), orc._ip(orc._i(desc.bond_j)), orc._dp(orc._d(desc.bond_r0)), orc._dp(orc._d(desc.bond_k)), orc._dp(pos), orc._dp(f))
    e += L.orc_angles(len(desc.angle_i), orc._ip(orc._i(desc.angle_i)), orc._ip(orc._i(desc.angle_j)), orc._ip(orc._i(desc.angle_k)),
                      orc._dp(orc._d(desc.angle_t0)), orc._dp(orc._d(desc.angle_kk)), orc._dp(pos), orc._dp(f))
    e += L.orc_exclusion_correction(len(desc.exc_i), orc._ip(orc._i(desc.exc_i)), orc._ip(orc._i(desc.exc_j)), orc._dp(pos), orc._dp(orc._d(desc.charges)),
                                    orc._dp(orc._d(desc.box).reshape(9)), 0, desc.pme_parameters()[0], orc._dp(f))
    assert np.abs(got[:-1].reshape(n, 3) - f).max() < 3/SCALE*64        # per-term fixed-point rounding only
    assert abs(got[-1] - e) < 1e-9*abs(e)


def test_round_robin_sharding_is_a_partition():
    for world in (2, 4, 8):
        for n in (0, 1, 7, 1000, 18773):
            seen = np.zeros(n, dtype=int)
            for r in range(world):
                seen[r::world] += 1
            assert (seen == 1).all()
        # contiguous chunks for the PME atoms (k_pme_spread): per = ceil(N/world)
        N = 23558
        per = (N + world - 1)//world
        cover = np.zeros(N, dtype=int)
        for r in range(world):
            cover[r*per:min(N, (r+1)*per)] += 1
        assert (cover == 1).all()


def _role_worker(rank, world, port, out):
    """DESIGN.md section 5 at world 2: rank 0 = direct space + bonded terms + exclusion correction, rank 1 = reciprocal space
    for all atoms (what engine.cu:role_split / role_nb arrange), joined by one int64 all-reduce of the force buffer."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from openmm_b200 import systems
    from oracle import port as orc
    desc = systems.water_box(3, cutoff=0.45, rigid=False).rounded()
    L = orc.lib()
    n = desc.natoms
    pos, q = orc._d(desc.positions), orc._d(desc.charges)
    box = orc._d(desc.box).reshape(9)
    alpha, nx, ny, nz = desc.pme_parameters()
    f = np.zeros((n, 3))
    e = 0.0
    if rank == world - 1:
        e += L.orc_pme_reciprocal(n, orc._dp(pos), orc._dp(q), orc._dp(box), alpha, nx, ny, nz, orc._dp(f))
        e += L.orc_self_energy(n, orc._dp(q), alpha)
    else:
        fd, ed, parts = orc.forces_energy(desc, pme=(alpha, nx, ny, nz))
        # everything but reciprocal space and the self term
        fr = np.zeros((n, 3))
        er = L.orc_pme_reciprocal(n, orc._dp(pos), orc._dp(q), orc._dp(box), alpha, nx, ny, nz, orc._dp(fr))
        f, e = fd - fr, ed - er - parts["self"]
    fixed = torch.from_numpy(np.rint(f*SCALE).astype(np.int64))
    dist.all_reduce(fixed)
    et = torch.tensor([e], dtype=torch.float64)
    dist.all_reduce(et)
    if rank == 0:
        np.save(out, np.concatenate([fixed.numpy().astype(np.float64).ravel()/SCALE, et.numpy()]))
    dist.destroy_process_group()


def test_role_split_direct_rank_plus_reciprocal_rank_world2(tmp_path):
    from openmm_b200 import systems
    from oracle import port as orc
    out = str(tmp_path/"role.npy")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_role_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    desc = systems.water_box(3, cutoff=0.45, rigid=False).rounded()
    f, e, _ = orc.forces_energy(desc)
    assert np.abs(got[:-1].reshape(desc.natoms, 3) - f).max() < 1e-6      # int64 rounding of two partial sums + one subtraction
    assert abs(got[-1] - e) < 1e-9*abs(e)
