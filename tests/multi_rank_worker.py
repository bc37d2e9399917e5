"""Worker of tests/test_gpu_multi.py (run under torchrun, one process per GPU): the multi-GPU engine (peer-memory data
plane: owner decomposition, NVLink stores, slab-decomposed PME) against the single-GPU engine on the same device.
  * forces and energy of one evaluation: bit-equal forces (every contribution is an int64 fixed-point sum), energy 1e-12
  * a short Langevin trajectory: same positions to 1e-5 nm (the centre-of-mass sums are reduced in a different order)
  * the state read back from every rank is the same
Prints one line `MULTI_OK ...` on rank 0 when everything holds."""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import numpy as np
import torch
import torch.distributed as dist
from openmm_b200 import systems, Engine, _lib

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
t0 = time.time()


def log(*a):
    print("[rank %d %.1fs]" % (rank, time.time() - t0), *a, flush=True)


torch.cuda.set_device(local)
dist.init_process_group("gloo")
os.environ.setdefault("B200MD_NCCL_LIB", os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so.2"))
uid = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    buf = C.create_string_buffer(128)
    assert _lib.load().b200md_comm_unique_id(C.cast(buf, C.c_void_p)) == 0
    uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
dist.broadcast(uid, 0)
name = sys.argv[1] if len(sys.argv) > 1 else "water"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
if name == "water":
    d = systems.water_box(10, cutoff=0.9).rounded()
else:
    d = systems.SystemDesc.load(os.path.join(ROOT, "data", name + ".npz")).rounded()
ref = Engine(d, device=local)
eref = ref.compute()
fref = ref.get_forces()
log("single-GPU engine: E = %.6f" % eref)
eng = Engine(d, device=local, comm=(rank, world, bytes(uid.numpy().tobytes())))
e = eng.compute()
f = eng.get_forces()
dfmax = float(np.abs(f - fref).max())
log("x%d engine: E = %.6f, max|dF| = %.3e" % (world, e, dfmax))
ok = dfmax == 0.0 and abs(e - eref) <= 1e-8*abs(eref)      # the energy is a double sum of fp32 terms reduced in another order
# a second evaluation (exercises the exchange epochs) and component evaluations
for terms in (31, 8 | 1 | 2 | 4, 16):
    e2 = eng.compute(terms); f2 = eng.get_forces()
    er2 = ref.compute(terms); fr2 = ref.get_forces()
    ok = ok and float(np.abs(f2 - fr2).max()) == 0.0 and abs(e2 - er2) <= 1e-8*max(1.0, abs(er2))
v0 = np.random.default_rng(5).standard_normal((d.natoms, 3))*0.3
for g in (ref, eng):
    g.set_velocities(v0)
    g.set_integrator(systems.INT_LANGEVIN, 0.002, 300.0, 1.0, 5)
ref.step(nsteps)
eng.step(nsteps)
x = eng.get_positions(); xr = ref.get_positions()
v = eng.get_velocities(); vr = ref.get_velocities()
dx = float(np.abs(x - xr).max()); dv = float(np.abs(v - vr).max())
log("%d steps: max|dx| = %.3e nm, max|dv| = %.3e nm/ps" % (nsteps, dx, dv))
ok = ok and dx < 1e-5 and dv < 2e-3 and np.isfinite(x).all()
ke = eng.kinetic_energy(); ker = ref.kinetic_energy()
ok = ok and abs(ke - ker) < 1e-4*abs(ker)
e3 = eng.compute(); e3r = ref.compute()
ok = ok and abs(e3 - e3r) < 1e-5*abs(e3r)
# every rank must hold the same state
xs = torch.from_numpy(x.copy()); xall = [torch.zeros_like(xs) for _ in range(world)]
dist.all_gather(xall, xs)
ok = ok and all(bool((xa == xall[0]).all()) for xa in xall)
st = eng.stats()
ok = ok and st["overflow"] == 0
flag = torch.tensor([1.0 if ok else 0.0]); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
torch.cuda.synchronize(); dist.all_reduce(torch.zeros(1))
times = {}
for nm, g in (("x1", ref), ("x%d" % world, eng)):
    g.step(200); g.synchronize()
    dist.all_reduce(torch.zeros(1))
    t = time.time(); g.step(1000); g.synchronize(); times[nm] = 1e3*(time.time() - t)
    dist.all_reduce(torch.zeros(1))
if rank == 0:
    print("%s %s world=%d dF=%.3e dx=%.3e us_per_step=%s" % ("MULTI_OK" if flag.item() == 1.0 else "MULTI_FAIL", name, world, dfmax, dx,
                                                             {k: round(val, 1) for k, val in times.items()}), flush=True)
dist.all_reduce(torch.zeros(1))
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1.0 else 1)
