"""2..8-rank smoke test of the force decomposition (run under torchrun): forces and a short trajectory must equal the
single-GPU engine's."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import torch, torch.distributed as dist
from openmm_b200 import systems, Engine, _lib

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
def log(*a):
    print("[rank %d %.1fs]" % (rank, time.time()-t0), *a, flush=True)
t0 = time.time()
torch.cuda.set_device(local)
dist.init_process_group("gloo")
os.environ.setdefault("B200MD_NCCL_LIB", os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so.2"))
uid = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    buf = C.create_string_buffer(128)
    assert _lib.load().b200md_comm_unique_id(C.cast(buf, C.c_void_p)) == 0
    uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
dist.broadcast(uid, 0)
log("uid broadcast done")
name = sys.argv[1] if len(sys.argv) > 1 else "water"
d = systems.water_box(10, cutoff=0.9).rounded() if name == "water" else systems.SystemDesc.load(os.path.join("data", name + ".npz")).rounded()
ref = Engine(d, device=local)
eref = ref.compute(); fref = ref.get_forces()
log("single-GPU reference computed", eref)
eng = Engine(d, device=local, comm=(rank, world, bytes(uid.numpy().tobytes())))
log("comm engine created")
e = eng.compute(); f = eng.get_forces()
log("decomposed compute: E %.6f vs %.6f  max|dF| %.3e" % (e, eref, np.abs(f-fref).max()))
for g in (ref, eng):
    g.set_integrator(systems.INT_LANGEVIN, 0.002, 300.0, 1.0, 5)
ref.step(40); eng.step(40)
x = eng.get_positions(); xr = ref.get_positions()
log("40 steps: max|dx| %.3e" % np.abs(x-xr).max())
torch.cuda.synchronize(); dist.all_reduce(torch.zeros(1))
for nm, g in (("single", ref), ("x%d" % world, eng)):
    g.step(200); g.synchronize()
    t = time.time(); g.step(1000); g.synchronize(); dt = time.time()-t
    log("%s: %.1f us/step" % (nm, 1e3*dt))
dist.all_reduce(torch.zeros(1))
log("done")
dist.destroy_process_group()
