"""Kernel timeline of a few MD steps of the multi-GPU engine, one table per rank (run under torchrun):
    torchrun --nproc-per-node 2 tools/gpu_trace_multi.py apoa1 8
CUPTI activity records via torch.profiler; writes gpurun_out/trace_<name>_x<world>_r<rank>.txt (start_us dur_us stream kernel)."""
import ctypes as C
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
from openmm_b200 import systems, Engine, _lib

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
comm = None
if world > 1:
    dist.init_process_group("gloo")
    os.environ.setdefault("B200MD_NCCL_LIB", os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so.2"))
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = C.create_string_buffer(128)
        assert _lib.load().b200md_comm_unique_id(C.cast(buf, C.c_void_p)) == 0
        uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    dist.broadcast(uid, 0)
    comm = (rank, world, bytes(uid.numpy().tobytes()))
name = sys.argv[1] if len(sys.argv) > 1 else "apoa1"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
d = systems.SystemDesc.load(os.path.join("data", name + ".npz")).rounded()
eng = Engine(d, device=local, comm=comm)
eng.set_integrator(systems.INT_LANGEVIN, 0.002, 300.0, 1.0, 7, 1e-5)
eng.step(600); eng.synchronize()
if world > 1:
    dist.all_reduce(torch.zeros(1))
os.makedirs("gpurun_out", exist_ok=True)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    eng.step(nsteps); eng.synchronize()
path = "gpurun_out/trace_%s_x%d_r%d.json" % (name, world, rank)
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
with open(path.replace(".json", ".txt"), "w") as out:
    out.write("# %s, %d ranks, rank %d: %d MD steps inside the step graphs (CUPTI via torch.profiler). start_us dur_us stream kernel; absolute t0 = %.1f us\n" % (name, world, rank, nsteps, t0))
    for e in ev:
        out.write("%9.1f %7.1f  s%-3s %s\n" % (e["ts"] - t0, e["dur"], e["args"].get("stream", "?"), e["name"][:70]))
os.remove(path)
print("rank", rank, "kernels", len(ev), "per step us", (ev[-1]["ts"] + ev[-1]["dur"] - t0)/nsteps, flush=True)
if world > 1:
    dist.all_reduce(torch.zeros(1))
    dist.destroy_process_group()
