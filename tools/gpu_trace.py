"""Kernel timeline of a few MD steps via torch.profiler (CUPTI activity records see every kernel of the process, graph
nodes included).  Writes gpurun_out/trace_<name>.json (chrome trace) and prints a compact table."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from openmm_b200 import systems, Engine

name = sys.argv[1] if len(sys.argv) > 1 else "dhfr"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
d = systems.SystemDesc.load(os.path.join("data", name + ".npz")).rounded()
eng = Engine(d)
eng.set_integrator(systems.INT_LANGEVIN, 0.002, 300.0, 1.0, 7, 1e-5)
eng.step(1000); eng.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    eng.step(nsteps); eng.synchronize()
path = "gpurun_out/trace_%s.json" % name
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
out = open("gpurun_out/trace_%s.txt" % name, "w")
for e in ev:
    line = "%9.1f %7.1f  s%-3s %s" % (e["ts"] - t0, e["dur"], e["args"].get("stream", "?"), e["name"][:60])
    out.write(line + "\n")
out.close()
print("kernels", len(ev), "span us", ev[-1]["ts"] + ev[-1]["dur"] - t0, "per step", (ev[-1]["ts"] + ev[-1]["dur"] - t0)/nsteps)
