"""Timing experiments: per-chunk step time with / without L2 flush and host sync."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from openmm_b200 import systems, Engine

d = systems.water_box(20, cutoff=0.9).rounded()
eng = Engine(d)
eng.set_integrator(systems.INT_LANGEVIN, 0.002, 300.0, 1.0, 7)
stream = torch.cuda.ExternalStream(eng.stream())
flush = torch.empty(256*1024*1024, dtype=torch.uint8, device="cuda")
eng.step(1500); eng.synchronize()

def run(label, nchunk, md, do_flush, sync_each):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(nchunk):
        if do_flush:
            with torch.cuda.stream(stream):
                flush.zero_()
        eng.step(md)
        if sync_each:
            eng.synchronize()
    e1.record(stream)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    print("%-40s device %.1f us/step  host-enqueue %.1f us/step  wall %.1f us/step" % (label, 1e3*ms/(nchunk*md), 1e6*(t1-t0)/(nchunk*md), 1e6*(t2-t0)/(nchunk*md)), flush=True)

for rep in range(2):
    run("5x500 noflush nosync", 5, 500, False, False)
    run("5x500 noflush sync", 5, 500, False, True)
    run("5x500 flush nosync", 5, 500, True, False)
    run("5x500 flush sync", 5, 500, True, True)
    run("25x100 flush nosync", 25, 100, True, False)
    run("1x2500 noflush", 1, 2500, False, False)
os.environ["X"] = "1"
st = eng.stats()
print(st)
