"""Minimal step driver for profiling (ncu wraps this): python tools/gpu_steps.py <workload> <md steps> [equilibration steps]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from openmm_b200 import systems, Engine

name = sys.argv[1] if len(sys.argv) > 1 else "dhfr"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
eq = int(sys.argv[3]) if len(sys.argv) > 3 else 200
d = bench.load_workload(name)
eng = Engine(d)
eng.set_integrator(systems.INT_LANGEVIN, 0.002, 300.0, 1.0, 7, 1e-5)
eng.step(eq)
eng.synchronize()
eng.step(n)
eng.synchronize()
print("done", name, eng.stats()["num_tiles"])
