"""First-contact diagnostic on a B200: component-by-component comparison against the oracle, prints everything."""
import sys, os, time, json, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openmm_b200 import systems, Engine
from openmm_b200.engine import fft3d_r2c, fft3d_c2r, TERM_ALL, TERM_NB_DIRECT, TERM_NB_RECIP, TERM_BONDS, TERM_ANGLES
from oracle import omm


def relerr(f, fr):
    d = np.abs(f - fr).max(axis=1)
    n = np.maximum(1.0, np.linalg.norm(fr, axis=1))
    return float((d/n).max())


def section(name):
    print("\n==== " + name, flush=True)


def run(name, fn):
    section(name)
    try:
        fn()
    except Exception:
        traceback.print_exc()


def t_fft():
    rng = np.random.default_rng(0)
    for shape in [(8, 8, 8), (12, 10, 14), (28, 25, 30), (21, 25, 27), (56, 56, 56), (88, 88, 88), (90, 90, 90), (128, 128, 128)]:
        x = rng.standard_normal(shape).astype(np.float32)
        ref = np.fft.rfftn(x.astype(np.float64))
        out = fft3d_r2c(x)
        e1 = np.abs(out - ref).max()/np.abs(ref).max()
        back = fft3d_c2r(ref.astype(np.complex64), shape[2])
        e2 = np.abs(back/np.prod(shape) - x).max()
        print(shape, "fwd rel err %.2e  inv abs err %.2e" % (e1, e2), flush=True)


def compare(desc, label, terms=TERM_ALL, groups=-1, pme=None):
    eng = Engine(desc)
    e = eng.compute(terms)
    f = eng.get_forces()
    st = eng.stats()
    if pme is None and desc.method == systems.NB_PME:
        pme = desc.pme_parameters()
    sim = omm.Simulation(desc, "Reference", pme=pme)
    fr, er = sim.forces_energy()
    print(label, "N=%d E=%.6f Eref=%.6f relE=%.2e relF=%.2e maxF=%.1f" % (desc.natoms, e, er, abs(e-er)/max(1, abs(er)), relerr(f, fr), np.abs(fr).max()), flush=True)
    print("   stats", {k: st[k] for k in ("num_tiles", "num_mask_tiles", "pairs_in_cutoff", "list_builds", "overflow")}, flush=True)
    worst = np.argmax(np.abs(f-fr).max(axis=1)/np.maximum(1, np.linalg.norm(fr, axis=1)))
    print("   worst atom", worst, f[worst], fr[worst], flush=True)
    return eng, sim


def t_nocutoff():
    compare(systems.cluster(70).rounded(), "nocutoff")
    compare(systems.cluster(500, method=systems.NB_CUTOFF_NONPERIODIC).rounded(), "cutoff nonperiodic")


def t_lj():
    compare(systems.lj_fluid(8, cutoff=1.0), "lj cutoff periodic")
    compare(systems.lj_fluid(8, cutoff=1.0, charged=True), "lj+q RF periodic")


def t_water_small():
    d = systems.water_box(7, cutoff=0.9).rounded()
    compare(d, "water7 pme all")
    d2 = systems.water_box(7, cutoff=0.9, rigid=False).rounded()
    compare(d2, "water7 flexible (bonds+angles)")


def t_ions():
    compare(systems.random_ions(894, 3.0, cutoff=1.0).rounded(), "ions cubic")
    compare(systems.random_ions(894, 3.0, cutoff=1.0, triclinic=True).rounded(), "ions triclinic")


def t_water_big():
    d = systems.water_box(20, cutoff=0.9).rounded()
    t = time.time()
    eng, sim = compare(d, "water20 pme")
    print("   wall", time.time()-t)
    for ph in ["list_build", "pair", "pme_spread", "pme_fft_conv", "pme_gather", "bonded"]:
        print("   phase %-12s %.4f ms" % (ph, eng.time_phase(ph, 20)), flush=True)
    eng.set_integrator(systems.INT_LANGEVIN, 0.002, 300.0, 1.0, 7)
    eng.set_velocities(np.zeros((d.natoms, 3)))
    for graph in (0, 1):
        eng.step(20); eng.synchronize()
        t = time.time(); eng.step(500); eng.synchronize(); dt = time.time()-t
        print("   500 steps: %.3f s -> %.1f ns/day (2 fs)" % (dt, 0.002*500*86400/dt/1000), flush=True)
    print("   phases while hot:", {ph: round(eng.time_phase(ph, 20)*1000, 1) for ph in ["pair", "pme_spread", "pme_fft_conv", "pme_gather", "bonded"]}, "us")
    e = eng.compute(); ke = eng.kinetic_energy()
    print("   after 1040 steps: PE %.1f KE %.1f  T=%.1f K" % (e, ke, 2*ke/(0.00831446*(3*d.natoms - len(d.con_i)))), eng.stats(), flush=True)
    x = eng.get_positions()
    o = x[0::3]; h1 = x[1::3]; h2 = x[2::3]
    print("   constraint check: OH %.6f..%.6f HH %.6f..%.6f" % (np.linalg.norm(o-h1, axis=1).min(), np.linalg.norm(o-h2, axis=1).max(),
                                                                np.linalg.norm(h1-h2, axis=1).min(), np.linalg.norm(h1-h2, axis=1).max()))


def t_verlet():
    d = systems.water_box(6, cutoff=0.9).rounded()
    rng = np.random.default_rng(3)
    v = rng.standard_normal((d.natoms, 3))*0.3
    for kind, name in ((systems.INT_VERLET, "verlet"), (systems.INT_LANGEVIN, "langevin T=0"), (systems.INT_LANGEVIN_MIDDLE, "middle T=0")):
        eng = Engine(d)
        eng.set_integrator(kind, 0.001, 0.0, 1.0, 7)
        sim = omm.Simulation(d, "Reference", integrator=(kind, 0.0, 1.0, 0.001), pme=d.pme_parameters())
        eng.set_velocities(v); sim.set_velocities(v)
        eng.apply_constraints(); eng.apply_velocity_constraints()
        eng.step(10); sim.step(10)
        x = eng.get_positions(); xr = sim.state(positions=True, velocities=True)
        eng.compute()
        print(name, "pos err %.2e vel err %.2e  KE %.4f vs %.4f" % (np.abs(x-xr["positions"]).max(), np.abs(eng.get_velocities()-xr["velocities"]).max(),
                                                                   eng.kinetic_energy(), sim.state(energy=True)["kinetic"]), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["fft", "nocutoff", "lj", "water_small", "ions", "verlet", "water_big"]
    for w in which:
        run(w, globals()["t_" + w])
