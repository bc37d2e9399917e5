"""GPU parity tests (pytest -m gpu): the CUDA hot path, called through the C-ABI, against
 (a) the reference itself (oracle/_ref: the unmodified Reference platform, double precision) on identical inputs,
 (b) the committed golden fixtures (Gromacs known answers held by the reference's own tests),
 (c) the plain-C oracle (oracle/md_oracle.c),
 (d) size-independent properties at the benchmark size (momentum conservation, constraint satisfaction, energy/force
     consistency by finite differences, determinism).
Tolerance: 1e-4 relative in the reference's own ASSERT_EQUAL_VEC form (BASELINE.json north_star)."""
import json
import os
import numpy as np
import pytest
from conftest import relative_force_error, GOLDEN

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def mods():
    from openmm_b200 import systems, Engine, engine
    from oracle import omm, port
    return systems, Engine, engine, omm, port


def _compare(mods, desc, pme=None, tol=TOL, etol=TOL):
    systems, Engine, engine, omm, port = mods
    eng = Engine(desc)
    e = eng.compute()
    f = eng.get_forces()
    if pme is None and desc.method == systems.NB_PME:
        pme = desc.pme_parameters()
    sim = omm.Simulation(desc, "Reference", pme=pme)
    fr, er = sim.forces_energy()
    assert relative_force_error(f, fr) < tol
    assert abs(e - er)/max(1.0, abs(er)) < etol
    assert eng.stats()["overflow"] == 0
    return eng, sim


# ---- the bespoke FFT against numpy (template: platforms/cuda/tests/TestCudaFFT3D.cpp:52-135, same odd sizes) ----
@pytest.mark.parametrize("shape", [(28, 25, 30), (28, 25, 25), (25, 28, 25), (25, 25, 28), (21, 25, 27), (56, 56, 56), (88, 88, 88), (90, 90, 90), (128, 128, 128), (6, 6, 6),
                                   (26, 39, 13), (27, 45, 33), (12, 40, 24), (144, 20, 36)])
def test_fft3d_matches_numpy(mods, shape):
    _, _, engine, _, _ = mods
    rng = np.random.default_rng(sum(shape))
    x = rng.standard_normal(shape).astype(np.float32)
    ref = np.fft.rfftn(x.astype(np.float64))
    out = engine.fft3d_r2c(x)
    assert np.abs(out - ref).max()/np.abs(ref).max() < 2e-6
    back = engine.fft3d_c2r(ref.astype(np.complex64), shape[2])
    assert np.abs(back/np.prod(shape) - x).max() < 1e-5        # unnormalised round trip scales by N (fftpack.h:80-92)


def test_fft_rejects_unsupported_size(mods):
    _, _, engine, _, _ = mods
    with pytest.raises(engine.EngineError):
        engine.fft3d_r2c(np.zeros((34, 8, 8), np.float32))      # 17 is not a supported radix


# ---- golden vectors held by the reference's own tests ----
def test_gromacs_triclinic_golden(mods):
    systems, Engine, *_ = mods
    g = json.load(open(os.path.join(GOLDEN, "ewald_triclinic_gromacs.json")))
    d = systems.SystemDesc(masses=np.ones(8), charges=np.array(g["charges"]), sigmas=np.array(g["sigmas"]), epsilons=np.array(g["epsilons"]),
                           positions=np.array(g["positions"]), box=np.array(g["box"]), method=systems.NB_PME, cutoff=g["cutoff"],
                           pme_alpha=g["alpha"], pme_grid=tuple(g["grid"]), use_dispersion=False)
    eng = Engine(d)
    e = eng.compute()
    assert relative_force_error(eng.get_forces(), np.array(g["expected_forces"])) < g["tolerance"]     # TestEwald.h:268
    assert abs(e - g["expected_energy"])/abs(g["expected_energy"]) < g["tolerance"]


def test_nacl_amorph_fixture(mods):
    systems, Engine, *_ = mods
    z = np.load(os.path.join(GOLDEN, "nacl_amorph.npz"))
    n, L = 894, float(z["box"])
    pme = z["pme"]
    d = systems.SystemDesc(masses=np.ones(n), charges=z["charges"], sigmas=np.ones(n), epsilons=np.zeros(n), positions=z["positions"],
                           box=np.diag([L, L, L]), method=systems.NB_PME, cutoff=float(z["cutoff"]),
                           pme_alpha=float(pme[0]), pme_grid=(int(pme[1]), int(pme[2]), int(pme[3])))
    eng = Engine(d)
    e = eng.compute()
    # +-1 e ions, 1.2 nm cutoff, 1e-5 Ewald tolerance: every atom sums ~240 pair forces of up to 2000 kJ/mol/nm that cancel
    # to ~400 (the reference's own CUDA-vs-Reference tolerance for this system is 1e-2, TestEwald.h:147-149).  Round 1 held
    # 1.1e-4 here; with the close pairs in double and double PME weights: 4.6e-5 (profiles/r02_parity_probe.md).
    assert relative_force_error(eng.get_forces(), z["reference_forces"]) < TOL
    assert abs(e - float(z["reference_energy"]))/abs(e) < 1e-5           # TestEwald.h:147-149 asks 1e-5 on the energy
    assert abs(e - float(z["gromacs_energy"]))/abs(e) < 1e-5


def test_water5_fixture_and_c_oracle(mods):
    systems, Engine, _, _, port = mods
    z = np.load(os.path.join(GOLDEN, "water5_reference.npz"))
    d = systems.water_box(5, cutoff=0.75).rounded()
    eng = Engine(d)
    e = eng.compute()
    f = eng.get_forces()
    assert relative_force_error(f, z["reference_forces"]) < TOL
    assert abs(e - float(z["reference_energy"]))/abs(e) < TOL
    fp, ep, _ = port.forces_energy(d)
    assert relative_force_error(f, fp) < TOL and abs(e - ep)/abs(ep) < TOL


# ---- against the reference itself on identical (fp32-representable) inputs ----
def test_nocutoff_cluster(mods):
    _compare(mods, mods[0].cluster(70).rounded())


def test_cutoff_nonperiodic(mods):
    _compare(mods, mods[0].cluster(500, method=mods[0].NB_CUTOFF_NONPERIODIC).rounded())


def test_cutoff_periodic_lj_and_reaction_field(mods):
    systems = mods[0]
    _compare(mods, systems.lj_fluid(8, cutoff=1.0).rounded())
    _compare(mods, systems.lj_fluid(8, cutoff=1.0, charged=True).rounded())


def test_switching_function_and_dispersion(mods):
    systems = mods[0]
    d = systems.lj_fluid(8, cutoff=1.0).rounded()
    d.use_switch, d.switch_distance = True, 0.8
    _compare(mods, d)


def test_pme_water_rigid_and_flexible(mods):
    systems = mods[0]
    _compare(mods, systems.water_box(7, cutoff=0.9).rounded())
    _compare(mods, systems.water_box(7, cutoff=0.9, rigid=False).rounded())


def test_pme_ions_cubic_and_triclinic(mods):
    systems = mods[0]
    _compare(mods, systems.random_ions(894, 3.0, cutoff=1.0).rounded())
    _compare(mods, systems.random_ions(894, 3.0, cutoff=1.0, triclinic=True).rounded())


def test_pme_grid_with_radix_11(mods):
    systems = mods[0]
    d = systems.water_box(7, cutoff=0.9).rounded()
    d.pme_alpha, d.pme_grid = d.pme_parameters()[0], (22, 22, 22)
    _compare(mods, d)


def test_exceptions_14_and_torsions(mods):
    systems = mods[0]
    rng = np.random.default_rng(5)
    d = systems.water_box(5, cutoff=0.75, rigid=False).rounded()
    n = d.natoms
    # a handful of artificial 1-4 exceptions with their own parameters + torsions over consecutive atoms
    o = 3*np.arange(20, dtype=np.int32)
    d.exc_i = np.concatenate([d.exc_i, o]); d.exc_j = np.concatenate([d.exc_j, o+3])
    d.exc_qq = np.concatenate([d.exc_qq, np.full(20, 0.2)]); d.exc_sigma = np.concatenate([d.exc_sigma, np.full(20, 0.3)])
    d.exc_eps = np.concatenate([d.exc_eps, np.full(20, 0.5)])
    d.tor_i, d.tor_j, d.tor_k, d.tor_l = o, o+1, o+3, o+4
    d.tor_n = rng.integers(1, 4, 20).astype(np.int32); d.tor_phase = rng.random(20)*3; d.tor_kk = rng.random(20)*10
    assert n > 70
    _compare(mods, d)


def test_benchmark_size_water_parity(mods):
    """S1 of SURVEY.md 8(d): 24,000 atoms, 56^3 grid -- parity at the size bench.py runs."""
    eng, sim = _compare(mods, mods[0].water_box(20, cutoff=0.9).rounded())
    st = eng.stats()
    assert st["num_tiles"] > 0 and st["pairs_in_cutoff"] > 3e6


@pytest.mark.parametrize("name", ["dhfr", "apoa1"])
def test_real_benchmark_systems_parity(mods, name):
    """The REAL BASELINE.json systems (data/*.npz, built by the reference's own forcefield.py: tools/make_benchmark_systems.py):
    DHFR 23,558 atoms amber99sb/tip3p PME 0.9 nm 56^3; ApoA1 92,224 atoms ff14SB/lipid17/tip3p PME 1.0 nm 88^3."""
    from conftest import ROOT
    systems = mods[0]
    d = systems.SystemDesc.load(os.path.join(ROOT, "data", name + ".npz")).rounded()
    # Both at 1e-4 in the floor-1 relative measure of ASSERT_EQUAL_VEC against the TOTAL Reference force.  ApoA1 is the hard
    # one (atoms whose ~1 kJ/mol/nm net force is the difference of ~100 kJ/mol/nm direct and reciprocal sums): round 1 had 57
    # of 92,224 atoms above 1e-4 (max 7.5e-4); DESIGN.md section 4 "Precision" says what closed it (9.6e-5, 0 atoms).
    eng, sim = _compare(mods, d, tol=TOL)
    st = eng.stats()
    assert st["pme_grid"] == ([56, 56, 56] if name == "dhfr" else [88, 88, 88])
    # a short constrained Langevin run keeps every HBonds constraint (SETTLE waters + X-H_n SHAKE clusters)
    eng.set_integrator(systems.INT_LANGEVIN, 0.002, 300.0, 1.0, 1, 1e-5)
    eng.step(50)
    x = eng.get_positions()
    sel = slice(None, None, 97)
    dist = np.linalg.norm(x[d.con_i[sel]] - x[d.con_j[sel]], axis=1)
    assert np.abs(dist - d.con_d[sel]).max() < 2e-5
    assert np.isfinite(x).all()


# ---- integrators and constraints ----
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_deterministic_integration_matches_reference(mods, kind):
    """Verlet, and Langevin / LangevinMiddle at T = 0 (no noise term, ReferenceStochasticDynamics.cpp:99-100) are
    deterministic: 10 steps must follow the Reference trajectory."""
    systems, Engine, _, omm, _ = mods
    d = systems.water_box(6, cutoff=0.9).rounded()
    v = np.random.default_rng(3).standard_normal((d.natoms, 3))*0.3
    eng = Engine(d)
    eng.set_integrator(kind, 0.001, 0.0, 1.0, 7)
    sim = omm.Simulation(d, "Reference", integrator=(kind, 0.0, 1.0, 0.001), pme=d.pme_parameters())
    eng.set_velocities(v); sim.set_velocities(v)
    eng.apply_velocity_constraints()
    eng.step(10); sim.step(10)
    ref = sim.state(positions=True, velocities=True, energy=True)
    assert np.abs(eng.get_positions() - ref["positions"]).max() < 5e-6
    assert np.abs(eng.get_velocities() - ref["velocities"]).max() < 1e-3      # fp32 positions: dx/dt noise ~1e-7/1e-3
    eng.compute()
    assert abs(eng.kinetic_energy() - ref["kinetic"])/ref["kinetic"] < 1e-3


def test_settle_constraints_hold_over_langevin_run(mods):
    """tests/TestSettle.h:45-98: constraint lengths within 1e-5 (relative to fp32 positions here: 2e-6 nm)."""
    systems, Engine, *_ = mods
    d = systems.water_box(8, cutoff=0.9).rounded()
    eng = Engine(d)
    eng.set_integrator(systems.INT_LANGEVIN, 0.002, 300.0, 2.0, 11)
    eng.step(500)
    x = eng.get_positions()
    for i, j, dist in zip(d.con_i[::7], d.con_j[::7], d.con_d[::7]):
        assert abs(np.linalg.norm(x[i]-x[j]) - dist) < 1e-5
    assert np.isfinite(x).all()


def test_langevin_temperature(mods):
    """tests/TestLangevinIntegrator.h:92 (statistical): the thermostat reaches the target temperature."""
    systems, Engine, *_ = mods
    d = systems.water_box(8, cutoff=0.9).rounded()
    eng = Engine(d)
    eng.set_integrator(systems.INT_LANGEVIN_MIDDLE, 0.002, 300.0, 5.0, 3)
    eng.step(1500)
    ke = []
    for _ in range(20):
        eng.step(50)
        eng.compute(energy=False)
        ke.append(eng.kinetic_energy())
    dof = 3*d.natoms - len(d.con_i)
    T = 2*np.mean(ke)/(dof*0.00831446261815324)
    assert abs(T - 300.0) < 12.0


def test_random_seed_reproducibility(mods):
    """tests/TestLangevinIntegrator.h:205: equal seeds give identical trajectories, different seeds do not."""
    systems, Engine, *_ = mods
    d = systems.water_box(5, cutoff=0.75).rounded()
    out = []
    for seed in (5, 5, 6):
        eng = Engine(d)
        eng.set_integrator(systems.INT_LANGEVIN, 0.002, 300.0, 1.0, seed)
        eng.step(25)
        out.append(eng.get_positions())
    assert np.array_equal(out[0], out[1])
    assert np.abs(out[0] - out[2]).max() > 1e-4


def test_energy_force_consistency_finite_difference(mods):
    """tests/TestEwald.h:160-178: E(x + h n) - E(x - h n) = -2h |F| along the force direction."""
    systems, Engine, *_ = mods
    d = systems.random_ions(300, 2.2, cutoff=1.0, seed=4).rounded()
    eng = Engine(d)
    eng.compute()
    f = eng.get_forces()
    norm = np.sqrt((f**2).sum())
    h = 1e-3
    x0 = d.positions
    eng.set_positions(x0 + 0.5*h*f/norm)
    e1 = eng.compute()
    eng.set_positions(x0 - 0.5*h*f/norm)
    e2 = eng.compute()
    assert abs((e2 - e1)/h - norm)/norm < 2e-2


def test_momentum_conservation_and_determinism_at_benchmark_size(mods):
    systems, Engine, *_ = mods
    d = systems.water_box(20, cutoff=0.9).rounded()
    eng = Engine(d)
    eng.compute()
    f1 = eng.get_forces()
    assert np.abs(f1.sum(axis=0)).max() < 0.5           # sum of ~1e7 kJ/mol/nm of |F|: direct space cancels exactly (fixed point), PME to ~1e-7
    eng.compute()
    assert np.array_equal(f1, eng.get_forces()) or np.abs(f1 - eng.get_forces()).max() < 1e-3   # float atomics in the spread only


def test_checkpoint_roundtrip(mods):
    systems, Engine, *_ = mods
    d = systems.water_box(5, cutoff=0.75).rounded()
    eng = Engine(d)
    eng.set_integrator(systems.INT_LANGEVIN, 0.002, 300.0, 1.0, 5)
    eng.step(10)
    blob = eng.checkpoint()
    eng.step(10)
    xa = eng.get_positions()
    eng.load_checkpoint(blob)
    eng.step(10)
    # not bit-identical: the tile list is rebuilt on load, which re-partitions the fp32 partial sums
    assert np.abs(xa - eng.get_positions()).max() < 1e-4


_ASYNC_SCRIPT = r"""
import sys, os, numpy as np
sys.path.insert(0, os.environ["B200MD_ROOT"])
from openmm_b200 import systems, Engine
d = systems.water_box(12, cutoff=0.9).rounded()
eng = Engine(d)
eng.set_integrator(systems.INT_LANGEVIN, 0.002, 300.0, 1.0, 11, 1e-5)
eng.step(400)
st = eng.stats()
x = eng.get_positions()
e = eng.compute(); f = eng.get_forces()          # through the (other) current list, after a synchronous check
fresh = Engine(d); fresh.set_positions(x)
e2 = fresh.compute(); f2 = fresh.get_forces()    # a list built from scratch at the same positions
print("RESULT", st["list_builds"], st["stale_list_steps"], abs(e - e2), np.abs(f - f2).max(), np.isfinite(x).all())
"""


def test_successor_list_built_beside_the_step(mods):
    """B200MD_ASYNC_LIST=1: the next neighbour list is built on a side stream while the current one still serves the tile
    kernel, and k_integrate's last block flips them.  400 steps of a 5,184-atom box must rebuild many times, never serve a
    stale list, and end in a state whose forces equal those of an engine that builds its list from scratch."""
    import subprocess, sys
    from conftest import ROOT
    env = dict(os.environ, B200MD_ASYNC_LIST="1", B200MD_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-c", _ASYNC_SCRIPT], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    builds, stale, de, df, finite = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][0].split()[1:]
    assert int(builds) > 50 and int(stale) == 0 and finite == "True"
    assert float(de) < 0.05 and float(df) < 0.05      # fp32 summation order differs between two lists, nothing else


def test_molecules_that_walk_away_are_wrapped_but_the_trajectory_stays_continuous(mods):
    """fp32 coordinates must stay small, the user's trajectory must stay continuous (the Reference platform never wraps:
    ReferenceUpdateStateDataKernel::getPositions).  A 648-atom water box is given a 40 nm/ps drift: in 300 steps of 1 fs
    every molecule crosses the 1.86 nm box six times.  Newtonian dynamics is Galilean invariant, so the run must equal
    the run without drift shifted by v*t; the internal coordinates (checkpoint blob) must stay within two box lengths."""
    systems, Engine, *_ = mods
    d = systems.water_box(6, cutoff=0.9).rounded()
    L = float(d.box[0][0])
    rng = np.random.default_rng(3)
    v0 = rng.normal(0.0, 0.4, size=(d.natoms, 3))
    drift = np.array([40.0, -25.0, 0.0])
    nsteps, dt = 300, 0.001
    out = []
    for dv in (np.zeros(3), drift):
        eng = Engine(d)
        eng.set_integrator(systems.INT_VERLET, dt, 0.0, 0.0, 0, 1e-6)
        eng.set_velocities(v0 + dv)
        eng.apply_velocity_constraints(1e-6)
        eng.step(nsteps)
        out.append((eng.get_positions(), eng.checkpoint(), eng.stats()))
    (xa, _, _), (xb, blob, st) = out
    # same trajectory in the co-moving frame: measured 1e-5 typical, 2.5e-3 max (fp32 rounding differs between the frames
    # and 0.3 ps of water dynamics amplifies it); a missed or doubled lattice vector would be 1.86 nm
    assert np.abs((xb - drift*dt*nsteps) - xa).max() < 1e-2
    assert np.abs(xb - xa).max() > 5*L                                # ... although everything left the box many times
    npad = st["padded_atoms"]
    hdr = len(blob) - 2*16*npad - 3*4*npad
    inner = np.frombuffer(blob, dtype=np.float32, count=4*npad, offset=hdr).reshape(npad, 4)[:d.natoms, :3]
    lo = d.positions.min(axis=0)
    assert (inner > lo - 1.1*L - 0.5).all() and (inner < lo + 2.1*L + 0.5).all()      # wrapped back whenever > 1 box away
    eng.load_checkpoint(blob)                                         # the lattice offsets travel with the checkpoint
    assert np.abs(eng.get_positions() - xb).max() < 1e-6


def test_box_too_small_is_an_error(mods):
    systems, Engine, engine, *_ = mods
    d = systems.water_box(5, cutoff=0.9)       # box 1.55 nm < 2*0.9
    with pytest.raises(engine.EngineError):
        Engine(d)


def _all_bonds(systems, d):
    """constraints=AllBonds on a real System: every harmonic bond that is not yet constrained becomes a constraint at its
    equilibrium length (what forcefield.py does, wrappers/python/openmm/app/forcefield.py createSystem): the protein becomes
    ONE general constraint network, i.e. CCMA (ReferenceConstraints.cpp:148-184)."""
    have = set((min(i, j), max(i, j)) for i, j in zip(d.con_i, d.con_j))
    ci, cj, cd = list(d.con_i), list(d.con_j), list(d.con_d)
    for i, j, r0 in zip(d.bond_i, d.bond_j, d.bond_r0):
        if (min(i, j), max(i, j)) not in have:
            ci.append(int(i)); cj.append(int(j)); cd.append(float(r0))
    d.con_i, d.con_j, d.con_d = np.array(ci, dtype=np.int32), np.array(cj, dtype=np.int32), np.array(cd)
    return d


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_ccma_general_constraint_network_follows_reference(mods, kind):
    """DHFR with constraints=AllBonds (2,500 coupled protein constraints + SETTLE waters): deterministic integrators must
    follow the Reference platform's CCMA (ReferenceCCMAAlgorithm.cpp:235-316) and keep every constraint to the tolerance."""
    systems, Engine, _, omm, _ = mods
    d = _all_bonds(systems, systems.SystemDesc.load(os.path.join(os.path.dirname(GOLDEN), "..", "data", "dhfr.npz")).rounded())
    v = np.random.default_rng(3).standard_normal((d.natoms, 3))*0.2
    eng = Engine(d)
    eng.set_integrator(kind, 0.001, 0.0, 1.0, 7, 1e-6)
    sim = omm.Simulation(d, "Reference", integrator=(kind, 0.0, 1.0, 0.001), constraint_tol=1e-6, pme=d.pme_parameters())
    eng.apply_constraints(1e-6); sim.apply_constraints(1e-6)
    x0 = sim.state(positions=True)["positions"]
    assert np.abs(eng.get_positions() - x0).max() < 2e-6
    eng.set_positions(x0)
    eng.set_velocities(v); sim.set_velocities(v)
    eng.step(10); sim.step(10)
    ref = sim.state(positions=True, velocities=True)
    x = eng.get_positions()
    # fp32 positions at 4-6 nm resolve 5e-7 nm; every CCMA iteration of every step rounds once more (the free-atom / SETTLE
    # trajectory test above holds 5e-6): 5e-5 nm after 10 steps, 2.1e-5 measured
    assert np.abs(x - ref["positions"]).max() < 5e-5
    dist = np.linalg.norm(x[d.con_i] - x[d.con_j], axis=1)
    assert np.abs(dist/d.con_d - 1).max() < 2e-5
    eng.step(300)
    x = eng.get_positions()
    dist = np.linalg.norm(x[d.con_i] - x[d.con_j], axis=1)
    assert np.isfinite(x).all() and np.abs(dist/d.con_d - 1).max() < 2e-5


def test_ccma_small_chain_and_methane(mods):
    """tests/TestVerletIntegrator.h:230-275 (testConstrainedChain) shape, plus a CH4-like centre with four partners (more than
    the three an X-H_n SHAKE cluster takes): both are general networks."""
    systems, Engine, *_ = mods
    d = systems.lj_fluid(4, cutoff=0.7)
    d.con_i = np.array([0, 1, 2, 10, 10, 10, 10], dtype=np.int32); d.con_j = np.array([1, 2, 3, 11, 12, 13, 14], dtype=np.int32)
    x = d.positions
    d.con_d = np.linalg.norm(x[d.con_i] - x[d.con_j], axis=1)
    eng = Engine(d.rounded())
    eng.set_velocities(np.random.default_rng(1).standard_normal((d.natoms, 3))*0.5)
    eng.set_integrator(systems.INT_VERLET, 0.002, 0, 0, 0, 1e-6)
    eng.step(500)
    xn = eng.get_positions()
    assert np.abs(np.linalg.norm(xn[d.con_i] - xn[d.con_j], axis=1)/d.con_d - 1).max() < 1e-5


def test_nve_energy_drift_verlet(mods):
    """Energy conservation of the fp32 state (no posqCorrection): Verlet 1 fs, rigid water, 10,000 steps; the drift of the
    total energy per degree of freedom must stay far below kT (2.49 kJ/mol at 300 K)."""
    systems, Engine, *_ = mods
    d = systems.water_box(8, cutoff=0.9).rounded()
    eng = Engine(d)
    eng.set_integrator(systems.INT_LANGEVIN, 0.002, 300.0, 5.0, 3)
    eng.step(1000)                                  # thermalise off the lattice
    eng.set_integrator(systems.INT_VERLET, 0.001, 0, 0, 0, 1e-6)
    dof = 3*d.natoms - len(d.con_i) - 3

    def total():
        e = eng.compute()
        return e + eng.kinetic_energy()
    e0 = total()
    es = []
    for _ in range(10):
        eng.step(1000)
        es.append(total())
    drift = (np.array(es) - e0)/dof
    assert np.abs(drift).max() < 0.01, drift        # kJ/mol per degree of freedom over 10 ps
