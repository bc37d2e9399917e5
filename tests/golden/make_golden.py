"""Generates the committed golden fixtures of tests/golden/ from the reference IN THIS CONTAINER (/root/reference is
not present on the GPU box).  Run: python tests/golden/make_golden.py

1. ewald_triclinic_gromacs.json -- the known-answer vector the reference's own test holds: 8 ions in a triclinic box,
   PME grid 32x40x48, alpha 3.45891, Gromacs forces and energy (tests/TestEwald.h:222-271; tolerance there 1e-4).
   Parsed from the header text, not retyped.
2. nacl_amorph.npz -- the 894-ion amorphous NaCl positions of tests/nacl_amorph.dat (used by tests/TestEwald.h:98-220)
   plus forces/energy of the reference's Reference platform (oracle/_ref/libOpenMM.so) with PME, cutoff 1.2,
   tol 1e-5 pinned to an FFT-friendly grid, and the Gromacs energy -3.82047e5 quoted by the test (Ewald, :150).
3. water5_reference.npz -- 375-atom TIP3P box: Reference-platform forces/energy (PME) for the seeded S1 recipe.
"""
import json
import os
import re
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def parse_triclinic():
    src = open(os.path.join(REF, "tests/TestEwald.h")).read()
    blk = src[src.index("void testTriclinic()"):src.index("void testTriclinic2()")]
    vec = r"Vec3\(([^,]+),([^,]+),([^)]+)\)"
    box = [[float(x) for x in m] for m in re.findall(vec, re.search(r"setDefaultPeriodicBoxVectors\((.*)\);", blk).group(1))]
    pos = [[float(x) for x in m.groups()[1:]] for m in re.finditer(r"positions\[(\d)\] = " + vec, blk)]
    frc = [[float(x) for x in m.groups()[1:]] for m in re.finditer(r"expectedForce\[(\d)\] = " + vec, blk)]
    alpha, nx, ny, nz = re.search(r"setPMEParameters\(([^,]+),([^,]+),([^,]+),([^)]+)\)", blk).groups()
    energy = float(re.search(r"expectedEnergy = ([-0-9.e+]+)", blk).group(1))
    parts = re.findall(r"addParticle\(([-0-9.]+), ([0-9.]+), ([0-9.]+)\)", blk)
    assert len(pos) == 8 and len(frc) == 8 and len(parts) == 2
    out = {"source": "tests/TestEwald.h:222-271 (Gromacs)", "box": box, "positions": pos, "expected_forces": frc,
           "expected_energy": energy, "alpha": float(alpha), "grid": [int(nx), int(ny), int(nz)], "cutoff": 1.0,
           "charges": [float(parts[0][0])]*4 + [float(parts[1][0])]*4,
           "sigmas": [float(parts[0][1])]*4 + [float(parts[1][1])]*4,
           "epsilons": [float(parts[0][2])]*4 + [float(parts[1][2])]*4, "tolerance": 1e-4}
    json.dump(out, open(os.path.join(HERE, "ewald_triclinic_gromacs.json"), "w"), indent=1)
    return out


def nacl():
    from openmm_b200 import systems
    from oracle import omm
    txt = open(os.path.join(REF, "tests/nacl_amorph.dat")).read()
    pos = np.array([[float(x) for x in m] for m in re.findall(r"Vec3\(([^,]+),([^,]+),([^)]+)\)", txt)])
    assert pos.shape == (894, 3)
    # identical inputs for the fp32 device path and the double oracle: positions rounded to fp32-representable values
    # (rounding 3 nm coordinates moves them by <= 1.2e-7 nm, which alone changes ion-ion forces by ~2e-4 relative)
    pos = pos.astype(np.float32).astype(np.float64)
    n = 894
    L = 3.00646
    q = np.concatenate([np.ones(n//2), -np.ones(n//2)])
    d = systems.SystemDesc(masses=np.concatenate([np.full(n//2, 22.99), np.full(n//2, 35.45)]), charges=q, sigmas=np.ones(n), epsilons=np.zeros(n),
                           positions=pos, box=np.diag([L, L, L]), method=systems.NB_PME, cutoff=1.2, ewald_tol=1e-5, name="nacl_amorph")
    pme = d.pme_parameters()
    sim = omm.Simulation(d, "Reference", pme=pme)
    f, e = sim.forces_energy()
    np.savez_compressed(os.path.join(HERE, "nacl_amorph.npz"), positions=pos, box=L, charges=q, cutoff=1.2, ewald_tol=1e-5,
                        pme=np.array(pme), reference_forces=f, reference_energy=e, gromacs_energy=-3.82047e5)
    print("nacl_amorph: Reference PME energy %.3f (Gromacs Ewald %.3f), grid %s" % (e, -3.82047e5, pme))


def water():
    from openmm_b200 import systems
    from oracle import omm
    d = systems.water_box(5, cutoff=0.75).rounded()
    pme = d.pme_parameters()
    sim = omm.Simulation(d, "Reference", pme=pme)
    f, e = sim.forces_energy()
    np.savez_compressed(os.path.join(HERE, "water5_reference.npz"), positions=d.positions, pme=np.array(pme), reference_forces=f, reference_energy=e)
    print("water5: Reference energy %.4f" % e)


if __name__ == "__main__":
    t = parse_triclinic()
    print("triclinic golden: E=%g, grid %s" % (t["expected_energy"], t["grid"]))
    nacl()
    water()
