"""Diagnose a parity failure: per-term comparison against the Reference platform, worst atoms."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openmm_b200 import systems, Engine
from openmm_b200.engine import TERM_NB_DIRECT, TERM_NB_RECIP, TERM_BONDS, TERM_ANGLES, TERM_TORSIONS, TERM_ALL
from oracle import omm

name = sys.argv[1] if len(sys.argv) > 1 else "apoa1"
d = systems.SystemDesc.load(os.path.join("data", name + ".npz")).rounded()
print(name, d.natoms, "pos range", d.positions.min(0), d.positions.max(0), "box", np.diag(d.box))
eng = Engine(d)
pme = d.pme_parameters()
sim = omm.Simulation(d, "Reference", pme=pme, recip_group=1)
for label, terms, groups in (("recip", TERM_NB_RECIP, 2), ("all-but-recip", TERM_ALL & ~TERM_NB_RECIP, 1), ("all", TERM_ALL, 3)):
    e = eng.compute(terms)
    f = eng.get_forces()
    fr, er = sim.forces_energy(groups)
    err = np.abs(f - fr).max(axis=1)
    rel = err/np.maximum(1, np.linalg.norm(fr, axis=1))
    w = np.argsort(-rel)[:8]
    print("%-14s E %.4f ref %.4f | max rel %.3e abs %.3e | n(rel>1e-4)=%d" % (label, e, er, rel.max(), err.max(), (rel > 1e-4).sum()))
    for a in w[:5]:
        print("     atom %6d q=%.3f sig=%.3f eps=%.3f nexc=%d  f=%s fref=%s" % (a, d.charges[a], d.sigmas[a], d.epsilons[a],
              int((d.exc_i == a).sum() + (d.exc_j == a).sum()), np.round(f[a], 3), np.round(fr[a], 3)))
print(eng.stats())
# direct-only pieces: engine direct vs C oracle pieces is too slow at 92k; check bonded separately against the C port
from oracle import port
fb = np.zeros((d.natoms, 3))
import ctypes
L = port.lib()
eb = L.orc_bonds(len(d.bond_i), port._ip(port._i(d.bond_i)), port._ip(port._i(d.bond_j)), port._dp(port._d(d.bond_r0)), port._dp(port._d(d.bond_k)), port._dp(port._d(d.positions)), port._dp(fb))
e = eng.compute(TERM_BONDS); f = eng.get_forces()
print("bonds   E %.4f port %.4f maxabs %.3e" % (e, eb, np.abs(f-fb).max()))
fb[:] = 0
ea = L.orc_angles(len(d.angle_i), port._ip(port._i(d.angle_i)), port._ip(port._i(d.angle_j)), port._ip(port._i(d.angle_k)), port._dp(port._d(d.angle_t0)), port._dp(port._d(d.angle_kk)), port._dp(port._d(d.positions)), port._dp(fb))
e = eng.compute(TERM_ANGLES); f = eng.get_forces()
print("angles  E %.4f port %.4f maxabs %.3e" % (e, ea, np.abs(f-fb).max()))
fb[:] = 0
et = L.orc_torsions(len(d.tor_i), port._ip(port._i(d.tor_i)), port._ip(port._i(d.tor_j)), port._ip(port._i(d.tor_k)), port._ip(port._i(d.tor_l)), port._ip(port._i(d.tor_n)), port._dp(port._d(d.tor_phase)), port._dp(port._d(d.tor_kk)), port._dp(port._d(d.positions)), port._dp(fb))
e = eng.compute(TERM_TORSIONS); f = eng.get_forces()
w = np.argmax(np.abs(f-fb).max(axis=1))
print("torsion E %.4f port %.4f maxabs %.3e at atom %d" % (e, et, np.abs(f-fb).max(), w))
