"""Error attribution of the fp32 force path against the Reference platform (run on the GPU box):

    python tools/gpu_parity_probe.py apoa1 [dhfr] [nacl]

For every workload the Reference platform gives the direct-space group (+ bonded) and the reciprocal-space group
separately; every variant of the CUDA path (close-pair cutoff of the double-precision path, fp32 / double spectral
pipeline) is run in a fresh process and compared component by component:
  rel_own   = max_i |dF_i| / max(1, |F_i component|)         (the number a per-component test would see)
  rel_total = max_i |dF_i| / max(1, |F_i total|)             (the reference's ASSERT_EQUAL_VEC form on the full force)
  abs99/absmax = 99th percentile / maximum of |dF_i| in kJ/mol/nm
Results: gpurun_out/parity_probe.json."""
import json
import os
import subprocess
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")


def load(name):
    from openmm_b200 import systems
    if name == "nacl":
        z = np.load(os.path.join(ROOT, "tests", "golden", "nacl_amorph.npz"))
        n, L = 894, float(z["box"])
        pme = z["pme"]
        return systems.SystemDesc(masses=np.ones(n), charges=z["charges"], sigmas=np.ones(n), epsilons=np.zeros(n), positions=z["positions"],
                                  box=np.diag([L, L, L]), method=systems.NB_PME, cutoff=float(z["cutoff"]),
                                  pme_alpha=float(pme[0]), pme_grid=(int(pme[1]), int(pme[2]), int(pme[3]))).rounded()
    return systems.SystemDesc.load(os.path.join(ROOT, "data", name + ".npz")).rounded()


def child(name, tag):
    from openmm_b200 import Engine
    d = load(name)
    eng = Engine(d)
    out = {}
    for key, terms in (("direct", 1 | 2 | 4 | 8), ("recip", 16), ("total", 31)):
        e = eng.compute(terms)
        out["f_" + key] = eng.get_forces()
        out["e_" + key] = e
    out["pair_us"] = eng.time_phase("pair", 30)*1e3
    out["fft_us"] = eng.time_phase("pme_fft_conv", 30)*1e3
    np.savez(os.path.join(OUT, "probe_%s_%s.npz" % (name, tag)), **out)


def metrics(f, fr, ftot):
    d = np.abs(f - fr).max(axis=1)
    own = np.maximum(1.0, np.linalg.norm(fr, axis=1))
    tot = np.maximum(1.0, np.linalg.norm(ftot, axis=1))
    return {"rel_own": float((d/own).max()), "rel_total": float((d/tot).max()), "abs99": float(np.percentile(d, 99)), "absmax": float(d.max()),
            "n_over_1e-4_total": int((d/tot > 1e-4).sum())}


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        return child(sys.argv[2], sys.argv[3])
    from oracle import omm
    names = sys.argv[1:] or ["apoa1"]
    variants = [("close0", {"B200MD_CLOSE_NM": "0"}, "f32"), ("close0.36", {"B200MD_CLOSE_NM": "0.36"}, "f32")]
    report = {}
    for name in names:
        d = load(name)
        sim = omm.Simulation(d, "Reference", pme=d.pme_parameters(), recip_group=1)
        fdir, edir = sim.forces_energy(1)
        frec, erec = sim.forces_energy(2)
        ftot, etot = sim.forces_energy(3)
        sim.close()
        report[name] = {}
        for tag, env, lib in variants:
            e = dict(os.environ)
            e.update(env)
            if lib == "f64":
                e["B200MD_LIB"] = os.path.join(ROOT, "openmm_b200", "libb200md_dbl.so")
                if not os.path.exists(e["B200MD_LIB"]):
                    continue
            rc = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", name, tag], env=e).returncode
            if rc != 0:
                report[name][tag] = {"failed": rc}
                continue
            z = np.load(os.path.join(OUT, "probe_%s_%s.npz" % (name, tag)))
            r = {"direct": metrics(z["f_direct"], fdir, ftot), "recip": metrics(z["f_recip"], frec, ftot), "total": metrics(z["f_total"], ftot, ftot),
                 "e_rel": abs(float(z["e_total"]) - etot)/abs(etot), "pair_us": float(z["pair_us"]), "fft_us": float(z["fft_us"])}
            report[name][tag] = r
            print(name, tag, json.dumps(r), flush=True)
            os.remove(os.path.join(OUT, "probe_%s_%s.npz" % (name, tag)))
    json.dump(report, open(os.path.join(OUT, "parity_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
