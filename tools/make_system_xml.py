"""Write the benchmark Systems as the reference's own XML (XmlSerializer::serialize<System>) plus a PDB file with the
coordinates, for C++ host applications (plugin/examples/run_system_xml.cpp):

    python tools/make_system_xml.py dhfr /tmp/dhfr        ->  /tmp/dhfr.xml, /tmp/dhfr.pdb

The System is rebuilt from data/<name>.npz (produced by the reference's forcefield.py, tools/make_benchmark_systems.py)
through the reference library itself (oracle/_ref/libOpenMM.so via the ctypes shim), so the XML is exactly what a user of
the reference would have serialized."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openmm_b200 import systems       # noqa: E402
from oracle import omm                # noqa: E402


def write_pdb(path, d):
    with open(path, "w") as f:
        if d.box is not None:
            f.write("CRYST1%9.3f%9.3f%9.3f  90.00  90.00  90.00 P 1           1\n" % (10*d.box[0][0], 10*d.box[1][1], 10*d.box[2][2]))
        for i, p in enumerate(d.positions):
            f.write("ATOM  %5d  X   UNK A%4d    %8.3f%8.3f%8.3f  1.00  0.00\n" % ((i + 1) % 100000, (i//3 + 1) % 10000, 10*p[0], 10*p[1], 10*p[2]))
        f.write("END\n")


def main():
    name, out = sys.argv[1], sys.argv[2]
    d = systems.SystemDesc.load(os.path.join(ROOT, "data", name + ".npz"))
    sim = omm.Simulation(d, "Reference", integrator=(systems.INT_LANGEVIN, 300.0, 1.0, 0.002), pme=d.pme_parameters())
    if omm.lib().omm_system_serialize(sim.sys, (out + ".xml").encode()) != 0:
        raise SystemExit("serialization failed")
    write_pdb(out + ".pdb", d)
    print("wrote %s.xml (%d bytes), %s.pdb (%d atoms)" % (out, os.path.getsize(out + ".xml"), out, d.natoms))


if __name__ == "__main__":
    main()
