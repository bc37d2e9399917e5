"""gpurun_out/r02_launches_<w>.csv + gpurun_out/r02_kernels_<w>.ncu-rep  ->  profiles/r02_<w>_launch_summary.csv,
profiles/r02_<w>_kernels.md (one row per kernel: time, DRAM bytes, issue-active, pipes, registers, occupancy) and
profiles/r02_<w>_k_pair_summary.json (read by bench.py for roofline.traffic).      python tools/summarize_profile_r02.py dhfr"""
import collections
import csv
import io
import json
import os
import statistics
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "profiles")
w = sys.argv[1] if len(sys.argv) > 1 else "dhfr"


def short(name):
    return name.split("(")[0].replace("void ", "").split("<")[0]


lpath = os.path.join(root, "gpurun_out", "r02_launches_%s.csv" % w)
if os.path.exists(lpath):
    rows = [r for r in csv.reader(l for l in open(lpath) if l.startswith('"'))]
    h = rows[0]
    ik, iv, im = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
    per = collections.defaultdict(list)
    for r in rows[1:]:
        if len(r) > iv and r[im] == "gpu__time_duration.sum":
            per[short(r[ik])].append(float(r[iv].replace(",", ""))/1e3)
    total = sum(sum(v) for v in per.values())
    with open(os.path.join(out, "r02_%s_launch_summary.csv" % w), "w") as f:
        f.write("# per-kernel device time over %d consecutive launches (40 MD steps) of tools/gpu_steps.py %s: ncu --metrics gpu__time_duration.sum --clock-control none; "
                "CUDA graphs and the stream fork off, so every launch is visible and serialised (cold-ish caches: compare SHARES)\n" % (sum(len(v) for v in per.values()), w))
        f.write("kernel,launches,total_us,median_us,max_us,share_pct\n")
        for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            f.write("%s,%d,%.1f,%.2f,%.1f,%.1f\n" % (k, len(v), sum(v), statistics.median(v), max(v), 100*sum(v)/total))
    print("launch summary:", len(per), "kernels, total %.0f us" % total)

rep = os.path.join(root, "gpurun_out", "r02_kernels_%s.ncu-rep" % w)
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    h, units = r[0], r[1]
    col = {k: i for i, k in enumerate(h)}

    def val(row, k):
        if k not in col or row[col[k]] in ("", "n/a"):
            return None
        x = float(row[col[k]].replace(",", ""))
        u = units[col[k]]
        return x*{"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "usecond": 1e3, "msecond": 1e6, "second": 1e9, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1.0)
    best = {}
    for row in r[2:]:
        name = short(row[col["Kernel Name"]])
        t = val(row, "gpu__time_duration.sum")
        if t is None:
            continue
        if name not in best or t > best[name][0]:            # keep the longest instance (a rebuild step for the list kernels)
            best[name] = (t, row)
    cols = [("time_us", "gpu__time_duration.sum", 1e-3), ("dram_rd_MB", "dram__bytes_read.sum", 1e-6), ("dram_wr_MB", "dram__bytes_write.sum", 1e-6),
            ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1), ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1),
            ("issue_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1), ("fma_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1),
            ("xu_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1), ("fp64_pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 1),
            ("lsu_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1), ("warp_inst_M", "smsp__inst_executed.sum", 1e-6),
            ("regs", "launch__registers_per_thread", 1), ("occ_pct", "sm__warps_active.avg.pct_of_peak_sustained_active", 1), ("grid", "launch__grid_size", 1), ("block", "launch__block_size", 1)]
    with open(os.path.join(out, "r02_%s_kernels.md" % w), "w") as f:
        f.write("# ncu --set full, one launch per kernel (the longest of ~5 steady-state steps), workload %s, B200, --clock-control none\n\n" % w)
        f.write("| kernel | " + " | ".join(c[0] for c in cols) + " |\n|---|" + "---|"*len(cols) + "\n")
        for name, (t, row) in sorted(best.items(), key=lambda kv: -kv[1][0]):
            cells = []
            for _, k, sc in cols:
                x = val(row, k)
                cells.append("-" if x is None else ("%.2f" % (x*sc) if sc != 1 or abs(x) < 1000 else "%d" % x))
            f.write("| %s | %s |\n" % (name, " | ".join(cells)))
    print("kernel table:", len(best), "kernels")
    if "k_pair" in best:
        t, row = best["k_pair"]
        summ = {"kernel": row[col["Kernel Name"]][:80], "workload": w, "source": "profiles/r02_%s_kernels.md (ncu --set full, 1 launch)" % w, "gpu_time_us": t*1e-3,
                "dram_bytes_read": int(val(row, "dram__bytes_read.sum")), "dram_bytes_write": int(val(row, "dram__bytes_write.sum")),
                "inst_executed_warp": int(val(row, "smsp__inst_executed.sum")), "issue_active_pct": val(row, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "registers": int(val(row, "launch__registers_per_thread"))}
        json.dump(summ, open(os.path.join(out, "r02_%s_k_pair_summary.json" % w), "w"), indent=1)
        print(json.dumps(summ))
