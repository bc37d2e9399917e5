"""Turn the raw ncu outputs of a profiling run (gpurun_out/) into the tracked summaries under profiles/.
  launches.csv  : ncu --metrics gpu__time_duration.sum --clock-control none --csv  (one row per launch)
  <rep>.ncu-rep : ncu --set full capture of one k_pair launch"""
import csv, io, json, os, subprocess, sys, collections, statistics

tag = sys.argv[1] if len(sys.argv) > 1 else "r01_final"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "profiles")
lpath = os.path.join(root, "gpurun_out", "launches.csv")
if os.path.exists(lpath):
    rows = [r for r in csv.reader(l for l in open(lpath) if l.startswith('"'))]
    h = rows[0]
    ik, iv, im = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
    per = collections.defaultdict(list)
    for r in rows[1:]:
        if len(r) > iv and r[im] == "gpu__time_duration.sum":
            name = r[ik].split("(")[0].replace("void ", "")
            per[name].append(float(r[iv].replace(",", ""))/1e3)          # ns -> us
    total = sum(sum(v) for v in per.values())
    with open(os.path.join(out, tag + "_launch_summary.csv"), "w") as f:
        f.write("# %s: per-kernel device time over %d launches of `bench.py --md-steps 40` on DHFR (ncu --metrics gpu__time_duration.sum "
                "--clock-control none; CUDA graphs and stream overlap off so that every launch is visible; cold-cache, serialised: compare SHARES)\n" % (tag, sum(len(v) for v in per.values())))
        f.write("kernel,launches,total_us,median_us,max_us,share_pct\n")
        for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            f.write("%s,%d,%.1f,%.2f,%.1f,%.1f\n" % (k, len(v), sum(v), statistics.median(v), max(v), 100*sum(v)/total))
    print("launch summary written,", len(per), "kernels")
rep = os.path.join(root, "gpurun_out", "pair_final.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    h, v = r[0], r[2]
    g = lambda k: float(v[h.index(k)].replace(",", ""))
    with open(os.path.join(out, tag + "_k_pair_details.csv"), "w") as f:
        f.write("metric,unit,value\n")
        for i, k in enumerate(h):
            if "__" in k: f.write("%s,%s,%s\n" % (k, r[1][i], v[i]))
    def unit_scale(k):
        u = r[1][h.index(k)]
        return {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
    summ = {"kernel": v[h.index("Kernel Name")], "workload": "dhfr", "source": "profiles/%s_k_pair_details.csv (ncu --set full, 1 launch)" % tag,
            "gpu_time_us": g("gpu__time_duration.sum")/ (1e3 if r[1][h.index("gpu__time_duration.sum")] == "ns" else 1.0),
            "dram_bytes_read": int(g("dram__bytes_read.sum")*unit_scale("dram__bytes_read.sum")),
            "dram_bytes_write": int(g("dram__bytes_write.sum")*unit_scale("dram__bytes_write.sum")),
            "inst_executed_warp": int(g("smsp__inst_executed.sum")),
            "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "registers": int(g("launch__registers_per_thread"))}
    for key, met in (("pipe_fma_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), ("pipe_xu_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                     ("pipe_lsu_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active")):
        if met in h: summ[key] = g(met)
    json.dump(summ, open(os.path.join(out, tag + "_k_pair_summary.json"), "w"), indent=1)
    print(json.dumps(summ))
