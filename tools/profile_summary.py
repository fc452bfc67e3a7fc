"""Summarise one round's profiling pass (tools/profile_round.sh TAG, files in gpurun_out/) into profiles/:
  profiles/TAG_launches.csv, TAG_c2_launches.csv   ncu launch lists (our kernels only)
  profiles/TAG_summary.json                        per kernel: duration, DRAM bytes, issue/occupancy/SIMT metrics
  profiles/TAG_regions_<kernel>.txt                source-page aggregation per function
  profiles/TAG_bench.json, TAG_c2_*.json           the plain (un-profiled) runs of the same commands
usage: python tools/profile_summary.py TAG"""
import csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
           "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
           "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors.sum", "lts__t_sectors.sum.per_second",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def to_bytes(v, unit):
    k = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit)
    return float(v) * k if k else float(v)


summary = {}
for name in ("primary", "c2_primary", "c2_bounce", "c2_resample"):
    rep = os.path.join(G, f"prof_{tag}_{name}.ncu-rep")
    if not os.path.exists(rep):
        continue
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units, vals = rows[0], rows[1], rows[2]
    d = {"kernel": vals[h.index("Kernel Name")]}
    for m in METRICS:
        if m in h:
            i = h.index(m)
            v = vals[i].replace(",", "")
            try:
                d[m] = to_bytes(v, units[i]) if "bytes" in m else float(v)
                if not "bytes" in m:
                    d[m + "#unit"] = units[i]
            except ValueError:
                d[m] = v
    summary[name] = d
    reg = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_regions.py"), rep], capture_output=True, text=True).stdout
    open(os.path.join(P, f"{tag}_regions_{name}.txt"), "w").write(f"# {d['kernel']}\n# ncu --set full --import-source on, aggregated per function by tools/ncu_regions.py\n" + reg)
# bench.py quotes these numbers only while the kernel sources are the ones they were captured from
# (the fingerprint the profiled run itself printed: the working tree may have moved on since the GPU call)
fp = None
try:
    lines = [l for l in open(os.path.join(G, f"{tag}_bench.log")).read().splitlines() if l.startswith("{")]
    fp = json.loads(lines[-1])["roofline"]["source_fingerprint"]
except Exception:
    pass
summary["fingerprint"] = fp
summary["fingerprint_of"] = "sha256[:16] over raytracer.js_b200/csrc/* + include/rt_b200.h (bench.source_fingerprint), as printed by the profiled run"
json.dump(summary, open(os.path.join(P, f"{tag}_summary.json"), "w"), indent=1)

for f in (f"{tag}_launches.csv", f"{tag}_c2_launches.csv"):
    src = os.path.join(G, f)
    if os.path.exists(src):
        rows = [r for r in csv.reader(open(src)) if r and (r[0] == "ID" or (len(r) > 4 and "rt_" in r[4]))]
        csv.writer(open(os.path.join(P, f), "w")).writerows(rows)
for f, dst in ((f"{tag}_bench.log", f"{tag}_bench.json"), (f"{tag}_c2_1spp.log", f"{tag}_c2_1spp.json"), (f"{tag}_c2_16spp.log", f"{tag}_c2_16spp.json")):
    src = os.path.join(G, f)
    if os.path.exists(src):
        lines = [l for l in open(src).read().splitlines() if l.startswith("{")]
        if lines:
            open(os.path.join(P, dst), "w").write(lines[-1] + "\n")
# the dominant kernel's DRAM traffic per launch, read by bench.py
if "primary" in summary:
    s = summary["primary"]
    json.dump({"kernel": s["kernel"], "dram_bytes_per_launch": s["dram__bytes_read.sum"] + s["dram__bytes_write.sum"],
               "read": s["dram__bytes_read.sum"], "write": s["dram__bytes_write.sum"],
               "source": f"profiles/{tag}_summary.json (ncu --set full, bench.py --steps 3 --warmup 3), L2 not flushed under ncu"},
              open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(json.dumps({k: {m: v for m, v in d.items() if m in ("gpu__time_duration.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
                                                             "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum")} for k, d in summary.items() if isinstance(d, dict)}, indent=1))
