"""Aggregate an ncu source page by code region (function) of rt_trace.cuh / rt_b200.cu: share of warp
instructions, active threads per instruction, share of stall samples.
usage: python tools/ncu_regions.py report.ncu-rep"""
import csv, re, subprocess, sys, os
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; L = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            k = (cur, int(r[0])); v = L.setdefault(k, [0, 0, 0])
            v[0] += int(d["Instructions Executed"]); v[1] += int(d["Thread Instructions Executed"]); v[2] += int(d["# Samples"])
        except Exception: pass
tot = sum(v[0] for v in L.values()); ts = sum(v[2] for v in L.values())
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "raytracer.js_b200", "csrc")
marks = {}
fn = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:RT_HD|RT_COLD|RT_D|__global__)[^;{]*?\b([A-Za-z_0-9]+)\s*\(")
for f in ("rt_trace.cuh", "rt_b200.cu"):
    src = open(os.path.join(root, f)).read().split("\n")
    ms = []
    for i, l in enumerate(src):
        m = fn.match(l.strip()) if not l.startswith((" ", "\t")) or "rt_" in l else None
        if l.startswith(("RT_HD", "RT_COLD", "RT_D", "__global__", "    rt_", "\trt_")) or (l.startswith("template") and i + 1 < len(src)):
            t = l if not l.startswith("template") else src[i + 1]
            m = re.search(r"\b([A-Za-z_0-9]+)\s*\(", t.replace("__launch_bounds__", "").replace("RT_WARPS_PER_CTA", ""))
            if m: ms.append((i + 1, m.group(1)))
    marks[f] = ms
agg = {}
for (f, ln), (wi, ti, sm) in L.items():
    k = f
    if f in marks:
        c = [n for m, n in marks[f] if m <= ln]
        k = f.split(".")[0][3:] + ":" + (c[-1] if c else "header")
    a = agg.setdefault(k, [0, 0, 0]); a[0] += wi; a[1] += ti; a[2] += sm
print(f"total warp-instr {tot:,} samples {ts:,}")
for k, (wi, ti, sm) in sorted(agg.items(), key=lambda x: -x[1][0])[:30]:
    print(f"{k:40s} inst {100*wi/tot:5.1f}%  thr/inst {ti/max(wi,1):5.1f}  stall {100*sm/ts:5.1f}%")
