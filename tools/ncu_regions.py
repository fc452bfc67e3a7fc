"""Aggregate an ncu source page by code region of rt_trace.cuh (regions = the functions / lambdas of the
packet walk).  usage: python tools/ncu_regions.py report.ncu-rep"""
import csv, subprocess, sys
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
src = open('raytracer.js_b200/csrc/rt_trace.cuh').read().split('\n')
pats = [('exact/candidate', 'exact float64 primitives'), ('per-ray walker', '// ------------------------------------------------------------------ the walk'),
        ('warp prims', 'the packet walk (camera rays)'), ('pierces_cube', 'RT_HD bool packet_pierces_cube'),
        ('meets_record', 'RT_HD bool packet_meets_record'), ('xor_permute', 'RT_HD unsigned xor_permute8'),
        ('packet setup', 'RT_HD void packet_primary_hits('), ('walk: head', 'int sp = 0;'), ('walk: push_children', 'auto push_children'),
        ('walk: scan', 'auto scan = '), ('walk: main loop', 'const int A = F.chain_node[k];'), ('shading helpers', 'shading helpers'),
        ('trace_path', 'Ray.trace (src/raytracer.ts:168-277)'), ('primary_terminal/store', 'primary stage: one 8x4 patch')]
marks = sorted((next(i + 1 for i, l in enumerate(src) if p in l), n) for n, p in pats)
agg = {}
for (f, ln), (wi, ti, sm) in L.items():
    k = f if f != 'rt_trace.cuh' else ([n for m, n in marks if m <= ln] or ['header'])[-1]
    a = agg.setdefault(k, [0, 0, 0]); a[0] += wi; a[1] += ti; a[2] += sm
print(f"total warp-instr {tot:,} samples {ts:,}")
for k, (wi, ti, sm) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print(f"{k:28s} inst {100*wi/tot:5.1f}%  thr/inst {ti/max(wi,1):5.1f}  stall {100*sm/ts:5.1f}%")
