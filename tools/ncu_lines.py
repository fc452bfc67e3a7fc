"""Summarise an ncu report per CUDA source line: instructions executed, avg active threads, stall samples.
usage: python tools_ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; hdr = None; lines = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            lines.append((cur_file, int(r[0]), r[1].strip()[:90], int(d["Instructions Executed"]), int(d["Thread Instructions Executed"]), int(d["# Samples"])))
        except Exception: pass
tot_i = sum(l[3] for l in lines); tot_s = sum(l[5] for l in lines)
print(f"total warp-instr {tot_i:,}  samples {tot_s:,}")
for f, ln, src, wi, ti, smp in sorted(lines, key=lambda l: -l[5])[:top]:
    print(f"{f}:{ln:<4d} inst {100*wi/tot_i:5.1f}%  thr/inst {ti/max(wi,1):5.1f}  stall {100*smp/tot_s:5.1f}%  | {src}")
