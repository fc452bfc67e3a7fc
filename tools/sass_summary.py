"""Per-kernel SASS summary of the shipped librt_b200.so (no GPU needed): instruction count, the instruction mix that
matters on this path (global / shared / LOCAL memory, FP32, FP64, votes and shuffles, branches, calls), and the
proof that nothing here is tensor-core or TMA code - a non-contraction path.
usage: python tools/sass_summary.py [out.txt]      (writes profiles/r2_sass_summary.txt by default)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "raytracer.js_b200", "librt_b200.so")
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_summary.txt")
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (source_fingerprint)

GROUPS = [
    ("LDG global loads", r"^LDG"), ("  of them 128-bit", r"^LDG\.E\.128|^LDG\.E\.[A-Z.]*128"), ("STG global stores", r"^STG"),
    ("LDS/STS shared", r"^(LDS|STS)"), ("LDL local loads (spills, stacks)", r"^LDL"), ("STL local stores", r"^STL"),
    ("LDC/LDCU constant (kernel params)", r"^LDCU?"), ("FP32 FFMA/FMUL/FADD/FMNMX/FSETP", r"^(FFMA|FMUL|FADD|FMNMX|FSETP|FSEL)"),
    ("FP64 DFMA/DMUL/DADD/DSETP", r"^(DFMA|DMUL|DADD|DSETP)"), ("MUFU (rcp, rsq, sqrt)", r"^MUFU"),
    ("VOTE / MATCH / REDUX", r"^(VOTE|MATCH|REDUX)"), ("SHFL", r"^SHFL"), ("ATOM / RED", r"^(ATOM|RED|ATOMG)"),
    ("BRA / BSSY / BSYNC", r"^(BRA|BSSY|BSYNC)"), ("CALL (out-of-line float64 div/sqrt/atan2, cold blocks)", r"^CALL"),
    ("tensor core (HMMA/IMMA/DMMA/UTCMMA/tcgen05)", r"^(HMMA|IMMA|DMMA|QMMA|UTC|TCGEN)"), ("TMA (UBLKCP/UTMA)", r"^(UBLK|UTMA)"),
]


def kernels():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, res = None, collections.OrderedDict()
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            res[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            res[cur].append(m.group(1))
    return res


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    except Exception:
        return n


def resources():
    txt = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    res, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            res[cur] = line.strip()
            cur = None
    return res


def main():
    ks, rs = kernels(), resources()
    lines = [f"# SASS summary of raytracer.js_b200/librt_b200.so (cuobjdump -sass, sm_100a), source fingerprint {bench.source_fingerprint()}",
             "# static instruction counts per kernel; `rt_*` are this repo's kernels, the rest is CUB (device tree builder)", ""]
    for name, ins in ks.items():
        d = demangle(name)
        if "rt_" not in d or "cub::" in d:
            continue
        lines.append(f"{d}")
        lines.append(f"    {len(ins)} instructions; {rs.get(name, '')}")
        for label, pat in GROUPS:
            c = sum(1 for i in ins if re.match(pat, i))
            if c or "tensor" in label or "TMA" in label or "local" in label.lower():
                lines.append(f"    {c:6d}  {label}")
        lines.append("")
    open(out_path, "w").write("\n".join(lines))
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
