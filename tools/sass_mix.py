#!/usr/bin/env python
"""SASS instruction mix of the product library's kernels (cuobjdump -sass instancefusion_b200/libef_track.so), per kernel:
opcode histogram of the static code, plus the Blackwell-specific mnemonics B200_PROFILING.md asks about.  No GPU needed."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "instancefusion_b200", "libef_track.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern = None
per = collections.OrderedDict()
for line in txt.split("\n"):
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        per[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)(?:\.[A-Z0-9_.]+)?\s", line)
    if m and kern:
        per[kern][m.group(1)] += 1


def short(name):
    out = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    out = re.sub(r"\(anonymous namespace\)::", "", out)
    return out[:100]


total = collections.Counter()
print(f"cuobjdump -sass {os.path.relpath(so, ROOT)}: {len(per)} kernels, static instruction counts\n")
for k, c in sorted(per.items(), key=lambda kv: -sum(kv[1].values())):
    n = sum(c.values())
    total.update(c)
    print(f"{short(k)}\n    {n} instructions: " + ", ".join(f"{op} {v}" for op, v in c.most_common(14)))
print("\nwhole library: " + ", ".join(f"{op} {v}" for op, v in total.most_common(30)))
probe = ["FFMA2", "FMUL2", "FADD2", "UTMALDG", "UTMASTG", "UTCMMA", "UTCHMMA", "UTCQMMA", "TCGEN05", "LDGSTS", "UBLKCP", "SYNCS", "DFMA", "DMUL", "REDUX", "SHFL",
         "LDG", "STG", "LDS", "STS", "BAR", "MEMBAR", "ATOMG", "RED"]
print("\nmnemonics of interest: " + ", ".join(f"{p} {total.get(p, 0)}" for p in probe))
print("""
Reading: no tensor-core (UTC*MMA / tcgen05) and no TMA (UTMALDG / UTMASTG) instructions -- by design: nothing on this path is a dense
contraction (28 FMAs per 48 gathered bytes), the stencils are 5x5 over tiles that fit shared memory, and the step kernels are
gather-bound.  No packed FFMA2 either: measured to have exactly the FMA throughput of two FFMAs on sm_100 and to free no issue
slots (tools/ffma2_tput.cu, profiles/r02_k_track_experiments.txt).  The 128-bit flagged-chunk traffic of the persistent kernel is
LDG.E.128.STRONG.GPU / STG.E.128.STRONG.GPU from ld/st.relaxed.gpu.global.b128.""")
