"""Static SASS summary of librl4afcs_b200.so: per kernel, instruction counts by class (cuobjdump -sass).

    python scripts/sass_summary.py > profiles/sass_summary_r02.txt

What to look for: the fused kernels are FP64 / FP32 vector-pipe code (DFMA / DMUL / DADD, FFMA / FMUL / FADD) with MUFU seeds,
shuffles (SHFL) and named barriers (BAR) in the pipeline kernel; there is no tensor-core (UTC*MMA / HMMA) or TMA
(UTMALDG / UBLKCP) instruction because no contraction on this path is larger than (1,10)@(10,3) and every chain's order is
fixed by the parity contract."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "rl4afcs_b200", "librl4afcs_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
CLASSES = [("fp64", r"^(DFMA|DMUL|DADD|DSETP|DMNMX)"), ("fp32", r"^(FFMA|FMUL|FADD|FSETP|FSEL|FMNMX|FCHK)"), ("mufu", r"^(MUFU)"),
           ("cvt", r"^(F2F|I2F|F2I|F2FP)"), ("int/mov", r"^(IMAD|IADD3|LOP3|MOV|SHF|LEA|ISETP|SEL|VIADD|UMOV|CS2R|PLOP3|PRMT|R2UR|S2R|S2UR|ULEA|UIADD3|VIMNMX|VIADDMNMX|IABS|UISETP|ULOP3|USEL|UIMAD)"),
           ("global", r"^(LDG|STG)"), ("local", r"^(LDL|STL)"), ("shared", r"^(LDS|STS)"), ("const", r"^(LDC|LDCU)"),
           ("shuffle", r"^(SHFL|VOTE|VOTEU|MATCH|WARPSYNC)"), ("barrier", r"^(BAR)"), ("branch/call", r"^(BRA|BSSY|BSYNC|CALL|RET|EXIT|NOP|BRX|JMP)"),
           ("tensor/TMA", r"^(UTC|HMMA|HGMMA|QGMMA|IGMMA|LDTM|STTM|UTMA|UBLKCP|LDGSTS)")]
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        for name, pat in CLASSES:
            if re.match(pat + ("" if name in ("tensor/TMA",) else r"$"), op):
                counts[cur][name] += 1
                break
        else:
            counts[cur]["other"] += 1
        counts[cur]["total"] += 1
names = [c[0] for c in CLASSES] + ["other", "total"]
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("static SASS instruction counts per kernel of rl4afcs_b200/librl4afcs_b200.so (sm_100a), by class")
print("  ".join(f"{n:>11}" for n in names) + "  kernel")
for (k, c), d in zip(counts.items(), demangle):
    # drop the parameter list (the LAST top-level parenthesis group; template arguments like `(int)1` stay)
    depth, cut = 0, len(d)
    for pos in range(len(d) - 1, -1, -1):
        if d[pos] == ")":
            depth += 1
        elif d[pos] == "(":
            depth -= 1
            if depth == 0:
                cut = pos
                break
    d = d[:cut]
    print("  ".join(f"{c[n]:>11}" for n in names) + "  " + d)
