"""Tuning experiments: builds build_variants/librl4_<name>.so from ONE translation unit recompiled with extra -D flags
(the other units are the objects of the regular build in rl4afcs_b200/build/), and times a script with each of them.

    python scripts/variants.py build sp  g4 "-DRL4_SP_TANH_GROUP=4"  m4 "-DRL4_MINB_FP64=4" ...
    python scripts/variants.py build nl  base "" ...
    python scripts/variants.py run [--match g] scripts/prof_sp.py --policy fp64       # RL4AFCS_LIB selects the library

build_variants/ is git-ignored; it travels to the GPU box with the snapshot."""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rl4afcs_b200 import build as B  # noqa: E402

OUT = os.path.join(ROOT, "build_variants")
UNIT = {"sp": "sp_kernels.cu", "nl": "nl_kernels.cu"}


def build(unit, pairs):
    B.build()                                    # the regular objects (reused for the other units)
    os.makedirs(OUT, exist_ok=True)
    src = UNIT[unit]
    flags = [f for f in B.NVCC_FLAGS if f != "-shared"]
    procs = []
    for name, extra in pairs:
        obj = os.path.join(OUT, f"{unit}_{name}.o")
        cmd = [B._nvcc()] + flags + extra.split() + ["-c", "-o", obj, os.path.join(B.CSRC, src)]
        procs.append((name, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, obj, pr in procs:
        out, _ = pr.communicate()
        open(os.path.join(OUT, f"{unit}_{name}.log"), "w").write(out)
        if pr.returncode:
            print(f"{name}: compile FAILED\n{out[-2000:]}")
            continue
        objs = [obj if s == src else B._obj(s) for s in B.SOURCES]
        lib = os.path.join(OUT, f"librl4_{unit}_{name}.so")
        subprocess.run([B._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs, check=True)
        pat = r"sp_run_kernelIddLb0ELi0ELb0E" if unit == "sp" else r"nl_run_kernelIfLi1ELb0ELb0E"
        m = re.search(pat + r".*?\n.*?\n\s*(\d+ bytes stack frame, \d+ bytes spill stores, \d+ bytes spill loads)\n.*?Used (\d+) registers", out)
        print(f"{name}: {m.group(2) + ' regs, ' + m.group(1) if m else '?'}")


def run(match, script_and_args):
    for so in sorted(glob.glob(os.path.join(OUT, "librl4_*.so"))):
        if match and match not in os.path.basename(so):
            continue
        env = dict(os.environ, RL4AFCS_LIB=so)
        r = subprocess.run(["timeout", "120", sys.executable] + script_and_args, env=env, capture_output=True, text=True)
        tail = (r.stdout + r.stderr).strip().splitlines()[-3:]
        print(f"== {os.path.basename(so)}: " + " | ".join(tail), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        a = sys.argv[3:]
        build(sys.argv[2], list(zip(a[0::2], a[1::2])))
    else:
        a = sys.argv[2:]
        match = None
        if a and a[0] == "--match":
            match, a = a[1], a[2:]
        run(match, a)
