"""Builds librl4afcs_b200.so in-tree with nvcc for sm_100a (the only target).

    python -m rl4afcs_b200.build [--force] [--verbose]

Flags that matter for parity (DESIGN.md "Arithmetic contract"): ``-fmad=false`` (no implicit
FMA contraction; the kernels spell every FMA), default ``-prec-div=true -prec-sqrt=true
-ftz=false`` (IEEE division / sqrt, denormals kept), ``-lineinfo`` for ncu source pages.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "librl4afcs_b200.so")
SOURCES = ["runtime.cu", "sp_kernels.cu", "nl_kernels.cu", "step_kernels.cu", "host_episode.cu"]
HEADERS = ["rl4_math.cuh", "sp_core.cuh", "rl4_runtime.h", os.path.join("..", "..", "include", "rl4afcs_b200.h"),
           os.path.join("..", "..", "include", "rl4_citation_surrogate.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared",     # host pass: no contraction either (trim solve / reset)
    "-Xptxas", "-v",
    "--threads", "4",          # the three translation units compile in parallel
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found; rl4afcs_b200 has no prebuilt or CPU fallback")


def _stale() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: str = None) -> str:
    """Build the library.  ``extra_flags`` / ``out`` exist for tuning experiments (scripts/)."""
    out = out or LIB_PATH
    if not force and out == LIB_PATH and not _stale():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags) + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(_HERE, "build.log")
    with open(log, "w") as fh:
        fh.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed ({res.returncode}); see {log}\n{res.stderr[-4000:]}")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
