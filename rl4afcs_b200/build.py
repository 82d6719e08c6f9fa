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
SOURCES = ["runtime.cu", "sp_kernels.cu", "nl_kernels.cu", "step_kernels.cu", "host_episode.cu", "dasmat_plant.cu"]
HEADERS = ["rl4_math.cuh", "sp_core.cuh", "nl_pipeline.cuh", "nl_core.cuh", "rl4_runtime.h", os.path.join("..", "..", "include", "rl4afcs_b200.h"),
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


# headers each translation unit depends on (a stale object is rebuilt, the others are reused: the nonlinear kernels take
# over a minute to compile, the rest seconds)
_COMMON = ["rl4_math.cuh", "rl4_runtime.h", os.path.join("..", "..", "include", "rl4afcs_b200.h"),
           os.path.join("..", "..", "include", "rl4_citation_surrogate.h")]
DEPS = {"runtime.cu": _COMMON, "sp_kernels.cu": _COMMON + ["sp_core.cuh"], "nl_kernels.cu": _COMMON + ["nl_pipeline.cuh", "nl_core.cuh"],
        "step_kernels.cu": _COMMON, "host_episode.cu": _COMMON,
        "dasmat_plant.cu": _COMMON + ["nl_core.cuh",
                            os.path.join("..", "..", "include", "rl4_lift_runtime.h")]}

# The 'dasmat' plant is the reference's own aircraft model, translated from its binary (tools/lift_plant.py).  Nothing
# derived from the binary is committed: csrc/_gen/ is produced here, where the reference exists, and compiled into the
# library; without it dasmat_plant.cu compiles to entry points that report "built without the reference's plant binary".
GEN_DIR = os.path.join(CSRC, "_gen")
PLANT_BINARY = "/root/reference/envs/nonlinear/extended_input/_citation.cp39-win_amd64.pyd"
LIFTER = os.path.join(_HERE, "tools", "lift_plant.py")


def _generate_plant() -> None:
    code, image = os.path.join(GEN_DIR, "dasmat_code_step.inc"), os.path.join(GEN_DIR, "dasmat_image.inc")
    if not os.path.isfile(PLANT_BINARY) or not os.path.isfile(LIFTER):
        return
    if os.path.isfile(code) and os.path.isfile(image) and os.path.getmtime(code) >= os.path.getmtime(LIFTER):
        return
    subprocess.run([sys.executable, LIFTER, "--variant", "extended_input", "--out", GEN_DIR], check=True, stdout=subprocess.DEVNULL)
    for tag in ("code_step", "code_init", "image"):
        shutil.copyfile(os.path.join(GEN_DIR, f"citation_extended_input_{tag}.inc"), os.path.join(GEN_DIR, f"dasmat_{tag}.inc"))
OBJ_DIR = os.path.join(_HERE, "build")


def _obj(src: str) -> str:
    return os.path.join(OBJ_DIR, src.replace(".cu", ".o"))


def _src_stale(src: str) -> bool:
    o = _obj(src)
    if not os.path.isfile(o):
        return True
    t = os.path.getmtime(o)
    deps = [os.path.join(CSRC, d) for d in [src] + DEPS[src]] + [os.path.abspath(__file__)]
    if src == "dasmat_plant.cu":
        deps = deps[:-1] + [g for g in (os.path.join(GEN_DIR, "dasmat_code_step.inc"), os.path.join(GEN_DIR, "dasmat_code_init.inc"), os.path.join(GEN_DIR, "dasmat_image.inc")) if os.path.isfile(g)]
    return any(os.path.getmtime(d) > t for d in deps)


def _stale() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: str = None) -> str:
    """Build the library: every stale translation unit is compiled to rl4afcs_b200/build/*.o (in parallel), then linked.
    ``extra_flags`` / ``out`` exist for tuning experiments (scripts/): they compile everything into one private library."""
    out = out or LIB_PATH
    _generate_plant()
    if not force and out == LIB_PATH and not _stale() and not _src_stale("dasmat_plant.cu"):
        return LIB_PATH
    log = os.path.join(_HERE, "build.log")
    if extra_flags or out != LIB_PATH:
        cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags) + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
        res = subprocess.run(cmd, capture_output=True, text=True)
        text = " ".join(cmd) + "\n" + res.stdout + res.stderr
        rc = res.returncode
    else:
        os.makedirs(OBJ_DIR, exist_ok=True)
        todo = [s for s in SOURCES if force or _src_stale(s)]
        compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]
        procs = []
        for s in todo:
            cmd = [_nvcc()] + compile_flags + ["-c", "-o", _obj(s), os.path.join(CSRC, s)]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        text, rc = "", 0
        for cmd, pr in procs:
            o, _ = pr.communicate()
            text += " ".join(cmd) + "\n" + o
            rc = rc or pr.returncode
        if rc == 0:
            cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + [_obj(s) for s in SOURCES]
            res = subprocess.run(cmd, capture_output=True, text=True)
            text += " ".join(cmd) + "\n" + res.stdout + res.stderr
            rc = res.returncode
        # keep the ptxas summaries of the units that were NOT recompiled in the log
        old = {}
        if os.path.isfile(log) and todo != SOURCES:
            try:
                cur = None
                for line in open(log):
                    if line.startswith(_nvcc()) and " -c " in line:
                        cur = line.rsplit("/", 1)[-1].strip()
                        old[cur] = line
                    elif cur is not None:
                        old[cur] += line
            except OSError:
                old = {}
        for s in SOURCES:
            if s not in todo and s in old:
                text += old[s]
    with open(log, "w") as fh:
        fh.write(text)
    if verbose:
        print(text)
    if rc != 0:
        raise RuntimeError(f"nvcc failed ({rc}); see {log}\n{text[-4000:]}")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
