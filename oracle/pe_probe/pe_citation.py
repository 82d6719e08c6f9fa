"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the in-process loader (pe_citation.c) and a stand-in for the
reference's ``extended_input.citation`` module backed by the REAL plant binary.

Only usable where /root/reference exists (the build container): the GPU box has the golden fixtures instead.
The model keeps process-global state, like the reference's module (envs/nonlinear/citation.py:62-69)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = "/root/reference/envs/nonlinear"
VARIANTS = ("extended_input", "nominal", "shift_cg", "shift_2")
_lib = None
_open_variant = None


def available(variant: str = "extended_input") -> bool:
    return os.path.isfile(binary_path(variant))


def binary_path(variant: str) -> str:
    return os.path.join(REF_DIR, variant, "_citation.cp39-win_amd64.pyd")


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
        L = ctypes.CDLL(os.path.join(_HERE, "..", "_ref", "libpe_citation.so"))
        L.pe_citation_error.restype = ctypes.c_char_p
        L.pe_citation_base.restype = ctypes.c_uint64
        _lib = L
    return _lib


def open_variant(variant: str = "extended_input") -> dict:
    """Maps the binary of `variant` (one image per process: the preferred base is taken by the first one)."""
    global _open_variant
    L = lib()
    if _open_variant is not None:
        if _open_variant != variant:
            raise RuntimeError(f"this process already mapped the '{_open_variant}' binary; use a fresh process for '{variant}'")
        return {}
    nb, nt = ctypes.c_int(), ctypes.c_int()
    rc = L.pe_citation_open(binary_path(variant).encode(), ctypes.byref(nb), ctypes.byref(nt))
    if rc:
        raise RuntimeError(f"pe_citation_open failed ({rc}): {L.pe_citation_error().decode()}")
    _open_variant = variant
    return {"imports_bound_to_glibc": nb.value, "imports_trapped": nt.value, "base": hex(L.pe_citation_base())}


def initialize():
    lib().pe_citation_initialize()


def terminate():
    lib().pe_citation_terminate()


def step(cmd) -> np.ndarray:
    """citation.step(cmd): cmd (11,) float64 -> state (12,) float64 (a new array, like the SWIG wrapper's)."""
    u = np.ascontiguousarray(cmd, dtype=np.float64)
    assert u.shape == (11,)
    out = np.empty(12)
    lib().pe_citation_step(u.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
    return out


def get_state():
    """(x[12], engine[4]): the model's 16 continuous states."""
    x, e = np.empty(12), np.empty(4)
    lib().pe_citation_get_state(x.ctypes.data_as(ctypes.c_void_p), e.ctypes.data_as(ctypes.c_void_p))
    return x, e


def set_state(x=None, engine=None):
    """Overwrites the continuous states (after the first step() following initialize(), which loads the initial condition)."""
    xp = None if x is None else np.ascontiguousarray(x, dtype=np.float64)
    ep = None if engine is None else np.ascontiguousarray(engine, dtype=np.float64)
    lib().pe_citation_set_state(None if xp is None else xp.ctypes.data_as(ctypes.c_void_p),
                                None if ep is None else ep.ctypes.data_as(ctypes.c_void_p))


def onestep(x, engine, u):
    """The one-step map of the real plant for n samples: x (n,12), engine (n,4), u (n,11) -> (x_next, engine_next)."""
    x = np.array(x, dtype=np.float64, order="C"); e = np.array(engine, dtype=np.float64, order="C")
    u = np.ascontiguousarray(u, dtype=np.float64)
    n = x.shape[0]
    assert x.shape == (n, 12) and e.shape == (n, 4) and u.shape == (n, 11)
    lib().pe_citation_onestep(x.ctypes.data_as(ctypes.c_void_p), e.ctypes.data_as(ctypes.c_void_p), u.ctypes.data_as(ctypes.c_void_p), n)
    return x, e


def run(cmd, n_steps: int) -> np.ndarray:
    u = np.ascontiguousarray(cmd, dtype=np.float64)
    out = np.empty((n_steps, 12))
    lib().pe_citation_run(u.ctypes.data_as(ctypes.c_void_p), int(n_steps), out.ctypes.data_as(ctypes.c_void_p))
    return out
