"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the TRANSLATED plant binary (oracle/_ref/libcitation_lifted_<variant>.so,
built by oracle/pe_probe/build_lifted.sh (translator: rl4afcs_b200/tools/lift_plant.py) from the reference's .pyd where /root/reference exists; the compiled library and the
image file travel to the GPU box with oracle/_ref/).  Unlike pe_citation (the binary executing natively, one process-global
model) this is re-entrant: any number of aircraft, any host architecture."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "..", "_ref")
_libs = {}


def _paths(variant):
    return (os.path.join(_REF, f"libcitation_lifted_{variant}.so"), os.path.join(_REF, "lifted", f"citation_{variant}_image.bin"))


def available(variant: str = "extended_input") -> bool:
    so, img = _paths(variant)
    if os.path.isfile(so) and os.path.isfile(img):
        return True
    return os.path.isfile(f"/root/reference/envs/nonlinear/{variant}/_citation.cp39-win_amd64.pyd")


def lib(variant: str = "extended_input"):
    if variant not in _libs:
        so, img = _paths(variant)
        if not (os.path.isfile(so) and os.path.isfile(img)):
            subprocess.run([os.path.join(_HERE, "build_lifted.sh"), variant], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        L = ctypes.CDLL(so)
        L.cit_lifted_create.restype = ctypes.c_void_p
        for f in ("cit_lifted_initialize", "cit_lifted_terminate", "cit_lifted_destroy"):
            getattr(L, f).argtypes = [ctypes.c_void_p]
        L.cit_lifted_step.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.cit_lifted_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.cit_lifted_get_state.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.cit_lifted_set_state.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.cit_lifted_onestep.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        blob = open(img, "rb").read()
        assert L.cit_lifted_set_image(blob, ctypes.c_uint64(len(blob))) == 0
        _libs[variant] = L
    return _libs[variant]


class Aircraft:
    """One instance of the model: initialize() / step(cmd) -> state, as envs/nonlinear/citation.py:62-69."""

    def __init__(self, variant: str = "extended_input"):
        self.L = lib(variant)
        self.h = ctypes.c_void_p(self.L.cit_lifted_create())
        assert self.h.value

    def initialize(self):
        self.L.cit_lifted_initialize(self.h)

    def terminate(self):
        self.L.cit_lifted_terminate(self.h)

    def step(self, cmd) -> np.ndarray:
        u = np.ascontiguousarray(cmd, dtype=np.float64)
        assert u.shape == (11,)
        out = np.empty(12)
        self.L.cit_lifted_step(self.h, u.ctypes.data, out.ctypes.data)
        return out

    def run(self, cmd, n_steps: int) -> np.ndarray:
        u = np.ascontiguousarray(cmd, dtype=np.float64)
        out = np.empty((int(n_steps), 12))
        self.L.cit_lifted_run(self.h, u.ctypes.data, int(n_steps), out.ctypes.data)
        return out

    def get_state(self):
        x, e = np.empty(12), np.empty(4)
        self.L.cit_lifted_get_state(self.h, x.ctypes.data, e.ctypes.data)
        return x, e

    def set_state(self, x=None, engine=None):
        xp = None if x is None else np.ascontiguousarray(x, dtype=np.float64)
        ep = None if engine is None else np.ascontiguousarray(engine, dtype=np.float64)
        self.L.cit_lifted_set_state(self.h, None if xp is None else xp.ctypes.data, None if ep is None else ep.ctypes.data)

    def __del__(self):
        try:
            self.L.cit_lifted_destroy(self.h)
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass
