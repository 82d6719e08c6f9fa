"""TEST INFRASTRUCTURE ONLY -- loads the *verbatim* reference pieces that can run here.

The reference (wingos80/RL4AFCS, mounted read-only at /root/reference in the build
container; absent on the GPU box) cannot be imported as a whole: ``objects.py:29``
imports TensorFlow, ``envs/linear/env.py:2-4`` imports gymnasium + matplotlib, none
of which are installed.  Two pieces are pure numpy and run unmodified:

* ``Ce500ShortPeriod`` (``envs/linear/env.py:7-264``) -- imported from the file where
  it lies, with stub ``gymnasium`` / ``matplotlib`` modules registered first;
* ``RLS`` (``objects.py:439-549``) -- the ``ClassDef`` node is extracted with ``ast``
  and exec'd with ``{'np': numpy}``.

* ``Ce500NonLinear`` (``envs/nonlinear/env.py:11-319``), the WRAPPER only -- imported from the file where it lies with a
  stand-in for its ``extended_input.citation`` module (the reference's plant is a source-less Windows binary): the
  stand-in's ``initialize / step / terminate`` drive the documented surrogate plant through the C oracle, so the
  verbatim action scaling, actuator, fault, reward and MDP-state code runs unmodified around it.
* ``utils.py`` (samplers, ``get_PSD``, ``get_convergence_time``, ``VD_A``, ``kl_divergence``) -- imported from the
  file where it lies with the matplotlib stub (scipy is installed).

Nothing is copied into this repo; the source is read from /root/reference at call
time.  Used by ``oracle/make_golden.py`` and ``tests/test_oracle_vs_reference.py``
(skipped when /root/reference is absent).  Never imported by the product package.
"""
from __future__ import annotations

import ast
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("RL4AFCS_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "envs", "linear", "env.py"))


def _install_stubs() -> None:
    if "gymnasium" not in sys.modules:
        gym = types.ModuleType("gymnasium")

        class Env:  # minimal stand-in for gymnasium.Env
            def reset(self, seed=None, options=None):
                return None

        gym.Env = Env
        spaces = types.ModuleType("gymnasium.spaces")
        spaces.Box = object
        gym.spaces = spaces
        sys.modules["gymnasium"] = gym
        sys.modules["gymnasium.spaces"] = spaces
    if "tqdm" not in sys.modules or not getattr(sys.modules["tqdm"], "_rl4_stub", False):
        tq = types.ModuleType("tqdm")

        class tqdm:  # silent stand-in: iterable with set_description
            def __init__(self, it=None, **kw):
                self._it = it

            def __iter__(self):
                return iter(self._it)

            def set_description(self, *a, **k):
                pass

        tq.tqdm = tqdm
        tq._rl4_stub = True
        sys.modules["tqdm"] = tq
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        plt.rcParams = {}
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt


def load_reference_linear_env():
    """Return the verbatim ``Ce500ShortPeriod`` class (envs/linear/env.py:7)."""
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    import importlib.util

    path = os.path.join(REFERENCE_ROOT, "envs", "linear", "env.py")
    spec = importlib.util.spec_from_file_location("_rl4afcs_ref_linear_env", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.Ce500ShortPeriod


def load_reference_rls():
    """Return the verbatim ``RLS`` class (objects.py:439-549)."""
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    import numpy as np

    path = os.path.join(REFERENCE_ROOT, "objects.py")
    with open(path, "r", encoding="utf-8") as fh:
        tree = ast.parse(fh.read(), filename=path)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == "RLS":
            ns = {"np": np}
            code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
            exec(code, ns)
            return ns["RLS"]
    raise RuntimeError("class RLS not found in reference objects.py")


def load_reference_utils():
    """Return the verbatim ``utils`` module (utils.py: get_PSD :188, samplers :238,:293, get_convergence_time :350,
    VD_A :391, kl_divergence :436)."""
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    import importlib.util

    path = os.path.join(REFERENCE_ROOT, "utils.py")
    spec = importlib.util.spec_from_file_location("_rl4afcs_ref_utils", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_nonlinear_env(integrator: str = "ode5", plant: str = "surrogate"):
    """Return (verbatim ``Ce500NonLinear`` class, plant stand-in module).  envs/nonlinear/env.py:5-9 imports
    ``extended_input.citation`` -- a SWIG wrapper of a Windows DLL -- so a stand-in module with the same three functions
    (envs/nonlinear/citation.py:62-69: ``initialize()``, ``step(cmd) -> state[12]``, ``terminate()``; process-global
    state like the original) is registered first.  ``plant="surrogate"`` integrates the surrogate of
    include/rl4_citation_surrogate.h with the C oracle, with the binary's output-then-update timing (``step`` returns
    the state BEFORE the step); ``plant="binary"`` runs the reference's REAL plant binary in-process
    (oracle/pe_probe/pe_citation.py) -- the aircraft the reference actually flies."""
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    import ctypes
    import importlib.util

    import numpy as np

    from . import nl_c

    L = nl_c.lib()
    L.orc_cit_plant_step.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_int]
    plant_name = plant
    plant = np.ascontiguousarray(nl_c.make_cfg()["plant"][:1])
    stub = types.ModuleType("extended_input.citation")
    stub._x = None
    stub._integrator = nl_c.INTEGRATOR[integrator]
    stub.dt = 0.01

    def initialize():
        stub._x = np.array([0, 0, 0, 90, 0.0576, 0, 0, 0.0576, 0, 2000, 0, 0], dtype=np.float64)   # built-in initial state

    def step(cmd):
        u = np.ascontiguousarray(cmd, dtype=np.float64)
        assert u.shape == (11,)
        out = stub._x.copy()                      # output-then-update, like the binary (oracle/pe_probe/README.md)
        L.orc_cit_plant_step(plant.ctypes.data, stub._x.ctypes.data, u.ctypes.data, stub.dt, stub._integrator)
        return out

    def terminate():
        stub._x = None

    stub.initialize, stub.step, stub.terminate = initialize, step, terminate
    if plant_name == "binary":
        from .pe_probe import pe_citation

        pe_citation.open_variant("extended_input")
        stub.initialize, stub.step, stub.terminate = pe_citation.initialize, pe_citation.step, pe_citation.terminate
    pkg = types.ModuleType("extended_input")
    pkg.citation = stub
    pkg.__path__ = []
    sys.modules["extended_input"] = pkg
    sys.modules["extended_input.citation"] = stub
    path = os.path.join(REFERENCE_ROOT, "envs", "nonlinear", "env.py")
    spec = importlib.util.spec_from_file_location("_rl4afcs_ref_nonlinear_env", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.Ce500NonLinear, stub


def load_reference_objects():
    """Return the verbatim ``objects`` module (objects.py: Network, Critic, Actor, Critic_big, Actor_big, RLS, IDHPsp,
    IDHPnonlin) executed on the TensorFlow stand-in of oracle/tf_shim.py (TensorFlow itself cannot be installed here).
    The stand-in supplies only the arithmetic of the TF ops (DESIGN.md section 3 contract); every line of control flow,
    call order, aliasing and numpy-side arithmetic is the reference's own."""
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    import importlib.util

    from . import tf_shim

    tf_shim.install()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)              # objects.py does `from utils import *`
    path = os.path.join(REFERENCE_ROOT, "objects.py")
    spec = importlib.util.spec_from_file_location("_rl4afcs_ref_objects", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, tf_shim


def load_reference_functions():
    """Return the verbatim ``functions`` module (functions.py: MC_run_seed :39, MC_run :62, MC_test_hparam :931) on the
    TensorFlow stand-in; its ``from objects import ...`` picks up the verbatim objects.py the same way."""
    load_reference_objects()                            # installs the stand-in, stubs and sys.path
    import importlib.util

    if not hasattr(sys.modules["matplotlib.pyplot"], "rcParams"):
        sys.modules["matplotlib.pyplot"].rcParams = {}
    path = os.path.join(REFERENCE_ROOT, "functions.py")
    spec = importlib.util.spec_from_file_location("_rl4afcs_ref_functions", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
