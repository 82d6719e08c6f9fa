"""The C-ABI library loads without a GPU and exports every symbol include/rl4afcs_b200.h declares;
the product refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "rl4afcs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rl4_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from rl4afcs_b200 import _lib, build

    build.build()
    L = _lib.load()
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), n
    assert sorted(_lib.EXPORTS) == names
    hdr = open(os.path.join(ROOT, "include", "rl4afcs_b200.h")).read()
    assert L.rl4_abi_version() == _lib.ABI_VERSION == int(re.search(r"#define RL4_ABI_VERSION (\d+)", hdr).group(1))
    # load() refuses a library whose struct layouts differ from the binding's
    assert L.rl4_sizeof_sp_params() == ctypes.sizeof(_lib.SpParams) and L.rl4_sizeof_nl_params() == ctypes.sizeof(_lib.NlParams)


def test_load_rejects_a_stale_library(monkeypatch):
    """A library built from an older header (other ABI version / struct sizes) must not load silently (by-value structs
    of another layout would corrupt kernel parameters)."""
    from rl4afcs_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "ABI_VERSION", _lib.ABI_VERSION + 1)
    with pytest.raises(_lib.Rl4Error, match="rebuild"):
        _lib.load()
    monkeypatch.setattr(_lib, "_lib", None)


def test_struct_layouts_match_header_constants():
    from rl4afcs_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "rl4afcs_b200.h")).read()
    for table, prefix in ((_lib.SPE, "RL4_SPE_"), (_lib.SPN, "RL4_SPN_"), (_lib.SPI, "RL4_SPI_"), (_lib.LF, "RL4_LF_"), (_lib.LB, "RL4_LB_"),
                          (_lib.NLL, "RL4_NLL_"), (_lib.NLF, "RL4_NLF_"), (_lib.OUT, "RL4_OUT_")):
        for k, v in table.items():
            m = re.search(prefix + k + r"\s*=\s*(\d+)", hdr)
            assert m and int(m.group(1)) == v, (prefix, k)
    for table, enum in ((_lib.HP, "rl4_sp_hp"), (_lib.HPI, "rl4_sp_hpi")):
        body = re.search(r"enum " + enum + r"\s*\{(.*?)\}", hdr, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = [t.strip().split("=")[0].strip() for t in body.split(",") if t.strip()]
        for i, nm in enumerate(names):
            key = nm.replace("RL4_HPI_", "").replace("RL4_HP_", "")
            assert table[key] == i, nm
    assert ctypes.sizeof(_lib.SpParams) == 8 * (16 + 8 + 1 + 14) + 4 * (8 + 2) + 8 * (14 + 8)
    for table, enum, prefix in ((_lib.SPS, "rl4_sp_stat_field", "RL4_SPS_"), (_lib.NLS, "rl4_nl_stat_field", "RL4_NLS_")):
        body = re.search(r"enum " + enum + r"\s*\{(.*?)\}", hdr, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = [t.strip().split("=")[0].strip() for t in body.split(",") if t.strip()]
        assert [table[nm.replace(prefix, "")] for nm in names] == list(range(len(names))), enum
    assert ctypes.sizeof(_lib.SpHostIO) == 8 * 9 + 8 and ctypes.sizeof(_lib.NlHostIO) == 8 * 11 + 8


def test_no_cpu_fallback():
    torch = pytest.importorskip("torch")
    from rl4afcs_b200 import _lib, sp_engine

    with pytest.raises(_lib.Rl4Error):
        sp_engine.SpEngine(4, policy="fp64", device="cpu")
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            sp_engine.SpEngine(4, policy="fp64", device="cuda")


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "rl4afcs_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("# oracle", ""), os.path.join(dp, f)


def test_nl_default_params_equal_oracle_defaults(oracle):
    """rl4_nl_default_params (host-only C entry point) == the oracle's default configuration, which
    tests/test_oracle_vs_reference.py pins to the verbatim idhp_nonlin.py; the plant block (trim solve, atmosphere
    series, reciprocals) must come out bit-identical from nvcc's host pass and from gcc."""
    from oracle import nl_c
    from rl4afcs_b200 import _lib

    L = _lib.load()
    p = _lib.NlParams()
    assert L.rl4_nl_default_params(ctypes.byref(p)) == 0
    cfg = nl_c.make_cfg()[0]
    for f in nl_c.PLANT_FIELDS + ["zeta_per_m"]:
        assert getattr(p.plant, f) == cfg["plant"][f], f
    assert list(p.plant.rho_poly) == list(cfg["plant"]["rho_poly"]) and list(p.plant.lapse_poly) == list(cfg["plant"]["lapse_poly"])
    assert list(p.trim_input) == list(cfg["trim_input"]) and p.dt == cfg["dt"]
    hp = _lib.NHP
    for key, name in (("ETA_A_H", "eta_a_h"), ("ETA_A_L", "eta_a_l"), ("ETA_C_H", "eta_c_h"), ("ETA_C_L", "eta_c_l"),
                      ("LAMBDA_H", "lambda_h"), ("LAMBDA_L", "lambda_l"), ("GAMMA", "gamma"), ("GAMMA_SQ", "gamma_sq"), ("TAU", "tau"),
                      ("LR_DECAY", "lr_decay"), ("RLS_GAMMA", "rls_gamma"), ("RLS_COV0", "rls_cov0"), ("Q_SYM", "Q_sym"),
                      ("LAMBDA_T", "lambda_t"), ("LAMBDA_S", "lambda_s"), ("DAMP_FACTOR", "damp_factor"), ("CG_SHIFT", "cg_shift")):
        assert p.hp[hp[key]] == cfg[name], key
    assert list(p.noise_std) == list(cfg["noise_std"]) and p.omega0 == cfg["omega0"] and p.omega_slow == cfg["omega_slow"]
    assert p.rate_limit == cfg["rate_limit"] and list(p.limit_deg) == list(cfg["limit_deg"]) and list(p.sat_limit) == list(cfg["sat_limit"])
    hi = _lib.NHPI
    assert p.hpi[hi["WARMUP_STEPS"]] == cfg["warmup_steps"] and p.hpi[hi["COOLDOWN_STEPS"]] == cfg["cooldown_steps"]
    assert p.hpi[hi["MULTISTEP"]] == cfg["multistep"] and p.hpi[hi["ELIG_A"]] == cfg["elig_a"] and p.hpi[hi["FAULT_STEP"]] == cfg["fault_step"]
    assert p.hpi[hi["FLIGHT_STEP"]] == cfg["flight_step"] == 5500 and p.hpi[hi["NUMPY2"]] == cfg["numpy2"] == 1 and p.integrator == cfg["integrator"]
