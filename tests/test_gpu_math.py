"""Arithmetic primitives of the kernels vs their CPU statements (through the C-ABI test hook)."""
import ctypes

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _probe(op, a, b=None):
    from rl4afcs_b200 import _lib

    L = _lib.load()
    ta = torch.as_tensor(a).cuda()
    tb = torch.as_tensor(b).cuda() if b is not None else None
    out = torch.empty_like(ta)
    _lib.check(L.rl4_test_math(op, ta.data_ptr(), tb.data_ptr() if tb is not None else None, out.data_ptr(),
                               ta.numel(), None), "rl4_test_math")
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _tanh_inputs(dtype, rng):
    parts = [rng.uniform(-20, 20, 200000), rng.uniform(-1, 1, 200000), rng.uniform(-1e-2, 1e-2, 200000),
             10.0 ** rng.uniform(-300 if dtype == np.float64 else -37, -2, 100000),
             np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 19.0624, 19.0625, 19.07, 9.124, 9.125, 9.2, 1e300, -1e300,
                       0.17328679, 0.1732868, 0.5198603, 5e-324, 1e-310, -1e-310])]
    return np.concatenate(parts).astype(dtype)


@pytest.mark.parametrize("dtype,op", [(np.float64, 0), (np.float32, 1)])
def test_device_tanh_t13_equals_oracle_t13_bitwise(oracle, dtype, op):
    x = _tanh_inputs(dtype, np.random.default_rng(0))
    if dtype == np.float32:
        x = x[np.isfinite(x) | np.isnan(x) | np.isinf(x)]
    got = _probe(op, x)
    want = oracle.tanh_t13(x)
    same = (got == want) | (np.isnan(got) & np.isnan(want))
    assert same.all(), (x[~same][:5], got[~same][:5], want[~same][:5])


def test_device_tanh_t13_accuracy_vs_libm():
    x = _tanh_inputs(np.float64, np.random.default_rng(1))
    x = x[np.isfinite(x)]
    got = _probe(0, x)
    want = np.tanh(x)
    ulp = np.abs(got - want) / np.maximum(np.spacing(np.abs(want)), 5e-324)
    assert ulp.max() <= 6.0, ulp.max()      # t13 <= 2.1 ulp of exact, np.tanh <= 3 ulp of libm


def test_shared_reciprocal_division_equals_ieee():
    rng = np.random.default_rng(2)
    n = 1 << 22
    a = rng.standard_normal(n) * 10.0 ** rng.uniform(-12, 12, n)
    b = rng.standard_normal(n) * 10.0 ** rng.uniform(-12, 12, n)
    # edge ranges: zeros, denormals, huge, inf, nan, exact quotients
    edge = np.array([0.0, -0.0, 5e-324, 1e-310, 2.0 ** -969, 2.0 ** -970, 1e-300, 1.0, 3.0, 1e300, 1.7e308, np.inf, -np.inf, np.nan])
    ea, eb = np.meshgrid(edge, edge)
    a = np.concatenate([a, ea.ravel(), rng.uniform(0, 2 ** 56, 100000)])
    b = np.concatenate([b, eb.ravel(), rng.uniform(2, 2 ** 56, 100000)])
    got = _probe(2, a, b)
    ieee = _probe(3, a, b)
    with np.errstate(all="ignore"):
        host = a / b
    for want in (ieee, host):
        same = (got == want) | (np.isnan(got) & np.isnan(want))
        assert same.all(), (a[~same][:5], b[~same][:5], got[~same][:5], want[~same][:5])


def test_float_t13_quotient_equals_ieee_for_every_float():
    """em/(em+2) re-spelled (MUFU.RCP + Newton + correction) == __fdiv_rn for all floats in [0, 2^28]."""
    from rl4afcs_b200 import _lib

    L = _lib.load()
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    hi = int(np.float32(2.0 ** 28).view(np.uint32)) + 1
    _lib.check(L.rl4_test_t13_div_f32(0, hi, cnt.data_ptr(), None), "rl4_test_t13_div_f32")
    torch.cuda.synchronize()
    assert int(cnt.item()) == 0


def test_double_t13_quotient_seed_is_the_ieee_float_reciprocal_for_every_float():
    """The double t13 quotient starts from the float32 reciprocal of the float32-rounded denominator; the kernel forms it as
    MUFU.RCP + one Newton step and the oracle as `1.0f / d`.  They must agree for EVERY float the denominator em + 2 can
    round to (2 <= d < 2^60; em < 2^56)."""
    from rl4afcs_b200 import _lib

    L = _lib.load()
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    lo = int(np.float32(2.0).view(np.uint32))
    hi = int(np.float32(2.0 ** 60).view(np.uint32))
    _lib.check(L.rl4_test_rcp_f32(lo, hi, cnt.data_ptr(), None), "rl4_test_rcp_f32")
    torch.cuda.synchronize()
    assert int(cnt.item()) == 0


def test_plant_sincos_and_atmosphere_equal_oracle_bitwise(oracle):
    """The surrogate plant's sine / cosine and ISA series (include/rl4_citation_surrogate.h) are IEEE basic operations
    only: device == host bit for bit, and accurate (sin / cos < 2 ulp of libm; density == ISA power law to 1e-15)."""
    import ctypes

    from oracle import nl_c
    from rl4afcs_b200 import _lib

    L = nl_c.lib()
    vp = ctypes.c_void_p
    L.orc_cit_sincos.argtypes = [vp, vp, vp, ctypes.c_int64]
    L.orc_cit_air.argtypes = [vp, vp, vp, vp, ctypes.c_int64]
    rng = np.random.default_rng(3)
    a = np.concatenate([rng.uniform(-0.8, 0.8, 300000), rng.uniform(-10, 10, 300000), rng.uniform(-1e4, 1e4, 100000),
                        rng.standard_normal(1000) * 1e12, [0.0, -0.0, np.pi / 4, np.pi / 2, np.pi, 1e-300, 0.0576, np.inf, np.nan]])
    s, c = np.empty_like(a), np.empty_like(a)
    L.orc_cit_sincos(a.ctypes.data, s.ctypes.data, c.ctypes.data, a.size)
    lib = _lib.load()
    d_a = torch.as_tensor(a).cuda()
    out = torch.empty_like(d_a)
    for op, want in ((4, s), (5, c)):
        _lib.check(lib.rl4_test_math(op, d_a.data_ptr(), None, out.data_ptr(), a.size, None), "rl4_test_math")
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), fin)                              # inf / nan inputs give nan on both sides
        bad = np.flatnonzero(got[fin].view(np.uint64) != want[fin].view(np.uint64))
        assert bad.size == 0, (op, a[fin][bad[:5]], got[fin][bad[:5]], want[fin][bad[:5]])
    small = np.abs(a) <= 1e4
    with np.errstate(invalid="ignore"):
        assert np.max(np.abs(s[small] - np.sin(a[small])) / np.spacing(np.abs(np.sin(a[small])) + 1e-300)) <= 2.0
        assert np.max(np.abs(c[small] - np.cos(a[small])) / np.spacing(np.abs(np.cos(a[small])) + 1e-300)) <= 2.0
    # atmosphere
    cfg = nl_c.make_cfg()
    plant = np.ascontiguousarray(cfg["plant"][:1])
    h = np.concatenate([rng.uniform(-2000, 11000, 200000), [0.0, 2000.0]])
    rho, lapse = np.empty_like(h), np.empty_like(h)
    L.orc_cit_air(plant.ctypes.data, h.ctypes.data, rho.ctypes.data, lapse.ctypes.data, h.size)
    d_p = torch.as_tensor(plant.view(np.uint8)).cuda()
    d_h = torch.as_tensor(h).cuda()
    out = torch.empty_like(d_h)
    for op, want in ((6, rho), (7, lapse)):
        _lib.check(lib.rl4_test_math(op, d_h.data_ptr(), d_p.data_ptr(), out.data_ptr(), h.size, None), "rl4_test_math")
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy().view(np.uint64), want.view(np.uint64))
    isa = 1.225 * (1 - 0.0065 * h / 288.15) ** (9.80665 / (0.0065 * 287.05) - 1)
    assert np.max(np.abs(rho - isa) / isa) < 2e-15 and np.max(np.abs(lapse - (isa / 1.225) ** 0.7) / lapse) < 2e-15
