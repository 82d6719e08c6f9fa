"""The machine model under the translated plant binary (include/rl4_lift_runtime.h: what each x86-64 instruction class does
to its destination and to ZF SF CF OF PF) against an independent big-integer model of the ISA written from the Intel SDM.
Runs anywhere (gcc only): the end-to-end check -- translation == binary -- is tests/test_citation_lifted.py."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def probe():
    out = os.path.join(HERE, "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "liblift_probe.so")
    subprocess.run(["gcc", "-O1", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, os.path.join(HERE, "lift_runtime_probe.c"), "-lm"], check=True)
    return ctypes.CDLL(so)


def parity(r):
    return int(bin(r & 0xff).count("1") % 2 == 0)


def model(op, w, a, b, cf):
    """(result, zf, sf, cf, of, pf); None for a flag the instruction leaves undefined / unchanged and the probe need not match"""
    m = (1 << w) - 1
    a &= m; b &= m
    sign = lambda v: (v >> (w - 1)) & 1  # noqa: E731
    sx = lambda v: v - (1 << w) if sign(v) else v  # noqa: E731
    if op in ("ADD", "ADC"):
        c = cf if op == "ADC" else 0
        full = a + b + c
        r = full & m
        return r, int(r == 0), sign(r), int(full > m), int(not (-(1 << (w - 1)) <= sx(a) + sx(b) + c < (1 << (w - 1)))), parity(r)
    if op in ("SUB", "SBB"):
        c = cf if op == "SBB" else 0
        r = (a - b - c) & m
        return r, int(r == 0), sign(r), int(a < b + c), int(not (-(1 << (w - 1)) <= sx(a) - sx(b) - c < (1 << (w - 1)))), parity(r)
    if op in ("AND", "OR", "XOR"):
        r = {"AND": a & b, "OR": a | b, "XOR": a ^ b}[op]
        return r, int(r == 0), sign(r), 0, 0, parity(r)
    if op == "INC":
        r = (a + 1) & m
        return r, int(r == 0), sign(r), cf, int(a == (1 << (w - 1)) - 1), parity(r)
    if op == "DEC":
        r = (a - 1) & m
        return r, int(r == 0), sign(r), cf, int(a == 1 << (w - 1)), parity(r)
    if op == "NEG":
        r = (-a) & m
        return r, int(r == 0), sign(r), int(a != 0), int(a == 1 << (w - 1)), parity(r)
    n = b & (63 if w == 64 else 31)
    if op in ("SHL", "SHR", "SAR"):
        if n == 0:
            return a, None, None, cf, None, None
        if op == "SHL":
            r = (a << n) & m
            c = ((a << n) >> w) & 1 if n <= w else 0
            return r, int(r == 0), sign(r), c, (sign(r) ^ c) if n == 1 else None, parity(r)
        if op == "SHR":
            r = a >> n
            c = (a >> (n - 1)) & 1 if n <= w else 0
            return r, int(r == 0), sign(r), c, sign(a) if n == 1 else None, parity(r)
        r = (sx(a) >> min(n, w - 1)) & m
        c = (sx(a) >> (min(n, w) - 1)) & 1
        return r, int(r == 0), sign(r), c, 0 if n == 1 else None, parity(r)
    if op in ("ROL", "ROR"):
        n %= w
        if n == 0:
            return a, None, None, None, None, None
        r = ((a << n) | (a >> (w - n))) & m if op == "ROL" else ((a >> n) | (a << (w - n))) & m
        return r, None, None, (r & 1) if op == "ROL" else sign(r), None, None
    raise KeyError(op)


@pytest.mark.parametrize("w", [8, 16, 32, 64])
def test_integer_instruction_semantics(probe, w):
    rng = np.random.default_rng(w)
    edge = [0, 1, 2, (1 << (w - 1)) - 1, 1 << (w - 1), (1 << (w - 1)) + 1, (1 << w) - 2, (1 << w) - 1, 0x55 & ((1 << w) - 1)]
    vals = edge + [int(v) & ((1 << w) - 1) for v in rng.integers(0, 1 << 63, 200, dtype=np.uint64) * 2 + rng.integers(0, 2, 200, dtype=np.uint64)]
    flags = (ctypes.c_uint8 * 5)()
    for op in ("ADD", "ADC", "SUB", "SBB", "AND", "OR", "XOR", "INC", "DEC", "NEG", "SHL", "SHR", "SAR", "ROL", "ROR"):
        fn = getattr(probe, f"probe_{op}{w}")
        fn.restype = ctypes.c_uint64
        fn.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p]
        shifts = op in ("SHL", "SHR", "SAR", "ROL", "ROR")
        for i, a in enumerate(vals):
            bs = [0, 1, 2, w - 1, w // 2, 31, 33 % w] if shifts else [vals[(i * 7 + 3) % len(vals)], vals[(i + 1) % len(vals)]]
            for b in bs:
                for cf in (0, 1):
                    r = fn(a, b, cf, flags)
                    want = model(op, w, a, b, cf)
                    assert r == want[0], (op, w, hex(a), hex(b), cf, hex(r), hex(want[0]))
                    for k, name in enumerate(("zf", "sf", "cf", "of", "pf")):
                        if want[1 + k] is not None:
                            assert flags[k] == want[1 + k], (op, w, hex(a), hex(b), cf, name, flags[k], want[1 + k])


def test_imul_truncates(probe):
    flags = (ctypes.c_uint8 * 5)()
    for w in (32, 64):
        fn = getattr(probe, f"probe_IMUL{w}")
        fn.restype = ctypes.c_uint64
        fn.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p]
        m = (1 << w) - 1
        for a, b in ((3, 5), (m, 2), (1 << (w - 1), 3), (0x12345678, 0x9abcdef), (m, m)):
            sa = a - (1 << w) if a >> (w - 1) else a
            sb = b - (1 << w) if b >> (w - 1) else b
            assert fn(a, b, 0, flags) == (sa * sb) & m


def test_comisd_flags_and_conversions(probe):
    probe.probe_comisd.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_void_p]
    flags = (ctypes.c_uint8 * 5)()
    for a, b, want in ((1.0, 2.0, (0, 0, 1, 0, 0)), (2.0, 1.0, (0, 0, 0, 0, 0)), (1.5, 1.5, (1, 0, 0, 0, 0)), (-0.0, 0.0, (1, 0, 0, 0, 0)),
                       (math.nan, 1.0, (1, 0, 1, 0, 1)), (1.0, math.nan, (1, 0, 1, 0, 1)), (math.inf, 1e308, (0, 0, 0, 0, 0))):
        probe.probe_comisd(a, b, flags)
        assert tuple(flags) == want, (a, b, tuple(flags))
    probe.probe_cvtt32.restype = probe.probe_cvtt64.restype = ctypes.c_uint64
    probe.probe_cvtt32.argtypes = probe.probe_cvtt64.argtypes = [ctypes.c_double]
    assert probe.probe_cvtt32(2.9) == 2 and probe.probe_cvtt32(-2.9) == 0xfffffffe and probe.probe_cvtt32(math.nan) == 0x80000000
    assert probe.probe_cvtt32(3e9) == 0x80000000 and probe.probe_cvtt32(-2147483648.0) == 0x80000000 and probe.probe_cvtt32(2147483647.9) == 0x7fffffff
    assert probe.probe_cvtt64(-1.5) == 0xffffffffffffffff and probe.probe_cvtt64(1e19) == 0x8000000000000000 and probe.probe_cvtt64(math.nan) == 0x8000000000000000
    # maxsd / minsd return the SECOND operand when either is NaN (and for equal operands, e.g. the two zeros)
    probe.probe_max.restype = probe.probe_min.restype = ctypes.c_double
    probe.probe_max.argtypes = probe.probe_min.argtypes = [ctypes.c_double, ctypes.c_double]
    assert probe.probe_max(1.0, 2.0) == 2.0 and probe.probe_max(2.0, 1.0) == 2.0 and probe.probe_min(1.0, 2.0) == 1.0
    assert math.isnan(probe.probe_max(1.0, math.nan)) and probe.probe_max(math.nan, 1.0) == 1.0 and probe.probe_min(math.nan, 3.0) == 3.0
    assert math.copysign(1.0, probe.probe_max(0.0, -0.0)) == -1.0 and math.copysign(1.0, probe.probe_min(-0.0, 0.0)) == 1.0
