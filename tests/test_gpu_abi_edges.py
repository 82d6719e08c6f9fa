"""Edge cases of the C-ABI on the GPU: empty and ragged batches, stride > n, argument validation, error text."""
import ctypes

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from tests import _util  # noqa: E402


def _engine(oracle, n, policy="mixed"):
    from rl4afcs_b200 import sp_engine

    eng = sp_engine.SpEngine(n, policy=policy)
    ic = oracle.default_idhp_config()
    sp_engine.apply_idhp_config(eng, ic, dt=0.02)
    base, amp = oracle.default_reference()
    eng.set_hp("REF_AMP", amp); eng.set_hpi("FAULT_STEP", -1); eng.set_hpi("FAULT_KIND", 0)
    eng.set_reference(base)
    return eng, ic, base, amp


def test_empty_batch_is_a_no_op(oracle):
    eng, ic, base, amp = _engine(oracle, 0)
    eng.init(np.zeros((0, 2)), np.zeros((0, 4)), np.zeros((0, 4)), np.zeros((0, 4)), np.zeros((0, 8)))
    eng.run(10)
    assert eng.stats()["sum_c"].numel() == 0


@pytest.mark.parametrize("n", [1, 31, 33, 127, 129, 1000])
def test_ragged_batch_sizes(oracle, n):
    eng, ic, base, amp = _engine(oracle, n, "fp64")
    rng = np.random.default_rng(n)
    x0 = np.deg2rad(rng.uniform(-2, 2, size=(n, 2)))
    w = oracle.init_weights(n, n)
    eng.init(x0, w["W1a"], w["W2a"], w["W1c"], w["W2c"])
    eng.run(120)
    cfg = oracle.make_cfg(ic)
    st = oracle.init_states("fp64", cfg, x0, w)
    oracle.run("fp64", cfg, base, st, 0, 120, tanh="t13")
    got = _util.engine_state_to_oracle(eng, oracle, lambda low: np.where(low, cfg["lambda_l"] * cfg["gamma"], cfg["lambda_h"] * cfg["gamma"]))
    assert _util.state_mismatches(got, st) == {}


def test_stride_larger_than_batch_leaves_padding_untouched(oracle):
    from rl4afcs_b200 import _lib
    from rl4afcs_b200._lib import SPE, SPI, SPN

    eng, ic, base, amp = _engine(oracle, 40)
    L = eng.lib
    n, stride = 40, 64
    env = torch.full((SPE["COUNT"], stride), 7.0, dtype=torch.float64, device="cuda")
    net = torch.full((SPN["COUNT"], stride), 7.0, dtype=torch.float32, device="cuda")
    ints = torch.full((SPI["COUNT"], stride), 7, dtype=torch.int32, device="cuda")
    st = _lib.SpState(env.data_ptr(), net.data_ptr(), ints.data_ptr(), stride)
    z = torch.zeros((8, n), dtype=torch.float64, device="cuda")
    _lib.check(L.rl4_sp_init(_lib.MIXED, ctypes.byref(eng.params), z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(),
                             z.data_ptr(), n, st, n, None), "init")
    _lib.check(L.rl4_sp_run(_lib.MIXED, ctypes.byref(eng.params), eng.ref_base.data_ptr(), 0, 50, st, n, 0,
                            _lib.SpLog(None, 0, 1, 0), None), "run")
    torch.cuda.synchronize()
    assert bool((env[:, n:] == 7.0).all()) and bool((net[:, n:] == 7.0).all()) and bool((ints[:, n:] == 7).all())
    assert not bool((env[0, :n] == 7.0).any())


def test_argument_validation_and_error_text(oracle):
    from rl4afcs_b200 import _lib

    eng, ic, base, amp = _engine(oracle, 8)
    L = eng.lib
    st = eng.state_struct()
    nolog = _lib.SpLog(None, 0, 1, 0)
    assert L.rl4_sp_run(99, ctypes.byref(eng.params), eng.ref_base.data_ptr(), 0, 5, st, 8, 0, nolog, None) == -1
    assert b"policy" in L.rl4_last_error()
    assert L.rl4_sp_run(_lib.MIXED, ctypes.byref(eng.params), None, 0, 5, st, 8, 0, nolog, None) == -1
    bad = _lib.SpState(st.env, st.net, st.ints, 4)           # stride < n
    assert L.rl4_sp_run(_lib.MIXED, ctypes.byref(eng.params), eng.ref_base.data_ptr(), 0, 5, bad, 8, 0, nolog, None) == -1
    assert b"stride" in L.rl4_last_error()
    # traces configured but the trace-free kernel requested
    eng.set_hpi("ELIG_A", 1)
    assert L.rl4_sp_run(_lib.MIXED, ctypes.byref(eng.params), eng.ref_base.data_ptr(), 0, 5, st, 8, 0, nolog, None) == -1
    assert b"elig" in L.rl4_last_error()
    with pytest.raises(_lib.Rl4Error):
        _lib.check(-1, "demo")
    # nonlinear path: fp32 policy is rejected loudly
    from rl4afcs_b200 import nl_engine
    with pytest.raises(_lib.Rl4Error):
        nl_engine.NlEngine(4, policy="fp32")


def test_reference_table_bounds_are_checked(oracle):
    eng, ic, base, amp = _engine(oracle, 4)
    eng.init(np.zeros((4, 2)), np.zeros((4, 4)), np.zeros((4, 4)), np.zeros((4, 4)), np.zeros((4, 8)))
    with pytest.raises(AssertionError):
        eng.run(len(base) + 1)


@pytest.mark.parametrize("n", [1000, 200000])
def test_host_buffer_episode_equals_device_path(oracle, n):
    """rl4_sp_episode_host (chunked, multi-stream H2D / kernels / D2H) returns exactly what init + run produce."""
    from rl4afcs_b200 import _lib
    from rl4afcs_b200._lib import SPE, SPI, SPN

    steps = 150
    eng, ic, base, amp = _engine(oracle, n, "mixed")
    L = eng.lib
    rng = np.random.default_rng(3)
    x0 = np.deg2rad(rng.uniform(-2, 2, size=(n, 2)))
    w = oracle.init_weights(n, 5)
    eng.init(x0, w["W1a"], w["W2a"], w["W1c"], w["W2c"])
    eng.run(steps)
    pin = lambda a: torch.as_tensor(np.ascontiguousarray(a.T)).pin_memory()     # noqa: E731  [rows][n]
    hx0, h1, h2, h3, h4 = pin(x0), pin(w["W1a"]), pin(w["W2a"]), pin(w["W1c"]), pin(w["W2c"])
    href = torch.as_tensor(base[:steps].copy()).pin_memory()
    oenv = torch.empty((SPE["COUNT"], n), dtype=torch.float64).pin_memory()
    onet = torch.empty((SPN["COUNT"], n), dtype=torch.float32).pin_memory()
    oint = torch.empty((SPI["COUNT"], n), dtype=torch.int32).pin_memory()
    ctx = ctypes.c_void_p()
    _lib.check(L.rl4_ctx_create(0, _lib.MIXED, n, steps, ctypes.byref(ctx)), "ctx")
    io = _lib.SpHostIO(hx0.data_ptr(), h1.data_ptr(), h2.data_ptr(), h3.data_ptr(), h4.data_ptr(), href.data_ptr(),
                       oenv.data_ptr(), onet.data_ptr(), oint.data_ptr())
    try:
        _lib.check(L.rl4_sp_episode_host(ctx, ctypes.byref(eng.params), ctypes.byref(io), n, steps, 0), "episode")
    finally:
        L.rl4_ctx_destroy(ctx)
    eq = lambda u, v: bool(((u == v) | (torch.isnan(u) & torch.isnan(v))).all())   # noqa: E731
    assert eq(oenv, eng.env[:, :n].cpu()) and eq(onet, eng.net[:, :n].cpu()) and torch.equal(oint, eng.ints[:, :n].cpu())


def test_ctx_create_reports_allocation_failure_and_cleans_up():
    """A capacity no GPU can hold: rl4_ctx_create returns the CUDA error (positive code, text in rl4_last_error),
    leaves *out NULL, frees whatever it had created, and the library keeps working afterwards."""
    from rl4afcs_b200 import _lib

    L = _lib.load()
    free0 = torch.cuda.mem_get_info()[0]
    ctx = ctypes.c_void_p()
    rc = L.rl4_ctx_create(0, _lib.FP64, 1 << 36, 3000, ctypes.byref(ctx))
    assert rc > 0 and not ctx.value
    assert b"cudaMalloc" in L.rl4_last_error()
    torch.cuda.synchronize()
    assert torch.cuda.mem_get_info()[0] >= free0 - (64 << 20)                 # nothing of the failed context is left behind
    ok = ctypes.c_void_p()
    _lib.check(L.rl4_ctx_create(0, _lib.FP64, 1024, 100, ctypes.byref(ok)), "ctx")
    assert ok.value
    L.rl4_ctx_destroy(ok)


def test_host_buffer_episode_output_mask(oracle):
    """rl4_sp_host_io.out_mask: only the selected field groups travel back (same full-plane host layout, unselected rows
    untouched); the default selection of bench.py's end-to-end leg is statistics + weights + RLS model."""
    from rl4afcs_b200 import _lib
    from rl4afcs_b200._lib import OUT, SPE, SPI, SPN

    n, steps = 70000, 60
    eng, ic, base, amp = _engine(oracle, n, "fp64")
    L = eng.lib
    rng = np.random.default_rng(4)
    x0 = np.deg2rad(rng.uniform(-2, 2, size=(n, 2)))
    w = oracle.init_weights(n, 6)
    eng.init(x0, w["W1a"], w["W2a"], w["W1c"], w["W2c"])
    eng.run(steps)
    pin = lambda a: torch.as_tensor(np.ascontiguousarray(a.T)).pin_memory()     # noqa: E731
    hx0, h1, h2, h3, h4 = pin(x0), pin(w["W1a"]), pin(w["W2a"]), pin(w["W1c"]), pin(w["W2c"])
    href = torch.as_tensor(base[:steps].copy()).pin_memory()
    ctx = ctypes.c_void_p()
    _lib.check(L.rl4_ctx_create(0, _lib.FP64, n, steps, ctypes.byref(ctx)), "ctx")
    sel = {"STATS": ([("SUM_C", 2)], [], True), "WEIGHTS": ([], [("W1A", 20)], False), "TARGET": ([], [("W1T", 12)], False),
           "RLS": ([("THETA", 15), ("EPS", 3)], [], False), "STATE": ([("X", 4), ("CGRAD_PREV", 1)], [("A", 2), ("MPREV", 6)], True),
           "TRACES": ([("EA", 20)], [], False)}
    try:
        for groups in (("STATS",), ("STATS", "WEIGHTS", "RLS"), ("STATE", "TRACES", "TARGET"), tuple(sel)):
            oenv = torch.full((SPE["COUNT"], n), -7.0, dtype=torch.float64).pin_memory()
            onet = torch.full((SPN["COUNT"], n), -7.0, dtype=torch.float64).pin_memory()
            oint = torch.full((SPI["COUNT"], n), -7, dtype=torch.int32).pin_memory()
            mask = sum(OUT[g] for g in groups)
            io = _lib.SpHostIO(hx0.data_ptr(), h1.data_ptr(), h2.data_ptr(), h3.data_ptr(), h4.data_ptr(), href.data_ptr(),
                               oenv.data_ptr(), onet.data_ptr(), oint.data_ptr(), mask, 0)
            _lib.check(L.rl4_sp_episode_host(ctx, ctypes.byref(eng.params), ctypes.byref(io), n, steps, 0), "episode")
            env_rows = np.zeros(SPE["COUNT"], bool); net_rows = np.zeros(SPN["COUNT"], bool); ints = False
            for g in groups:
                for f, c in sel[g][0]:
                    env_rows[SPE[f]:SPE[f] + c] = True
                for f, c in sel[g][1]:
                    net_rows[SPN[f]:SPN[f] + c] = True
                ints |= sel[g][2]
            eq = lambda u, v: bool(((u == v) | (torch.isnan(u) & torch.isnan(v))).all())   # noqa: E731
            er, nr = torch.as_tensor(env_rows), torch.as_tensor(net_rows)
            assert eq(oenv[er], eng.env[:, :n].cpu()[er]) and bool((oenv[~er] == -7.0).all()), groups
            assert eq(onet[nr], eng.net[:, :n].cpu()[nr]) and bool((onet[~nr] == -7.0).all()), groups
            assert torch.equal(oint, eng.ints[:, :n].cpu()) if ints else bool((oint == -7).all()), groups
    finally:
        L.rl4_ctx_destroy(ctx)


def test_nonlinear_noise_stream_is_deterministic_standard_normal():
    """rl4_nl_noise_fill: the Philox / Box-Muller N(0,1) stream is a pure function of (seed, step, agent) -- any split
    of the steps or the agents reproduces the same numbers -- with the moments of a standard normal."""
    from rl4afcs_b200 import _lib

    L = _lib.load()
    n, steps = 5000, 64
    full = torch.empty((steps, n), dtype=torch.float32, device="cuda")
    _lib.check(L.rl4_nl_noise_fill(1234, 0, 0, steps, n, full.data_ptr(), n, None), "noise")
    part = torch.empty((24, n - 1000), dtype=torch.float32, device="cuda")
    _lib.check(L.rl4_nl_noise_fill(1234, 1000, 40, 24, n - 1000, part.data_ptr(), n - 1000, None), "noise")   # agents 1000.., steps 40..
    other = torch.empty((steps, n), dtype=torch.float32, device="cuda")
    _lib.check(L.rl4_nl_noise_fill(1235, 0, 0, steps, n, other.data_ptr(), n, None), "noise")
    torch.cuda.synchronize()
    assert torch.equal(part, full[40:, 1000:]) and not torch.equal(other, full)
    assert bool(torch.isfinite(full).all())
    assert abs(float(full.mean())) < 0.01 and abs(float(full.std()) - 1.0) < 0.01
    assert abs(float((full.abs() < 1.0).float().mean()) - 0.6827) < 0.01 and float(full.abs().max()) > 3.5
    assert abs(float((full[1:] * full[:-1]).mean())) < 0.01 and abs(float((full[:, 1:] * full[:, :-1]).mean())) < 0.01


@pytest.mark.parametrize("noise_on_host", [True, False])
def test_nonlinear_host_buffer_episode_equals_device_path(noise_on_host):
    """rl4_nl_episode_host (IDHPnonlin(...).train() for a batch from HOST buffers: chunked over agents and steps, noise
    either supplied as a host plane or drawn on the device by rl4_nl_noise_fill) returns exactly what rl4_nl_init +
    rl4_nl_run produce on device buffers with the same noise."""
    from oracle import nl_c
    from rl4afcs_b200 import _lib, nl_engine
    from rl4afcs_b200._lib import NLE, NLI, NLN, OUT

    n, steps, seed = 70000, 320, 99
    eng = nl_engine.NlEngine(n, policy="mixed", device="cuda:0")
    L = eng.lib
    th = nl_engine.theta_reference()
    eng.set_reference(th)
    w = nl_c.init_weights(n, 8)
    noise = torch.empty((steps, n), dtype=torch.float32, device="cuda")
    if noise_on_host:
        noise.copy_(torch.as_tensor(np.random.default_rng(2).standard_normal((steps, n)).astype(np.float32)))
    else:
        _lib.check(L.rl4_nl_noise_fill(seed, 0, 0, steps, n, noise.data_ptr(), n, None), "noise")
    eng.init(w["W1a"], w["W2a"], w["W1c"], w["W2c"])
    eng.run(steps, noise)
    pin = lambda a: torch.as_tensor(np.ascontiguousarray(a.T)).pin_memory()     # noqa: E731
    h = [pin(w[k]) for k in ("W1a", "W2a", "W1c", "W2c")]
    href = torch.as_tensor(th[:steps].copy()).pin_memory()
    hnoise = noise.cpu().pin_memory() if noise_on_host else None
    oenv = torch.full((NLE["COUNT"], n), -7.0, dtype=torch.float64).pin_memory()
    onet = torch.full((NLN["COUNT"], n), -7.0, dtype=torch.float32).pin_memory()
    oint = torch.full((NLI["COUNT"], n), -7, dtype=torch.int32).pin_memory()
    ctx = ctypes.c_void_p()
    _lib.check(L.rl4_ctx_create(0, _lib.MIXED, n, steps, ctypes.byref(ctx)), "ctx")
    eq = lambda u, v: bool(((u == v) | (torch.isnan(u) & torch.isnan(v))).all())   # noqa: E731
    try:
        io = _lib.NlHostIO(h[0].data_ptr(), h[1].data_ptr(), h[2].data_ptr(), h[3].data_ptr(), href.data_ptr(),
                           hnoise.data_ptr() if noise_on_host else None, seed, 0, oenv.data_ptr(), onet.data_ptr(), oint.data_ptr(), 0, 0)
        _lib.check(L.rl4_nl_episode_host(ctx, ctypes.byref(eng.params), ctypes.byref(io), n, steps), "nl episode")
        assert eq(oenv, eng.env[:, :n].cpu()) and eq(onet, eng.net[:, :n].cpu()) and torch.equal(oint, eng.ints[:, :n].cpu())
        # statistics only: RSE / n_z / learning rates + the int plane
        oenv.fill_(-7.0); onet.fill_(-7.0); oint.fill_(-7)
        io.out_mask = OUT["STATS"]
        _lib.check(L.rl4_nl_episode_host(ctx, ctypes.byref(eng.params), ctypes.byref(io), n, steps), "nl episode (stats)")
        rows = np.zeros(NLE["COUNT"], bool); rows[NLE["RSE"]:NLE["EA"]] = True; rows[NLE["RSE_FLIGHT"]:] = True
        r = torch.as_tensor(rows)
        assert eq(oenv[r], eng.env[:, :n].cpu()[r]) and bool((oenv[~r] == -7.0).all()) and bool((onet == -7.0).all())
        assert torch.equal(oint, eng.ints[:, :n].cpu())
    finally:
        L.rl4_ctx_destroy(ctx)
    assert int((eng.int_field("DIVERGED_STEP") >= 0).sum()) < n // 2
