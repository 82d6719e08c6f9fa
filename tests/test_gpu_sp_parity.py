"""GPU parity of the fused short-period kernel against the CPU oracle (C restatement of
envs/linear/env.py:156-220 + objects.py:142-281,439-549,551-1004), through the C-ABI.

Bar: with the t13 tanh on both sides every state variable is BIT-IDENTICAL (all three dtype
policies, all trace modes, faults, per-agent hyper-parameters, divergence).  With the oracle on
libm tanh (the reference's primitive is np.tanh / tf.tanh, itself host-dependent) the fp64 path
must stay within 1e-10 relative per step (teacher-forced) -- BASELINE.json's tolerance.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from tests import _util  # noqa: E402


def _make(oracle, n, policy, *, seed=0, fault=None, elig=(None, None), ms=2, x0_scale=2.0, per_agent=False,
          eta_a_h=None, kappa=None):
    from rl4afcs_b200 import sp_engine

    ic = oracle.default_idhp_config()
    ic["multistep"] = ms
    ic["actor_config"]["elig"], ic["critic_config"]["elig"] = elig
    if eta_a_h is not None:
        ic["actor_config"]["eta_h"] = eta_a_h
    if kappa is not None:
        ic["kappa"] = kappa
    rng = np.random.default_rng(seed)
    x0 = np.deg2rad(rng.uniform(-x0_scale, x0_scale, size=(n, 2)))
    w = oracle.init_weights(n, seed + 1)
    base, amp = oracle.default_reference()
    cfg = oracle.make_cfg(ic, fault_scenario=fault, n=n if per_agent else 1)
    eng = sp_engine.SpEngine(n, policy=policy, device="cuda:0")
    sp_engine.apply_idhp_config(eng, ic, dt=0.02)
    eng.set_hp("REF_AMP", amp)
    eng.set_hpi("FAULT_STEP", -1 if fault is None else int(20 / 0.02))
    eng.set_hpi("FAULT_KIND", 0 if fault is None else {"invert_elevator": 1, "damp_elevator": 2, "shift_cg": 3}[fault])
    if per_agent:
        # per-agent sweep: actor/critic lr, rls forgetting factor, reference amplitude, fault time, trace modes
        ea = rng.uniform(2.5, 4.7, n); ec = rng.uniform(0.2, 0.55, n)
        rg = rng.uniform(0.99, 1.0, n); am = np.deg2rad(rng.uniform(1, 10, n))
        fs = rng.integers(200, 900, n).astype(np.int32); fk = rng.integers(0, 4, n).astype(np.int32)
        ela = rng.integers(0, 3, n).astype(np.int32); elc = rng.integers(0, 3, n).astype(np.int32)
        msv = rng.integers(0, 2, n).astype(np.int32)
        cfg["eta_a_h"], cfg["eta_c_h"], cfg["rls_gamma"], cfg["ref_amp"] = ea, ec, rg, am
        cfg["fault_step"], cfg["elig_a"], cfg["elig_c"], cfg["multistep"] = fs, ela, elc, msv
        from oracle import sp_c
        k = sp_c.ce500_coeffs(); A, B = sp_c.ce500_A(k), sp_c.ce500_B(k)
        names = [None, "invert_elevator", "damp_elevator", "shift_cg"]
        for i in range(n):
            Af, Bf = sp_c.ce500_fault(k, A, B, names[fk[i]])
            cfg["A_fault"][i] = Af.reshape(-1); cfg["B_fault"][i] = Bf.reshape(-1)
        eng.set_hp("ETA_A_H", ea); eng.set_hp("ETA_C_H", ec); eng.set_hp("RLS_GAMMA", rg); eng.set_hp("REF_AMP", am)
        eng.set_hpi("FAULT_STEP", fs); eng.set_hpi("FAULT_KIND", fk)
        eng.set_hpi("ELIG_A", ela); eng.set_hpi("ELIG_C", elc); eng.set_hpi("MULTISTEP", msv)
    eng.set_reference(base)
    eng.init(x0, w["W1a"], w["W2a"], w["W1c"], w["W2c"])
    st = oracle.init_states(policy, cfg, x0, w)
    return eng, st, cfg, base, ic


def _gl_of(cfg, ic):
    def f(low):
        lh = cfg["lambda_h"] * cfg["gamma"]
        ll = cfg["lambda_l"] * cfg["gamma"]
        return np.where(low, ll, lh)
    return f


@pytest.mark.parametrize("policy", ["fp64", "mixed", "fp32"])
def test_init_matches_oracle(oracle, policy):
    eng, st, cfg, base, ic = _make(oracle, 257, policy)
    got = _util.engine_state_to_oracle(eng, oracle, _gl_of(cfg, ic))
    assert _util.state_mismatches(got, st) == {}


@pytest.mark.parametrize("policy", ["fp64", "mixed", "fp32"])
@pytest.mark.parametrize("case", [
    dict(),                                                  # idhp_sp.py defaults (multistep, no traces)
    dict(ms=0),
    dict(fault="shift_cg"),
    dict(fault="invert_elevator", elig=("accumulating", "accumulating"), ms=0),
    dict(fault="damp_elevator", elig=("replacing", "replacing")),
    dict(elig=("accumulating", None)),
])
def test_full_episode_bit_exact_t13(oracle, policy, case):
    n, steps = 384, 3000
    eng, st, cfg, base, ic = _make(oracle, n, policy, seed=11, **case)
    oracle.run(policy, cfg, base, st, 0, steps, tanh="t13")
    eng.run(steps)
    got = _util.engine_state_to_oracle(eng, oracle, _gl_of(cfg, ic))
    assert _util.state_mismatches(got, st) == {}


@pytest.mark.parametrize("policy", ["fp64", "mixed", "fp32"])
def test_per_agent_hparams_and_chunked_resume_bit_exact(oracle, policy):
    n, steps = 512, 1500
    eng, st, cfg, base, ic = _make(oracle, n, policy, seed=5, per_agent=True)
    oracle.run(policy, cfg, base, st, 0, steps, tanh="t13")
    for chunk in (1, 1, 1, 97, 400, 1000):                   # resume must not change anything
        eng.run(chunk)
    got = _util.engine_state_to_oracle(eng, oracle, _gl_of(cfg, ic))
    assert _util.state_mismatches(got, st) == {}


@pytest.mark.parametrize("policy", ["fp64", "mixed", "fp32"])
def test_divergence_is_reproduced(oracle, policy):
    # critic learning rates from sane to absurd: part of the batch blows up to NaN.  The NaN
    # convention (objects.py:991, Q19) must freeze the same agents at the same step.
    n, steps = 512, 1200
    eng, st, cfg, base, ic = _make(oracle, n, policy, seed=3, x0_scale=20.0)
    eta_c = np.random.default_rng(9).uniform(0.3, 4.0, n)
    cfg = np.repeat(cfg, n)
    cfg["eta_c_h"] = eta_c
    eng.set_hp("ETA_C_H", eta_c)
    rng = np.random.default_rng(3)
    x0 = np.deg2rad(rng.uniform(-20.0, 20.0, size=(n, 2)))
    w = oracle.init_weights(n, 4)
    eng.init(x0, w["W1a"], w["W2a"], w["W1c"], w["W2c"])
    st = oracle.init_states(policy, cfg, x0, w)
    oracle.run(policy, cfg, base, st, 0, steps, tanh="t13")
    eng.run(steps)
    got = _util.engine_state_to_oracle(eng, oracle, _gl_of(cfg, ic))
    nd = int((st["diverged_step"] >= 0).sum())
    assert 0 < nd < n, f"want a mix of diverged and healthy agents, got {nd}/{n}"
    assert np.array_equal(got["diverged_step"], st["diverged_step"])
    assert np.array_equal(got["x_nan"], st["x_nan"])
    ok = st["diverged_step"] < 0
    assert _util.state_mismatches(got[ok], st[ok]) == {}


def test_log_matches_oracle_log(oracle):
    n, steps, nlog = 64, 600, 16
    policy = "mixed"
    eng, st, cfg, base, ic = _make(oracle, n, policy, seed=21, fault="shift_cg", elig=("accumulating", "accumulating"))
    olog = oracle.run(policy, cfg, base, st, 0, steps, tanh="t13", n_log=nlog)
    from rl4afcs_b200 import _lib
    lg = eng.run(steps, log_level=_lib.LOG_FULL, log_agents=nlog).cpu().numpy()     # (rows, fields, agents)
    lg = np.transpose(lg, (2, 0, 1))                                                  # (agents, rows, fields)
    LF, LB = _lib.LF, _lib.LB
    def eq(a, b):
        return np.array_equal(a, b, equal_nan=True)
    assert eq(lg[:, :, LB["X"]:LB["X"] + 2], olog["x"])
    assert eq(lg[:, :, LB["A"]], olog["a"])
    assert eq(lg[:, :, LB["C"]], olog["c"])
    assert eq(lg[:, :, LB["REF"]], olog["ref"])
    assert eq(lg[:, :, LB["E"]], olog["e"])
    assert eq(lg[:, :, LF["AW1"]:LF["AW1"] + 4], olog["a_w1"])
    assert eq(lg[:, :, LF["AW2"]:LF["AW2"] + 4], olog["a_w2"])
    assert eq(lg[:, :, LF["CW1"]:LF["CW1"] + 4], olog["c_w1"])
    assert eq(lg[:, :, LF["CW2"]:LF["CW2"] + 8], olog["c_w2"])
    assert eq(lg[:, :, LF["AE"]:LF["AE"] + 8], olog["a_e"])
    assert eq(lg[:, :, LF["CE"]:LF["CE"] + 4], olog["c_e"][:, :, 0:4])
    assert eq(lg[:, :, LF["CE"] + 4:LF["CE"] + 8], olog["c_e"][:, :, 8:12])
    assert eq(lg[:, :, LF["CE"] + 8:LF["CE"] + 12], olog["c_e"][:, :, 20:24])
    assert eq(lg[:, :, LF["AGRAD"]:LF["AGRAD"] + 8], olog["a_all_grad"])
    assert eq(lg[:, :, LF["CGRAD"]:LF["CGRAD"] + 12], olog["c_all_grad"])
    assert eq(lg[:, :, LF["PARAMS"]:LF["PARAMS"] + 6], olog["params"])
    assert eq(lg[:, :, LF["COV"]:LF["COV"] + 9], olog["cov"])
    assert eq(lg[:, :, LF["EPS_NORM"]], olog["eps_norm"])
    assert eq(lg[:, :, LF["LAM"]:LF["LAM"] + 2], olog["lam"])
    assert eq(lg[:, :, LF["TD"]:LF["TD"] + 2], olog["td"])
    assert eq(lg[:, :, LF["DADZ"]], olog["dadz"])
    assert eq(lg[:, :, LF["M"]:LF["M"] + 4], olog["M"])
    assert eq(lg[:, :, LF["LOSS_GRAD"]], olog["loss_grad"])


def _rel(a, b, scale):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), scale))


def test_fp64_teacher_forced_vs_libm_tanh_oracle_1e10(oracle):
    """BASELINE.json: fp64 path within 1e-10 relative per step over the first 1000 steps.
    The oracle runs libm tanh (the reference's primitive is not bit-reproducible); the kernel is
    re-seeded from the oracle state every step, so this measures the per-step deviation."""
    n, steps = 64, 1000
    policy = "fp64"
    eng, st, cfg, base, ic = _make(oracle, n, policy, seed=31)
    gl_low = float(cfg["lambda_l"][0] * cfg["gamma"][0])
    worst = 0.0
    for k in range(steps):
        _util.oracle_state_to_engine(st, eng, gl_low=gl_low)
        eng.k = k
        oracle.run(policy, cfg, base, st, k, 1, tanh="libm")
        eng.run(1)
        got = _util.engine_state_to_oracle(eng, oracle, _gl_of(cfg, ic))
        for f, scale in (("x", 1e-3), ("a", 1e-3), ("W1a", 1e-3), ("W2a", 1e-3), ("W1c", 1e-3), ("W2c", 1e-3),
                         ("W1t", 1e-3), ("W2t", 1e-3), ("theta", 1e-3), ("M_prev", 1e-3)):
            worst = max(worst, _rel(got[f], st[f], scale))
        # covariance: relative to its own magnitude (entries span 1e-3 .. 1e6)
        worst = max(worst, _rel(got["cov"], st["cov"], np.abs(st["cov"]).max(axis=1, keepdims=True) * 1e-6))
        assert np.array_equal(got["cooldown"], st["cooldown"])
    assert worst < 1e-10, worst


def test_fp32_single_step_within_1e4_of_fp64_oracle(oracle):
    """BASELINE.json: fp32 path within 1e-4 relative on per-step state and weight updates from
    identical inputs (the fp64 oracle with libm tanh is the yardstick)."""
    n, steps = 64, 300
    eng, st32, cfg, base, ic = _make(oracle, n, "fp32", seed=41)
    _, st64, _, _, _ = _make(oracle, n, "fp64", seed=41)
    gl_low = float(cfg["lambda_l"][0] * cfg["gamma"][0])
    worst = 0.0
    for k in range(steps):
        # identical inputs: the fp32-representable state of the fp32 oracle trajectory
        st64 = st32.copy()
        _util.oracle_state_to_engine(st32, eng, gl_low=gl_low)
        eng.k = k
        oracle.run("fp64", cfg, base, st64, k, 1, tanh="libm")
        oracle.run("fp32", cfg, base, st32, k, 1, tanh="t13")
        eng.run(1)
        got = _util.engine_state_to_oracle(eng, oracle, _gl_of(cfg, ic))
        assert _util.state_mismatches(got, st32, fields=("x", "a", "W1a", "W2a", "W1c", "W2c", "theta")) == {}
        for f, scale in (("x", 1e-2), ("a", 1e-2), ("W1a", 1e-2), ("W2a", 1e-2), ("W1c", 1e-2), ("W2c", 1e-2)):
            worst = max(worst, _rel(got[f], st64[f], scale))
    assert worst < 1e-4, worst


@pytest.mark.parametrize("policy", ["fp64", "mixed", "fp32"])
def test_full_size_batch_by_replication_property(oracle, policy):
    """BASELINE.json configs[1] size (2^20 agents): agents are independent, so a batch that replicates a 1024-agent
    block 1024 times must reproduce that block bit for bit at every position (no cross-agent leakage, no
    position-dependent arithmetic), and the block itself equals the oracle."""
    from rl4afcs_b200 import sp_engine

    n_small, reps, steps = 1024, 1024, 400
    n = n_small * reps
    ic = oracle.default_idhp_config()
    base, amp = oracle.default_reference()
    rng = np.random.default_rng(77)
    x0 = np.deg2rad(rng.uniform(-2, 2, size=(n_small, 2)))
    w = oracle.init_weights(n_small, 78)
    cfg = oracle.make_cfg(ic)
    st = oracle.init_states(policy, cfg, x0, w)
    oracle.run(policy, cfg, base, st, 0, steps, tanh="t13")
    eng = sp_engine.SpEngine(n, policy=policy)
    sp_engine.apply_idhp_config(eng, ic, dt=0.02)
    eng.set_hp("REF_AMP", amp); eng.set_hpi("FAULT_STEP", -1); eng.set_hpi("FAULT_KIND", 0)
    eng.set_reference(base)
    tile = lambda a: torch.as_tensor(a).cuda().repeat(reps, 1)    # noqa: E731
    eng.init(tile(x0), tile(w["W1a"]), tile(w["W2a"]), tile(w["W1c"]), tile(w["W2c"]))
    eng.run(steps)
    for plane in (eng.env, eng.net, eng.ints):
        v = plane.reshape(plane.shape[0], reps, n_small)
        same = (v == v[:, :1, :]) | (torch.isnan(v) & torch.isnan(v[:, :1, :])) if plane.is_floating_point() else (v == v[:, :1, :])
        assert bool(same.all())
    small = sp_engine.SpEngine(n_small, policy=policy)
    small.env.copy_(eng.env[:, :n_small]); small.net.copy_(eng.net[:, :n_small]); small.ints.copy_(eng.ints[:, :n_small])
    got = _util.engine_state_to_oracle(small, oracle, _gl_of(cfg, ic))
    assert _util.state_mismatches(got, st) == {}
