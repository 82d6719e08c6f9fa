"""GPU parity of the nonlinear path (Ce500NonLinear wrapper + surrogate plant + IDHPnonlin) against the CPU
oracle (oracle/nl_oracle.c: restatement of envs/nonlinear/env.py:60-311 and objects.py:283-437,1006-1564).

The surrogate plant is written with IEEE basic operations only (polynomial sin / cos, binomial-series atmosphere,
explicit FMA chains), the networks use the t13 tanh: kernel == oracle BIT FOR BIT over free-running episodes
(test_free_run_bit_exact).  The teacher-forced tolerance tests below predate that and stay as a second, looser net.
The surrogate is calibrated against the reference's plant binary, not a restatement of it; the reference's OWN model on the GPU
(plant='dasmat', translated from the binary) has its tests in tests/test_gpu_dasmat.py."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from tests import _util_nl  # noqa: E402


@pytest.fixture(scope="module")
def nl(oracle):
    from oracle import nl_c
    nl_c.lib()
    return nl_c


def _setup(nl, n, policy, *, seed=0, fault=None, fault_time=60, integrator="ode5", elig="accumulating", ms=0, **over):
    from rl4afcs_b200 import _lib, nl_engine

    cfg = nl.make_cfg(fault=fault, fault_time=fault_time, integrator=integrator, elig_a=elig, multistep=ms, **over)
    w = nl.init_weights(n, seed)
    st = nl.init_states(policy, cfg, w, n)
    eng = nl_engine.NlEngine(n, policy=policy)
    damp, sat = nl_engine.split_fault(fault)
    eng.set_hpi("FAULT_DAMP", damp); eng.set_hpi("FAULT_SAT", sat)
    eng.set_hpi("FAULT_STEP", int(cfg["fault_step"][0]))
    eng.set_hpi("ELIG_A", _lib.ELIG[elig]); eng.set_hpi("MULTISTEP", 1 if ms else 0)
    eng.params.integrator = _lib.INTEGRATOR[integrator]
    for k, v in over.items():
        eng.set_hp(k.upper(), v)
    th = nl.theta_reference()
    eng.set_reference(th)
    eng.init(w["W1a"], w["W2a"], w["W1c"], w["W2c"])
    return eng, st, cfg, th


@pytest.mark.parametrize("policy", ["mixed", "fp64"])
def test_reset_and_prologue(nl, policy):
    eng, st, cfg, th = _setup(nl, 70, policy)
    got = _util_nl.engine_to_oracle(eng, nl)
    assert np.array_equal(got["x_full"], st["x_full"])                          # 1001 trim steps of the plant, bit for bit
    for f in ("W1a", "W2a", "W1c", "W2c", "W1t", "W2t", "cov", "eta_a", "eta_c", "lambdaa", "lr_a", "lr_c", "gl"):
        assert np.array_equal(got[f], st[f]), f
    assert np.array_equal(got["diverged_step"], st["diverged_step"])


@pytest.mark.parametrize("policy,case", [
    ("mixed", dict()),                                                           # idhp_nonlin.py defaults
    ("fp64", dict()),
    ("mixed", dict(integrator="rk4", elig=None, ms=1)),
    ("mixed", dict(fault="damp_elevator_and_saturate_elevator", fault_time=2.0, elig="replacing")),
    ("fp64", dict(fault="shift_cg", fault_time=1.5, ms=1)),
    ("mixed", dict(fault="slow_all", fault_time=1.0)),
])
def test_teacher_forced_steps(nl, policy, case):
    n, steps = 48, 450
    eng, st, cfg, th = _setup(nl, n, policy, seed=3, **case)
    rng = np.random.default_rng(1)
    noise = rng.standard_normal((steps, n)).astype(np.float32)
    worst = {}
    f32 = policy == "mixed"
    for k in range(steps):
        _util_nl.oracle_to_engine(st, eng)
        eng.k = k
        nl.run(policy, cfg, th, noise[k:k + 1], st, k, 1, tanh="t13")
        eng.run(1, noise[k:k + 1])
        got = _util_nl.engine_to_oracle(eng, nl)
        for f, floor in (("x_full", 1e-3), ("x_act", 1e-4), ("x_lon", 1e-3), ("theta", 1e-4), ("eps", 1e-6), ("rse", 1e-3),
                         ("eta_a", 1e-3), ("eta_c", 1e-3), ("lambdaa", 1e-3), ("gl", 1e-3)):
            worst[f] = max(worst.get(f, 0.0), _util_nl.max_rel(got[f], st[f], floor))
        worst["cov"] = max(worst.get("cov", 0.0), _util_nl.max_rel(got["cov"], st["cov"], np.abs(st["cov"]).max(axis=1, keepdims=True) * 1e-6))
        for f, floor in (("s", 1e-3), ("a", 1e-3), ("W1a", 1e-2), ("W2a", 1e-2), ("W1c", 1e-2), ("W2c", 1e-2), ("W1t", 1e-2),
                         ("W2t", 1e-2), ("M_prev", 1e-2), ("Ea", 1e-2), ("lr_a", 1e-3), ("lr_c", 1e-3)):
            worst[f] = max(worst.get(f, 0.0), _util_nl.max_rel(got[f], st[f], floor))
        assert np.array_equal(got["cooldown"], st["cooldown"]) and np.array_equal(got["stepp"], st["stepp"]), k
        assert np.array_equal(got["diverged_step"], st["diverged_step"]), k
    tol64 = 1e-9
    tol_net = 5e-6 if f32 else 1e-9       # a last-bit plant difference can flip one float32 rounding of s
    for f in ("x_full", "x_act", "x_lon", "rse", "eta_a", "eta_c", "lambdaa", "gl"):
        assert worst[f] < tol64, (f, worst[f])
    for f in ("theta", "cov", "eps"):
        assert worst[f] < (1e-4 if f32 else 1e-7), (f, worst[f])
    for f in ("s", "a", "W1a", "W2a", "W1c", "W2c", "W1t", "W2t", "M_prev", "Ea", "lr_a", "lr_c"):
        assert worst[f] < tol_net, (f, worst[f])


@pytest.mark.parametrize("policy", ["mixed", "fp64"])
def test_free_run_tracks_oracle_and_chunking_is_exact(nl, policy):
    n, steps = 64, 600
    eng, st, cfg, th = _setup(nl, n, policy, seed=5)
    rng = np.random.default_rng(2)
    noise = rng.standard_normal((steps, n)).astype(np.float32)
    olog = nl.run(policy, cfg, th, noise, st, 0, steps, tanh="t13", n_log=n)
    lg = eng.run(steps, noise, log_agents=n).cpu().numpy()          # (rows, fields, agents)
    from rl4afcs_b200 import _lib
    x_gpu = np.transpose(lg[:, _lib.NLL["XFULL"]:_lib.NLL["XFULL"] + 12, :], (2, 0, 1))
    # early part of the run: still tightly together
    assert _util_nl.max_rel(x_gpu[:, :200, :9], olog["x_full"][:, :200, :9], 1e-2) < 1e-6
    got = _util_nl.engine_to_oracle(eng, nl)
    assert np.array_equal(got["diverged_step"] >= 0, st["diverged_step"] >= 0)
    # the same run in chunks gives the same bits (resume exactness on the device)
    eng2, _, _, _ = _setup(nl, n, policy, seed=5)
    k = 0
    for chunk in (1, 7, 92, 500):
        eng2.run(chunk, noise[k:k + chunk]); k += chunk
    same = lambda u, v: bool(((u == v) | (torch.isnan(u) & torch.isnan(v))).all())   # noqa: E731
    assert same(eng.env, eng2.env) and same(eng.net, eng2.net) and torch.equal(eng.ints, eng2.ints)


def test_agents_learn_to_track_on_the_surrogate(nl):
    """Sanity of the whole loop on the GPU: with the reference's hyper-parameters (idhp_nonlin.py:123-146) most
    agents survive the 90 s episode on the surrogate plant and end with a small pitch-tracking error."""
    n, steps = 256, 9000
    eng, st, cfg, th = _setup(nl, n, "mixed", seed=8)
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    e_last = None
    for k0 in range(0, steps, 1500):
        noise = torch.randn((1500, n), generator=g, device="cuda", dtype=torch.float32)
        lg = eng.run(1500, noise, log_agents=n)
        e_last = lg[:, eng_field("E_THETA"), :]
    alive = ~eng.stats()["diverged"]
    assert alive.float().mean() > 0.8
    rms = torch.sqrt((e_last[-1000:] ** 2).mean(dim=0))[alive]
    assert torch.rad2deg(rms.median()) < 3.0


def eng_field(name):
    from rl4afcs_b200 import _lib
    return _lib.NLL[name]


@pytest.mark.parametrize("policy", ["mixed", "fp64"])
def test_step_level_actor_big_critic_big(nl, policy):
    """Critic_big.__call__, Actor_big.__call__ (+ trace, da/ds), get_weight_update, soft_update vs the oracle's
    per-step log from identical inputs (objects.py:294-339, 374-427)."""
    from rl4afcs_b200.objects import Actor_big, Critic_big
    from rl4afcs_b200 import nl_engine

    n, steps = 40, 30
    eng, st, cfg, th = _setup(nl, n, policy, seed=9)
    rng = np.random.default_rng(4)
    noise = rng.standard_normal((steps, n)).astype(np.float32)
    actor = Actor_big(eng, "W1A", "W2A", 1, "accumulating")
    critic = Critic_big(eng, "W1C", "W2C", 3, 1233)
    target = Critic_big(eng, "W1T", "W2T", 3, 1233)
    for k in range(steps):
        pre = st.copy()
        lg = nl.run(policy, cfg, th, noise[k:k + 1], st, k, 1, tanh="t13", n_log=n)[:, 0]
        _util_nl.oracle_to_engine(pre, eng)                      # objects now see the pre-step weights / trace
        lam = critic(pre["s_prev"])
        lam_t = target(lg["s_next"])
        a, dads = actor(pre["s_prev"], trace=True, return_input_gradient=True)
        assert np.array_equal(lam.double().cpu().numpy().reshape(n, 3), lg["lam"]), k
        assert np.array_equal(lam_t.double().cpu().numpy().reshape(n, 3), lg["lam_t"]), k
        assert np.array_equal(a.double().cpu().numpy().ravel(), lg["a_next"]), k
        assert np.array_equal(dads.double().cpu().numpy(), lg["dads"]), k
        if k > 0:
            # second trace pass at s_random (objects.py:1375-1378), then loss * E
            s_rand = (torch.as_tensor(noise[k]).cuda().to(eng.tn)[:, None] * torch.as_tensor(cfg["noise_std"][0]).cuda().to(eng.tn)[None, :]
                      + torch.as_tensor(pre["s_prev"]).cuda().to(eng.tn))
            a_r = actor(s_rand, trace=True)
            assert np.array_equal(a_r.double().cpu().numpy().ravel(), lg["a_random"]), k
            W1u, W2u = actor.get_weight_update(torch.as_tensor(lg["loss_grad"]).reshape(n, 1, 1))
            lr = torch.as_tensor(pre["lr_a"]).cuda().to(eng.tn)
            W1, W2 = actor.trainable_weights
            W1.copy_(W1 - lr[:, None, None] * W1u); W2.copy_(W2 - lr[:, None, None] * W2u)
            assert np.array_equal(W1.double().cpu().numpy().reshape(n, 40), st["W1a"]), k
            assert np.array_equal(W2.double().cpu().numpy().reshape(n, 10), st["W2a"]), k
            assert np.array_equal(actor.E.cpu().numpy().reshape(n, 50), st["Ea"]), k


_FULL_LOG_MAP = (  # (kernel field, oracle log-row field, relative-error floor, network-side quantity?)
    ("ETA_A", "eta_a", 1e-3, False), ("XFULL", "x_full", 1e-3, False), ("RSE", "rse_step", 1e-3, False),
    ("A_CMD", None, 1e-4, False), ("A_EFF", None, 1e-4, False), ("S", None, 1e-3, False), ("YREF", "yref_theta", 1e-3, False),
    ("E", "e_theta", 1e-4, False), ("A_W1", "W1a", 1e-2, True), ("A_W2", "W2a", 1e-2, True), ("C_W1", "W1c", 1e-2, True),
    ("C_W2", "W2c", 1e-2, True), ("A_GRAD", "a_grad", 1e-3, True), ("C_GRAD", "c_grad", 1e-3, True),
    ("RLS_PARAMS", "theta", 1e-4, None), ("RLS_EPS", "eps", 1e-6, None), ("RLS_EPS_NORM", "eps_norm", 1e-6, None),
    ("A", "a_next", 1e-3, True), ("REWARD", "reward", 1e-4, False))


@pytest.mark.parametrize("policy,case", [("mixed", dict()), ("fp64", dict(ms=1, elig=None)),
                                         ("mixed", dict(fault="damp_elevator_and_saturate_elevator", fault_time=0.5))])
def test_full_log_rows_teacher_forced(nl, policy, case):
    """log.level 2 (the layout of IDHPnonlin._log, objects.py:1083-1176) against the oracle's per-step record."""
    from rl4afcs_b200 import _lib

    n, steps = 24, 160
    eng, st, cfg, th = _setup(nl, n, policy, seed=11, **case)
    rng = np.random.default_rng(4)
    noise = rng.standard_normal((steps, n)).astype(np.float32)
    f32 = policy == "mixed"
    worst = {}
    for k in range(steps):
        _util_nl.oracle_to_engine(st, eng)
        eng.k = k
        olog = nl.run(policy, cfg, th, noise[k:k + 1], st, k, 1, tanh="t13", n_log=n)[:, 0]
        row = eng.run(1, noise[k:k + 1], log_agents=n, log_level=2)[0].cpu().numpy()      # (fields, agents)
        assert row.shape == (_lib.NLF["COUNT"], n)
        for name, oname, floor, _ in _FULL_LOG_MAP:
            off, w = _lib.NLF_FIELDS[name]
            got = row[off:off + w].T
            if name == "A_CMD":
                want = olog["surf"][:, :1]
            elif name == "A_EFF":
                want = olog["model_input"][:, :1]
            elif name == "S":
                want = olog["x_full"][:, 4:5]
            else:
                want = olog[oname].reshape(n, w)
            worst[name] = max(worst.get(name, 0.0), _util_nl.max_rel(got, want, floor))
        off, w = _lib.NLF_FIELDS["X"]
        assert np.array_equal(row[off:off + w], row[[1 + 4, 1 + 7, 1 + 1]])            # x_lon = x_full[4, 7, 1]
        off, w = _lib.NLF_FIELDS["RLS_COV"]
        worst["RLS_COV"] = max(worst.get("RLS_COV", 0.0), _util_nl.max_rel(
            row[off:off + w].T, olog["cov"], np.abs(olog["cov"]).max(axis=1, keepdims=True) * 1e-6))
        if k == 0:
            off, w = _lib.NLF_FIELDS["A_GRAD"]
            assert not row[off:off + 120].any()                                       # objects.py:1149: untouched at i = 0
    for name, _, _, net in _FULL_LOG_MAP:
        tol = (5e-6 if f32 else 1e-9) if net else ((1e-4 if f32 else 1e-7) if net is None else 1e-9)
        assert worst[name] < tol, (name, worst[name])
    assert worst["RLS_COV"] < (1e-4 if f32 else 1e-7)


def test_full_log_nan_rows_after_divergence(nl):
    """objects.py:1168-1175: the row of the step whose plant state turned NaN, and every later row, are NaN."""
    from rl4afcs_b200 import _lib

    n, steps = 8, 12
    eng, st, cfg, th = _setup(nl, n, "mixed", seed=2)
    noise = np.zeros((steps, n), dtype=np.float32)
    eng.run(4, noise[:4])
    eng.env_field("XFULL", 12)[3, 5] = float("nan")                                  # agent 5: airspeed becomes NaN
    lg = eng.run(8, noise[4:], log_agents=n, log_level=2).cpu().numpy()              # (rows, fields, agents)
    assert np.isnan(lg[:, :, 5]).all()
    assert not np.isnan(lg[:, :, [0, 1, 2, 3, 4, 6, 7]]).any()
    assert int(eng.int_field("DIVERGED_STEP")[5]) == 4
    lg1 = eng.run(2, np.zeros((2, n), dtype=np.float32), log_agents=n, log_level=1).cpu().numpy()
    assert lg1.shape[1] == _lib.NLL["COUNT"] and np.isnan(lg1[:, :, 5]).all()


def test_full_size_batch_by_replication_property(nl):
    """BASELINE.json configs[2] size (262144 agents): a batch that replicates a 256-agent block 1024 times must give
    every replica the same bits (agents independent, no position-dependent arithmetic), per-agent fault settings
    included, and the block itself stays with the oracle over the early part of the run."""
    n_small, reps, steps = 256, 1024, 240
    n = n_small * reps
    from rl4afcs_b200 import _lib, nl_engine

    cfg = nl.make_cfg()
    w = nl.init_weights(n_small, 31)
    st = nl.init_states("mixed", cfg, w, n_small)
    th = nl.theta_reference()
    rng = np.random.default_rng(9)
    noise = rng.standard_normal((steps, n_small)).astype(np.float32)
    olog = nl.run("mixed", cfg, th, noise, st, 0, steps, tanh="t13", n_log=n_small)
    eng = nl_engine.NlEngine(n, policy="mixed")
    eng.set_reference(th)
    tile = lambda a: torch.as_tensor(a).cuda().repeat(reps, 1)    # noqa: E731
    eng.init(tile(w["W1a"]), tile(w["W2a"]), tile(w["W1c"]), tile(w["W2c"]))
    lg = eng.run(steps, torch.as_tensor(noise).cuda().repeat(1, reps), log_agents=n_small)
    for plane in (eng.env, eng.net, eng.ints):
        v = plane[:, :n].reshape(plane.shape[0], reps, n_small)
        same = (v == v[:, :1, :]) | (torch.isnan(v) & torch.isnan(v[:, :1, :])) if plane.is_floating_point() else (v == v[:, :1, :])
        assert bool(same.all())
    x_gpu = np.transpose(lg.cpu().numpy()[:, _lib.NLL["XFULL"]:_lib.NLL["XFULL"] + 12, :], (2, 0, 1))
    assert _util_nl.max_rel(x_gpu[:, :200, :9], olog["x_full"][:, :200, :9], 1e-2) < 1e-6


_EXACT_FIELDS = ("x_full", "x_act", "x_lon", "x_prev_lon", "theta", "cov", "eps", "eps_norm", "rse", "rse_flight", "nz_peak",
                 "eta_a", "eta_c", "lambdaa", "gl", "Ea", "s", "s_prev", "a", "a_prev", "W1a", "W2a", "W1c", "W2c", "W1t", "W2t",
                 "M_prev", "lr_a", "lr_c", "cooldown", "diverged_step", "stepp")


@pytest.mark.parametrize("policy,case", [
    ("mixed", dict()),
    ("fp64", dict()),
    ("mixed", dict(integrator="rk4", elig="replacing", ms=1)),
    ("fp64", dict(fault="shift_cg", fault_time=1.5, ms=1, elig=None)),
    ("mixed", dict(fault="damp_elevator_and_saturate_elevator", fault_time=2.0)),
])
def test_free_run_bit_exact(nl, policy, case):
    """The surrogate plant uses IEEE basic operations only (polynomial sin / cos, binomial-series atmosphere), the
    networks the t13 tanh: a FREE-RUNNING episode on the GPU equals the oracle bit for bit, state and full log."""
    from rl4afcs_b200 import _lib

    n, steps = 96, 1500
    eng, st, cfg, th = _setup(nl, n, policy, seed=21, **case)
    rng = np.random.default_rng(5)
    noise = rng.standard_normal((steps, n)).astype(np.float32)
    olog = nl.run(policy, cfg, th, noise, st, 0, steps, tanh="t13", n_log=n)
    lg = eng.run(steps, noise, log_agents=n, log_level=2).cpu().numpy()          # (rows, fields, agents)
    got = _util_nl.engine_to_oracle(eng, nl)
    for f in _EXACT_FIELDS:
        assert np.array_equal(got[f], st[f], equal_nan=(got[f].dtype.kind == "f")), f
    assert np.array_equal(got["cgrad_prev"][:, 2], st["cgrad_prev"][:, 2], equal_nan=True)
    for name, oname in (("XFULL", "x_full"), ("A_W1", "W1a"), ("C_W2", "W2c"), ("A_GRAD", "a_grad"), ("C_GRAD", "c_grad"),
                        ("RLS_PARAMS", "theta"), ("RLS_COV", "cov"), ("RSE", "rse_step"), ("ETA_A", "eta_a")):
        off, w = _lib.NLF_FIELDS[name]
        assert np.array_equal(np.transpose(lg[:, off:off + w, :], (2, 0, 1)), olog[oname].reshape(n, steps, w), equal_nan=True), name


def test_log_levels_and_stride_are_consistent(nl):
    """The three log layouts (compact / full / MC row) and a log stride > 1 describe the same run: common quantities agree
    bit for bit at the common rows, and logging never changes the trajectory."""
    from rl4afcs_b200 import _lib

    n, steps, every = 40, 210, 7
    rng = np.random.default_rng(6)
    noise = rng.standard_normal((steps, n)).astype(np.float32)
    runs = {}
    for level, ev in ((0, 1), (1, 1), (2, 1), (3, every)):
        eng, st, cfg, th = _setup(nl, n, "mixed", seed=13, fault="damp_all", fault_time=1.0)
        lg = eng.run(steps, noise, log_agents=(n if level else 0), log_every=ev, log_level=max(level, 1))
        runs[level] = (eng, None if lg is None else lg.cpu().numpy())
    base = runs[0][0]
    for level in (1, 2, 3):
        e = runs[level][0]
        same = lambda u, v: bool(((u == v) | (torch.isnan(u) & torch.isnan(v))).all())   # noqa: E731
        assert same(base.env, e.env) and same(base.net, e.net) and torch.equal(base.ints, e.ints), level
    l1, l2, l3 = runs[1][1], runs[2][1], runs[3][1]
    F, M, L = _lib.NLF_FIELDS, _lib.NLM, _lib.NLL
    assert l3.shape == (steps // every, M["COUNT"], n) and l2.shape == (steps, _lib.NLF["COUNT"], n)
    rows = np.arange(0, steps, every)
    xf = F["XFULL"][0]
    assert np.array_equal(l1[:, L["XFULL"]:L["XFULL"] + 12], l2[:, xf:xf + 12])
    assert np.array_equal(l1[:, L["E_THETA"]], l2[:, F["E"][0]]) and np.array_equal(l1[:, L["SURF"]], l2[:, F["A_CMD"][0]])
    for mname, col in (("E", F["E"][0]), ("THETA", xf + 7), ("ALPHA", xf + 4), ("Q", xf + 1), ("V", xf + 3), ("H", xf + 9),
                       ("A_CMD", F["A_CMD"][0]), ("A_EFF", F["A_EFF"][0]), ("RLS_EPS", F["RLS_EPS_NORM"][0])):
        assert np.array_equal(l3[:, M[mname]], l2[rows, col]), mname
    o, w = F["A_W1"]
    assert np.allclose(l3[:, M["WA_NORM"]], np.sqrt((l2[rows, o:o + w] ** 2).sum(axis=1)), rtol=1e-14)


def test_per_agent_configuration_bit_exact(nl):
    """Every agent with its own hyper-parameters, trace mode, multistep switch and fault (family, time, severity): the
    kernel's per-agent instantiation against the oracle run with one configuration record per agent, free run, bit for
    bit; and the same batch with the overrides removed equals the uniform instantiation's result for agent 0's settings."""
    from rl4afcs_b200 import _lib, nl_engine

    n, steps = 96, 700
    rng = np.random.default_rng(17)
    names = ["none", "damp_elevator", "damp_all", "shift_cg", "slow_all", "saturate_elevator", "damp_elevator_and_saturate_elevator"]
    cfg = nl.make_cfg(n)
    per = dict(eta_a_h=rng.uniform(20, 40, n), eta_a_l=rng.uniform(3, 8, n), eta_c_h=rng.uniform(0.8, 1.6, n), eta_c_l=rng.uniform(0.4, 0.8, n),
               lambda_h=rng.uniform(0.85, 0.98, n), lambda_l=rng.uniform(0.6, 0.9, n), damp_factor=rng.uniform(0.2, 0.5, n),
               cg_shift=rng.uniform(-0.5, 0.0, n), lr_decay=rng.uniform(0.99, 0.999, n))
    for k, v in per.items():
        cfg[k] = v
    picks = rng.integers(0, len(names), n)
    ds = np.asarray([nl_engine.split_fault(names[p]) for p in picks], dtype=np.int32)
    cfg["fault_damp"], cfg["fault_sat"] = ds[:, 0], ds[:, 1]
    cfg["fault_step"] = np.where((ds[:, 0] == 0) & (ds[:, 1] == 0), -1, rng.integers(100, 500, n)).astype(np.int32)
    cfg["multistep"] = rng.integers(0, 2, n)
    cfg["elig_a"] = rng.integers(0, 3, n)
    cfg["warmup_steps"] = rng.integers(100, 300, n)
    cfg["cooldown_steps"] = rng.integers(50, 250, n)
    w = nl.init_weights(n, 23)
    st = nl.init_states("mixed", cfg, w, n)
    th = nl.theta_reference()
    noise = rng.standard_normal((steps, n)).astype(np.float32)
    nl.run("mixed", cfg, th, noise, st, 0, steps, tanh="t13")
    eng = nl_engine.NlEngine(n, policy="mixed")
    for k, v in per.items():
        eng.set_hp(k.upper(), v)
    for k in ("fault_damp", "fault_sat", "fault_step", "multistep", "elig_a", "warmup_steps", "cooldown_steps"):
        eng.set_hpi(k.upper(), cfg[k])
    eng.set_reference(th)
    eng.init(w["W1a"], w["W2a"], w["W1c"], w["W2c"])
    eng.run(steps, noise)
    got = _util_nl.engine_to_oracle(eng, nl)
    for f in _EXACT_FIELDS:
        assert np.array_equal(got[f], st[f], equal_nan=(got[f].dtype.kind == "f")), f
