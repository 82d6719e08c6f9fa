"""Live re-check of the oracle against the verbatim reference code (build container only:
/root/reference is not present on the GPU box, where the committed golden vectors stand in)."""
import os

import numpy as np
import pytest

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference not mounted")


def test_numpy_matmul_order_assumptions():
    """The term orders the oracle hard-codes are those numpy executes on this host."""
    import ctypes
    libm = ctypes.CDLL("libm.so.6"); libm.fma.restype = ctypes.c_double; libm.fma.argtypes = [ctypes.c_double] * 3
    f = libm.fma
    rng = np.random.default_rng(0)
    for _ in range(500):
        A = rng.normal(size=(2, 2)); x = rng.normal(size=(2, 1))
        y = A @ x
        assert y[0, 0] == f(A[0, 0], x[0, 0], A[0, 1] * x[1, 0])                       # (2,2)@(2,1): terms 1,0
        C = rng.normal(size=(3, 3)); X = rng.normal(size=(3, 1))
        z = C @ X
        assert z[1, 0] == f(C[1, 2], X[2, 0], f(C[1, 0], X[0, 0], C[1, 1] * X[1, 0]))   # (3,3)@(3,1): 1,0,2
        P = rng.normal(size=(3, 2))
        w = P.T @ X
        assert w[0, 0] == f(P[2, 0], X[2, 0], f(P[1, 0], X[1, 0], P[0, 0] * X[0, 0]))   # transposed lhs: in order
        d = X.T @ z
        assert d[0, 0] == f(X[2, 0], z[2, 0], f(X[1, 0], z[1, 0], X[0, 0] * z[0, 0]))
        assert np.linalg.norm(x) == np.sqrt(f(x[1, 0], x[1, 0], x[0, 0] * x[0, 0]))


@pytest.mark.parametrize("seed,fault,elig,ms", [
    (21, None, (None, None), 2), (22, "damp_elevator", ("accumulating", None), 2), (23, "shift_cg", (None, "replacing"), 0)])
def test_c_oracle_equals_numpy_loop_on_verbatim_objects(oracle, seed, fault, elig, ms):
    from oracle import sp_numpy

    Env = ref_loader.load_reference_linear_env(); RLS = ref_loader.load_reference_rls()
    base, amp = oracle.default_reference()
    ic = oracle.default_idhp_config(); ic["multistep"] = ms
    ic["actor_config"]["elig"], ic["critic_config"]["elig"] = elig
    rng = np.random.default_rng(seed)
    x0 = np.deg2rad(rng.uniform(-2, 2, size=2))
    steps = 1300
    env = Env({"state_dim": 2, "action_dim": 1, "x0": x0.reshape(2, 1).copy(), "dt": 0.02, "t_end": 60, "fault_time": 20,
               "fault_scenario": fault, "reference": {"tracked_state": ["alpha"], "signal": [amp * base]}})
    w = oracle.init_weights(1, seed)
    loop = sp_numpy.IDHPspLoop(env, ic, {k: v[0] for k, v in w.items()},
                               tanh_fn=lambda a: oracle.tanh_t13(np.asarray(a)), rls=RLS(ic["rls_config"]))
    lg = loop.train(steps)
    cfg = oracle.make_cfg(ic, fault_scenario=fault)
    st = oracle.init_states("mixed", cfg, x0.reshape(1, 2), w)
    cl = oracle.run("mixed", cfg, base, st, 0, steps, tanh="t13", n_log=1)[0]
    for k in ("x", "a", "ref", "a_w1", "a_w2", "c_w1", "c_w2", "a_e", "c_e"):
        assert np.array_equal(lg[k], cl[k], equal_nan=True), k
    for k in ("params", "cov", "eps_norm"):
        assert np.array_equal(lg[k][2:], cl[k][2:], equal_nan=True), k


@pytest.mark.parametrize("seed,fault,elig,ms,tanh,steps,warmup", [
    (31, None, (None, None), 2, "t13", 3000, 3.0),                              # full default episode: LR switch, RLS reset
    (32, "damp_elevator", ("accumulating", None), 0, "libm", 1300, 3.0),
    (33, "invert_elevator", ("replacing", "accumulating"), 2, "t13", 1300, 3.0),
    (34, None, (None, None), 2, "t13", 900, 1.0),       # warm-up shorter than the cooldown: the case where Q7 (numpy 1.x) would matter
])
def test_c_oracle_equals_verbatim_idhpsp_on_tf_standin(oracle, seed, fault, elig, ms, tanh, steps, warmup):
    """The reference's OWN agent code -- objects.py's IDHPsp.train(), Actor / Critic call and trace code, RLS,
    _adapt_check, _log -- executed unmodified on the TensorFlow stand-in (oracle/tf_shim.py supplies only the arithmetic
    of the TF ops) around the verbatim Ce500ShortPeriod, against the C oracle: every logged quantity
    bit for bit (the reward to 1 ulp: `e**2` is libm pow in the reference).  numpy >= 2 here, so Q7 is off."""
    O, tf = ref_loader.load_reference_objects()
    Env = ref_loader.load_reference_linear_env()
    base, amp = oracle.default_reference()
    ic = oracle.default_idhp_config(); ic["multistep"] = ms; ic["warmup_time"] = warmup
    ic["actor_config"]["elig"], ic["critic_config"]["elig"] = elig
    rng = np.random.default_rng(seed)
    x0 = np.deg2rad(rng.uniform(-2, 2, size=2))
    fault_time = 8.0
    env = Env({"state_dim": 2, "action_dim": 1, "x0": x0.reshape(2, 1).copy(), "dt": 0.02, "t_end": steps * 0.02,
               "fault_time": fault_time, "fault_scenario": fault, "reference": {"tracked_state": ["alpha"], "signal": [amp * base]}})
    assert int(env.t_end / env.dt) == steps
    tf.set_tanh((lambda a: oracle.tanh_t13(np.asarray(a, dtype=np.float32))) if tanh == "t13" else None)
    try:
        idhp = O.IDHPsp(env, ic, verbose=False, seed=seed)                      # weights drawn by the stand-in's initializer
        w = {"W1a": idhp.actor.get_weights()[0].reshape(1, 4).astype(np.float64), "W2a": idhp.actor.get_weights()[1].reshape(1, 4).astype(np.float64),
             "W1c": idhp.critic.get_weights()[0].reshape(1, 4).astype(np.float64), "W2c": idhp.critic.get_weights()[1].reshape(1, 8).astype(np.float64)}
        idhp.train()
    finally:
        tf.set_tanh(None)
    cfg = oracle.make_cfg(ic, fault_time=fault_time, fault_scenario=fault, q7_numpy1=0)
    st = oracle.init_states("mixed", cfg, x0.reshape(1, 2), w)
    cl = oracle.run("mixed", cfg, base, st, 0, steps, tanh=tanh, n_log=1)[0]
    ref = {"x": idhp.x_hist, "a": idhp.a_hist.reshape(-1), "ref": idhp.ref_hist, "a_w1": idhp.a_weights_hist1, "a_w2": idhp.a_weights_hist2,
           "c_w1": idhp.c_weights_hist1, "c_w2": idhp.c_weights_hist2, "a_e": idhp.a_e_hist, "c_e": idhp.c_e_hist,
           "a_all_grad": idhp.a_all_grad_hist, "c_all_grad": idhp.c_all_grad_hist}
    for k, v in ref.items():
        assert np.array_equal(np.asarray(v, dtype=np.float64).reshape(cl[k].shape), cl[k], equal_nan=True), k
    for k, v in (("params", idhp.params_hist), ("cov", idhp.cov_hist), ("eps_norm", idhp.eps_norm_hist)):
        assert np.array_equal(np.asarray(v, dtype=np.float64).reshape(cl[k].shape)[2:], cl[k][2:], equal_nan=True), k   # objects.py:704
    c_ref = np.asarray(idhp.c_hist, dtype=np.float64)
    assert np.all(np.abs(c_ref - cl["c"]) <= 2 * np.spacing(np.abs(c_ref)))      # pow(e, 2) vs e * e, then * kappa
    # final object state == final oracle state
    assert np.array_equal(idhp.model.params.ravel(), st["theta"][0]) and np.array_equal(idhp.model.Cov.ravel(), st["cov"][0])
    assert np.array_equal(idhp.target_critic.get_weights()[1].ravel().astype(np.float64), st["W2t"][0])


@pytest.mark.parametrize("seed,kw", [
    (51, dict(steps=1300, warmup=2.0)),                                                    # decay + two LR hand-overs
    (52, dict(steps=700, warmup=1.0, ms=1, elig=None, integrator="rk4", fault="shift_cg", fault_time=2.0)),
    (53, dict(steps=700, warmup=1.0, elig="replacing", fault="slow_all", fault_time=2.5)),
])
def test_c_oracle_equals_verbatim_idhpnonlin_on_tf_standin(oracle, seed, kw):
    """Live form of tests/test_oracle_golden.py::test_nl_full_loop_matches_verbatim_idhpnonlin: the verbatim IDHPnonlin
    (TensorFlow stand-in, verbatim env wrapper, surrogate plant) is run here and compared with the C oracle."""
    from oracle import make_golden as mg
    from oracle import nl_c
    from tests.test_oracle_golden import _NL_LOG10_MAP, _NL_LOG_MAP, _nl_loop_cfg, _oracle_nl_view

    g = mg.nl_loop_fixture(seed, **kw)
    steps = int(g["steps"])
    g = {k: np.asarray(v) for k, v in g.items()}
    cfg = _nl_loop_cfg(nl_c, g)
    w = {k: g[f"w_{k}"][None] for k in ("W1a", "W2a", "W1c", "W2c")}
    st = nl_c.init_states("mixed", cfg, w, 1)
    olog = nl_c.run("mixed", cfg, g["theta_ref"], g["noise"].reshape(steps, 1), st, 0, steps, tanh="t13", n_log=1)[0]
    for key, oname in _NL_LOG_MAP:
        assert np.array_equal(g[f"log_{key}"], _oracle_nl_view(olog, oname), equal_nan=True), key
    for key, oname in _NL_LOG10_MAP:
        assert np.array_equal(g[f"log10_{key}"], _oracle_nl_view(olog, oname)[9::10], equal_nan=True), key
    assert np.array_equal(g["final_E"], st["Ea"][0], equal_nan=True)        # case 52 diverges: NaN convention included


def _driver_namespace(script, prefer_else_of=("MC",)):
    """Executes only the plain assignments of a reference driver's `if __name__ == '__main__':` block (descending into the
    single-run branch of `if MC:`), skipping everything that needs TensorFlow / the environments / plotting."""
    import ast

    src = open(os.path.join(ref_loader.REFERENCE_ROOT, script)).read()
    tree = ast.parse(src)
    ns = {"np": np, "NO_TRACE": None, "A_TRACE": "accumulating", "R_TRACE": "replacing", "timeit": __import__("timeit")}

    def run(stmts):
        for st in stmts:
            if isinstance(st, (ast.Assign, ast.AugAssign)):
                try:
                    exec(compile(ast.Module(body=[st], type_ignores=[]), script, "exec"), ns)
                except Exception:
                    pass
            elif isinstance(st, ast.If):
                name = st.test.id if isinstance(st.test, ast.Name) else None
                run(st.orelse if name in prefer_else_of else st.body)
    for node in tree.body:
        if isinstance(node, ast.If) and isinstance(node.test, ast.Compare):
            run(node.body)
    return ns


def test_default_configurations_equal_the_verbatim_drivers(oracle):
    """idhp_sp.py / idhp_nonlin.py: the shipped single-run configurations and reference signals, evaluated from the
    driver sources themselves, equal the defaults the oracle, the C-ABI (`rl4_nl_default_params`) and bench.py use."""
    from oracle import nl_c

    sp = _driver_namespace("idhp_sp.py")
    assert sp["idhp_config"] == oracle.default_idhp_config()
    base, amp = oracle.default_reference()
    assert np.array_equal(sp["ref"], amp * base)                                            # idhp_sp.py:41-44,174
    assert sp["env_config"]["dt"] == 0.02 and sp["env_config"]["t_end"] == 60 and sp["env_config"]["fault_time"] == 20
    assert not sp["env_config"]["x0"].any() and sp["seed"] == 4

    nlns = _driver_namespace("idhp_nonlin.py")
    assert np.array_equal(nlns["theta_ref"], nl_c.theta_reference())                        # idhp_nonlin.py:40-48
    ic, cfg = nlns["idhp_config"], nl_c.make_cfg()[0]
    assert np.array_equal(nlns["trim_input"], cfg["trim_input"]) and nlns["env_config"]["dt"] == cfg["dt"]
    for key, val in (("gamma", ic["gamma"]), ("tau", ic["tau"]), ("lambda_h", ic["lambda_h"]), ("lambda_l", ic["lambda_l"]),
                     ("lr_decay", ic["lr_decay"]), ("eta_a_h", ic["actor_config"]["eta_h"]), ("eta_a_l", ic["actor_config"]["eta_l"]),
                     ("eta_c_h", ic["critic_config"]["eta_h"]), ("eta_c_l", ic["critic_config"]["eta_l"]),
                     ("rls_gamma", ic["rls_config"]["rls_gamma"]), ("rls_cov0", ic["rls_config"]["rls_cov"]), ("Q_sym", ic["kappa"][1])):
        assert cfg[key] == val, key
    assert cfg["multistep"] == (1 if ic["multistep"] > 0 else 0) and cfg["elig_a"] == oracle.ELIG[ic["actor_config"]["elig"]]
    assert cfg["warmup_steps"] == int(ic["warmup_time"] / 0.01) and cfg["cooldown_steps"] == int(ic["cooldown_time"] / 0.01)
    assert nlns["env_config"]["fault_time"] == 60 and nlns["env_config"]["fault_scenario"] == "none" and nlns["seed"] == 8
