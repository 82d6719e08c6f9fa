"""Maths checks of the restated TensorFlow parts (no TF here, so closed forms are validated
against torch.autograd) and of the t13 tanh."""
import numpy as np
import pytest


def test_t13_tanh_accuracy(oracle):
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 200
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.uniform(-20, 20, 1500), rng.uniform(-1, 1, 1500), rng.uniform(-1e-2, 1e-2, 1500),
                         10.0 ** rng.uniform(-300, -2, 500)])
    got = oracle.tanh_t13(xs)
    worst = 0.0
    for x, v in zip(xs, got):
        ex = mp.tanh(mp.mpf(float(x)))
        worst = max(worst, float(abs((mp.mpf(float(v)) - ex) / mp.mpf(float(np.spacing(abs(float(ex))))))))
    assert worst < 2.5, worst
    xs32 = xs[np.abs(xs) > 1e-30].astype(np.float32)
    got32 = oracle.tanh_t13(xs32)
    err = np.abs(got32.astype(np.float64) - np.tanh(xs32.astype(np.float64))) / np.spacing(np.abs(got32))
    assert err.max() < 2.5, err.max()
    assert oracle.tanh_t13(np.array([0.0, np.inf, -np.inf, 25.0])).tolist() == [0.0, 1.0, -1.0, 1.0]
    assert np.isnan(oracle.tanh_t13(np.array([np.nan]))[0])


def test_traces_and_dadz_equal_autograd(oracle):
    """Critic / Actor Jacobian traces (objects.py:161-188, 236-254) and da/dz (objects.py:876-878)
    of the fp64 oracle vs torch.autograd on the same 1-4-k bias-free tanh nets."""
    torch = pytest.importorskip("torch")
    n, steps = 6, 40
    base, amp = oracle.default_reference()
    cfg = oracle.make_cfg()
    rng = np.random.default_rng(3)
    x0 = np.deg2rad(rng.uniform(-2, 2, size=(n, 2)))
    w = oracle.init_weights(n, 3)
    st = oracle.init_states("fp64", cfg, x0, w)
    for k in range(steps):
        pre = st.copy()
        lg = oracle.run("fp64", cfg, base, st, k, 1, tanh="libm", n_log=n)[:, 0]
        for i in range(n):
            z = torch.tensor([[lg["e"][i]]], dtype=torch.float64, requires_grad=True)
            W1a = torch.tensor(pre["W1a"][i].reshape(1, 4), requires_grad=True)
            W2a = torch.tensor(pre["W2a"][i].reshape(4, 1), requires_grad=True)
            a = torch.tanh(torch.tanh(z @ W1a) @ W2a)
            gz, g1, g2 = torch.autograd.grad(a.sum(), [z, W1a, W2a])
            # 1 - h*h (objects.py:131) cancels for saturated units: compare relative to the vector scale
            close = lambda u, v: np.allclose(u, v, rtol=1e-9, atol=1e-9 * max(np.abs(v).max(), 1e-300))  # noqa: E731
            assert close(lg["dadz"][i], gz.numpy().ravel())
            assert close(lg["a_e"][i][:4], g2.numpy().ravel())
            assert close(lg["a_e"][i][4:], g1.numpy().ravel())
            W1c = torch.tensor(pre["W1c"][i].reshape(1, 4), requires_grad=True)
            W2c = torch.tensor(pre["W2c"][i].reshape(4, 2), requires_grad=True)
            lam = torch.tanh(z @ W1c) @ W2c
            assert np.allclose(lg["lam"][i], lam.detach().numpy().ravel(), rtol=1e-12, atol=1e-15)
            for q in range(2):
                g1, g2 = torch.autograd.grad(lam[0, q], [W1c, W2c], retain_graph=True)
                row = lg["c_e"][i][12 * q:12 * q + 12]
                assert close(row[0:8], g2.numpy().T.ravel())                  # layout objects.py:203
                assert close(row[8:12], g1.numpy().ravel())


def test_rls_identifies_the_discrete_plant(oracle):
    """functions.py:462-470 prints I + dt*A and dt*B next to the identified params; with the action
    regressor normalised (objects.py:955,973) G converges to dt*B*20*pi/180."""
    base, amp = oracle.default_reference()
    cfg = oracle.make_cfg()
    w = oracle.init_weights(1, 4)
    st = oracle.init_states("mixed", cfg, np.zeros((1, 2)), w)
    oracle.run("mixed", cfg, base, st, 0, 3000, tanh="libm")
    A = cfg["A"][0].reshape(2, 2); B = cfg["B"][0].reshape(2, 1)
    theta = st["theta"][0].reshape(3, 2)
    F, G = theta[:2].T, theta[2:].T
    assert np.allclose(F, np.eye(2) + 0.02 * A, atol=2e-3)
    assert np.allclose(G, 0.02 * B * 20 * np.pi / 180, atol=2e-4)


def test_quirk_flags_change_what_they_should(oracle):
    base, amp = oracle.default_reference()
    w = oracle.init_weights(1, 4)
    x0 = np.array([[0.02, -0.03]])
    out = {}
    for q3 in (0, 1):
        cfg = oracle.make_cfg(q3_alias=q3)
        st = oracle.init_states("mixed", cfg, x0, w)
        oracle.run("mixed", cfg, base, st, 0, 3, tanh="t13")
        out[q3] = st["theta"][0].copy()
    assert not np.array_equal(out[0], out[1])               # Q3 changes the first RLS regressor
    cd = {}
    for q7 in (0, 1):
        cfg = oracle.make_cfg(q7_numpy1=q7)
        st = oracle.init_states("mixed", cfg, x0, w)
        oracle.run("mixed", cfg, base, st, 0, 2, tanh="t13")
        cd[q7] = int(st["cooldown"][0])
    assert cd == {0: 0, 1: 100}                             # Q7: cooldown starts at k = 1 under numpy 1.x


def test_symmetric_flight_plant_equals_full_model(oracle):
    """include/rl4_citation_surrogate.h: the longitudinal form of the plant step (taken by the CUDA kernel in symmetric
    flight) returns the same values as the full 6-DOF step for symmetric states -- both integrators, random states,
    controls, flap / gear / c.g. settings, several consecutive steps; a non-symmetric state is refused."""
    import ctypes

    from oracle import nl_c

    L = nl_c.lib()
    vp = ctypes.c_void_p
    L.orc_cit_plant_step.argtypes = [vp, vp, vp, ctypes.c_double, ctypes.c_int]
    L.orc_cit_plant_step_lon.argtypes = [vp, vp, vp, ctypes.c_double, ctypes.c_int]
    L.orc_cit_plant_step_lon.restype = ctypes.c_int
    plant = np.ascontiguousarray(nl_c.make_cfg()["plant"][:1])
    rng = np.random.default_rng(0)
    for integ in (0, 1):
        for _ in range(1500):
            x = np.zeros(12)
            x[1], x[3], x[4], x[7] = rng.normal(0, 0.3), rng.uniform(40, 160), rng.normal(0.06, 0.2), rng.normal(0.06, 0.4)
            x[9], x[10], x[11] = rng.uniform(0, 6000), rng.uniform(0, 1e4), rng.choice([0.0, 3.5])
            u = np.array([rng.uniform(-0.3, 0.3), 0, 0, -0.02855, 0, 0, rng.choice([0, 0.2]), rng.choice([0, 1.0]), rng.uniform(0, 1),
                          rng.uniform(0, 1), rng.choice([0, -0.5])])
            a, b = x.copy(), x.copy()
            for _step in range(4):
                L.orc_cit_plant_step(plant.ctypes.data, a.ctypes.data, u.ctypes.data, 0.01, integ)
                assert L.orc_cit_plant_step_lon(plant.ctypes.data, b.ctypes.data, u.ctypes.data, 0.01, integ) == 1
            assert np.array_equal(a, b)
    # degenerate states (huge / tiny airspeed, huge rates and angles, i.e. agents about to diverge): the guarded step
    # (in-range check before and after, else the full equations) reproduces the full model including its inf / nan pattern
    rng = np.random.default_rng(1)
    for integ in (0, 1):
        for _ in range(2500):
            x = np.zeros(12)
            x[1] = rng.normal(0, 0.3) * 10.0 ** rng.choice([0, 0, 3, 8, 20, 40, 80, 150, 300])
            x[3] = rng.uniform(40, 160) * 10.0 ** rng.choice([0, 0, -5, -20, -40, 5, 20, 60, 160])
            x[4] = rng.normal(0.06, 0.2) * rng.choice([1, 1, 50, 1e5]); x[7] = rng.normal(0.06, 0.4) * rng.choice([1, 1, 50, 1e5])
            x[9] = rng.uniform(0, 6000)
            u = np.array([rng.uniform(-0.3, 0.3), 0, 0, -0.02855, 0, 0, 0, 0, 0.55, 0.55, rng.choice([0, -0.5])])
            a, b = x.copy(), x.copy()
            for _step in range(5):
                L.orc_cit_plant_step(plant.ctypes.data, a.ctypes.data, u.ctypes.data, 0.01, integ)
                if L.orc_cit_plant_step_lon(plant.ctypes.data, b.ctypes.data, u.ctypes.data, 0.01, integ) == 0:
                    L.orc_cit_plant_step(plant.ctypes.data, b.ctypes.data, u.ctypes.data, 0.01, integ)
                assert np.array_equal(a, b, equal_nan=True)
    x = np.zeros(12); x[3] = 90.0; x[5] = 1e-9                                   # a sideslip: not symmetric
    u = np.zeros(11)
    assert L.orc_cit_plant_step_lon(plant.ctypes.data, x.ctypes.data, u.ctypes.data, 0.01, 1) == 0
    x[5] = 0.0; u[1] = 0.01                                                      # an aileron input: not symmetric
    assert L.orc_cit_plant_step_lon(plant.ctypes.data, x.ctypes.data, u.ctypes.data, 0.01, 1) == 0
