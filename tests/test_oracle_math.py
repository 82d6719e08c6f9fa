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
