"""The TRANSLATED plant binary (oracle/pe_probe/lift.py -> oracle/_ref/libcitation_lifted_*.so: the reference's
`_citation.cp39-win_amd64.pyd`, envs/nonlinear/citation.py:62-69, turned instruction by instruction into portable C).

* Everywhere the compiled translation exists: it reproduces the golden trajectories of the binary BIT FOR BIT (the fixtures
  were produced by the binary executing natively with the same glibc), instances are independent (the reference's module is
  process-global, the translation is re-entrant).
* Where /root/reference exists: translation == binary on fresh random three-axis inputs, configuration changes and on the
  one-step map from random states -- every double identical."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
TRIM = np.array([-0.02855, 0, 0, 0, 0, 0, 0, 0, 0.55, 0.55, 0])


@pytest.fixture(scope="module")
def lifted():
    from oracle.pe_probe import lifted as lf

    if not lf.available():
        pytest.skip("neither the compiled translation nor the reference's plant binary is present")
    lf.lib()
    return lf


@pytest.mark.parametrize("name", ["trim", "elevator_doublet", "elevator_doublet_large", "elevator_step", "aileron_rudder", "shift_cg",
                                  "damped_elevator"])
def test_translation_reproduces_the_binarys_golden_trajectories_bitwise(lifted, name):
    g = np.load(os.path.join(GOLD, f"citation_{name}.npz"))
    ac = lifted.Aircraft()
    ac.initialize()
    x = np.array([ac.step(u) for u in g["u"]])
    assert np.array_equal(x, g["x"])
    assert np.array_equal(ac.get_state()[1], g["engine_final"])


def test_instances_are_independent_and_reinitialisable(lifted):
    a, b = lifted.Aircraft(), lifted.Aircraft()
    a.initialize(); b.initialize()
    u2 = TRIM.copy(); u2[0] -= 0.05; u2[1] = 0.02
    xa = a.run(TRIM, 300)
    xb = b.run(u2, 300)                                     # a different flight on the second aircraft ...
    a2 = lifted.Aircraft(); a2.initialize()
    assert np.array_equal(a2.run(TRIM, 300), xa)            # ... does not disturb the first one's trajectory
    assert not np.array_equal(xa[-1], xb[-1])
    b.initialize()
    assert np.array_equal(b.run(TRIM, 300), xa)             # initialize() restores the initial condition


def _binary():
    from oracle.pe_probe import pe_citation as pc

    if not pc.available():
        pytest.skip("the reference tree (and its plant binary) is not present on this machine")
    pc.open_variant("extended_input")
    return pc


def test_translation_equals_the_native_binary_on_random_inputs(lifted):
    pc = _binary()
    rng = np.random.default_rng(11)
    ac = lifted.Aircraft()
    ac.initialize(); pc.initialize()
    u = TRIM.copy()
    for k in range(1500):
        if k % 40 == 0:                                      # piecewise-constant commands on every input channel
            u = TRIM.copy()
            u[0:3] += rng.uniform(-0.08, 0.08, 3)
            u[6] = rng.choice([0.0, 0.3, 1.0]); u[7] = rng.choice([0.0, 1.0])
            u[8:10] = rng.uniform(0.3, 0.9, 2)
            u[10] = rng.choice([0.0, -0.5, 0.3])
        assert np.array_equal(ac.step(u), pc.step(u)), k


def test_translation_equals_the_native_binary_on_the_one_step_map(lifted):
    pc = _binary()
    rng = np.random.default_rng(5)
    n = 400
    x = np.tile([0, 0, 0, 90, 0.0576, 0, 0, 0.0576, 0, 2000, 0, 0.], (n, 1))
    x[:, 0:3] += rng.uniform(-0.4, 0.4, (n, 3)); x[:, 3] += rng.uniform(-30, 40, n); x[:, 4] += rng.uniform(-0.1, 0.2, n)
    x[:, 5] += rng.uniform(-0.1, 0.1, n); x[:, 6:9] += rng.uniform(-0.5, 0.5, (n, 3)); x[:, 9] += rng.uniform(-1500, 3000, n)
    u = np.tile(TRIM, (n, 1)); u[:, 0:3] += rng.uniform(-0.2, 0.2, (n, 3)); u[:, 8:10] = rng.uniform(0.2, 1.0, (n, 2))
    pc.initialize(); pc.step(TRIM)
    e0 = np.tile(pc.get_state()[1], (n, 1)) * rng.uniform(0.8, 1.2, (n, 4))
    xb, eb = pc.onestep(x, e0, u)
    ac = lifted.Aircraft(); ac.initialize(); ac.step(TRIM)
    for i in range(n):
        ac.set_state(x[i], e0[i])
        ac.step(u[i])
        xi, ei = ac.get_state()
        assert np.array_equal(xi, xb[i]) and np.array_equal(ei, eb[i]), i
