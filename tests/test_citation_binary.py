"""The reference's REAL plant binary (`_citation.cp39-win_amd64.pyd`, run in-process by oracle/pe_probe) against its golden
trajectories, and the surrogate plant of include/rl4_citation_surrogate.h against the same trajectories.

* Where /root/reference exists (the build container) the binary is mapped and re-run: bit for bit equal to the committed
  fixtures (tests/golden/citation_*.npz), deterministic, output-then-update, 16 continuous states.
* Everywhere (the fixtures travel): the surrogate replays the fixtures' inputs from the same initial state and must stay
  inside the DOCUMENTED fidelity bounds -- this is a stand-in calibrated against the binary, not a restatement, so the
  bounds are error budgets (profiles/citation_fidelity_r02.json), not parity."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
SCENARIOS = ["trim", "elevator_doublet", "elevator_doublet_large", "elevator_step", "aileron_rudder", "shift_cg", "damped_elevator"]


def _binary():
    from oracle.pe_probe import pe_citation as pc

    if not pc.available():
        pytest.skip("the reference tree (and its plant binary) is not present on this machine")
    pc.open_variant("extended_input")
    return pc


@pytest.mark.parametrize("name", ["trim", "elevator_doublet", "shift_cg"])
def test_binary_reproduces_its_golden_trajectories(name):
    pc = _binary()
    g = np.load(os.path.join(GOLD, f"citation_{name}.npz"))
    pc.initialize()
    x = np.array([pc.step(u) for u in g["u"]])
    assert np.array_equal(x, g["x"])


def test_binary_is_output_then_update_with_sixteen_states():
    """What the loader established about the plant's contract (oracle/pe_probe/README.md)."""
    pc = _binary()
    trim = np.array([-0.02855, 0, 0, 0, 0, 0, 0, 0, 0.55, 0.55, 0])
    pc.initialize()
    first = pc.step(trim)
    assert np.array_equal(first, [0, 0, 0, 90, 0.0576, 0, 0, 0.0576, 0, 2000, 0, 0])          # the initial condition, not a stepped state
    pc.initialize(); pc.step(trim)
    u2 = trim.copy(); u2[0] -= 0.1
    a = pc.step(u2)
    pc.initialize(); pc.step(trim)
    b = pc.step(trim)
    assert np.array_equal(a, b)                              # the output of call k does not depend on the input of call k
    x, e = pc.get_state()
    nxt = pc.step(trim)
    assert np.array_equal(nxt, x)                            # ... it IS the state the plant carried into the call
    # the 16 states determine the one-step map: same (x, engine, u) after different histories -> same successor
    xs = np.array([[0, 0.01, 0, 85, 0.07, 0, 0, 0.05, 0, 1500, 0, 0.]])
    r1 = pc.onestep(xs, e[None], trim[None])[0]
    for _ in range(25):
        pc.step(u2)
    r2 = pc.onestep(xs, e[None], trim[None])[0]
    assert np.array_equal(r1, r2)


# error budgets of the calibrated surrogate after the reset phase: (rms error) / (peak-to-peak excursion of the binary)
BUDGET = {"elevator_doublet": {"q": 0.12, "alpha": 0.12, "theta": 0.15}, "elevator_doublet_large": {"q": 0.12, "alpha": 0.10, "theta": 0.25},
          "elevator_step": {"q": 0.20, "alpha": 0.20, "theta": 0.25}, "damped_elevator": {"q": 0.12, "alpha": 0.12, "theta": 0.15},
          "shift_cg": {"q": 0.10, "alpha": 0.20, "theta": 0.12}}
IDX = {"q": 1, "alpha": 4, "theta": 7}


@pytest.mark.parametrize("name", list(BUDGET))
def test_surrogate_plant_stays_inside_its_fidelity_budget(name):
    from oracle.pe_probe.make_citation_golden import run_surrogate

    g = np.load(os.path.join(GOLD, f"citation_{name}.npz"))
    s = run_surrogate(g["u"])
    x = g["x"]
    w = slice(1001, None)
    for st, bound in BUDGET[name].items():
        j = IDX[st]
        rel = np.sqrt(np.mean((s[w, j] - x[w, j]) ** 2)) / np.ptp(x[w, j])
        assert rel < bound, (name, st, rel)


def test_surrogate_holds_the_reference_trim_like_the_binary():
    """After Ce500NonLinear.reset's 1001 settling calls both aircraft sit at the reference's trim point (the binary has a slow
    phugoid residue; the surrogate is an exact equilibrium there)."""
    from oracle.pe_probe.make_citation_golden import run_surrogate

    g = np.load(os.path.join(GOLD, "citation_trim.npz"))
    s = run_surrogate(g["u"])
    for x in (g["x"], s):
        assert np.all(np.abs(x[1000:, 3] - 90.0) < 0.2) and np.all(np.abs(x[1000:, 4] - 0.0576) < 2.5e-3)
        assert np.all(np.abs(x[1000:, 7] - 0.0576) < 4e-3) and np.all(np.abs(x[1000:, 1]) < 5e-4)
        assert not x[:, [0, 2, 5, 6, 8, 11]].any()            # symmetric flight stays symmetric
