"""TEST INFRASTRUCTURE ONLY -- golden trajectories of the reference's REAL plant binary, run in-process by pe_citation.

    python oracle/pe_probe/make_citation_golden.py            # writes tests/golden/citation_*.npz
    python oracle/pe_probe/make_citation_golden.py --report   # + surrogate-vs-binary error per state (profiles/citation_fidelity_r02.json)

Needs /root/reference (build container).  Each fixture holds the inputs u [N][11] and what `citation.step(u[k])` RETURNED
x [N][12] (the state before step k: the binary is an output-then-update block), starting right after `initialize()`.
tests/test_citation_binary.py replays them: against the binary itself where the reference exists (bit for bit), and against
the surrogate plant everywhere (error bounds = the documented fidelity of the stand-in)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pe_probe import pe_citation as pc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
TRIM = np.array([-0.02855, 0, 0, 0, 0, 0, 0, 0, 0.55, 0.55, 0])       # idhp_nonlin.py:53


def scenario_inputs():
    """name -> u [N][11]; every scenario starts with the 1001 trim calls of Ce500NonLinear.reset (envs/nonlinear/env.py:288-291)."""
    n = 1001 + 900
    k = np.arange(n) - 1001
    def base():
        return np.tile(TRIM, (n, 1))
    out = {}
    out["trim"] = np.tile(TRIM, (1001 + 3000, 1))
    d = np.deg2rad(2.0) * (((k >= 100) & (k < 200)).astype(float) - ((k >= 200) & (k < 300)).astype(float))
    u = base(); u[:, 0] += d; out["elevator_doublet"] = u
    u = base(); u[:, 0] += 3.0 * d; out["elevator_doublet_large"] = u
    u = base(); u[:, 0] += np.deg2rad(-3.0) * (k >= 100); out["elevator_step"] = u
    u = base(); u[:, 1] += 0.5 * d; u[:, 2] += -0.5 * np.roll(d, 150); out["aileron_rudder"] = u
    u = base(); u[:, 10] = np.where(k >= 200, -0.5, 0.0); out["shift_cg"] = u                       # envs/nonlinear/env.py:141-143
    u = base(); u[:, 0] = TRIM[0] + 0.3 * (d + TRIM[0] * 0); u[k >= 400, 0] *= 1.0; out["damped_elevator"] = u   # eff[0] *= 0.3 (env.py:135)
    return out


def run_binary(u):
    pc.initialize()
    return np.array([pc.step(row) for row in u])


def run_surrogate(u, integrator="ode5"):
    import ctypes

    from oracle import nl_c

    L = nl_c.lib()
    L.orc_cit_plant_step.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_int]
    plant = np.ascontiguousarray(nl_c.make_cfg()["plant"][:1])
    x = np.array([0, 0, 0, 90, 0.0576, 0, 0, 0.0576, 0, 2000, 0, 0], dtype=np.float64)
    out = np.empty((u.shape[0], 12))
    for k in range(u.shape[0]):
        out[k] = x                                           # output-then-update
        L.orc_cit_plant_step(plant.ctypes.data, x.ctypes.data, np.ascontiguousarray(u[k]).ctypes.data, 0.01, nl_c.INTEGRATOR[integrator])
    return out


STATE_NAMES = ["p", "q", "r", "V", "alpha", "beta", "phi", "theta", "psi", "h", "xe", "ye"]


def main():
    info = pc.open_variant("extended_input")
    report = {"binary": "envs/nonlinear/extended_input/_citation.cp39-win_amd64.pyd", "loader": info, "scenarios": {}}
    for name, u in scenario_inputs().items():
        x = run_binary(u)
        again = run_binary(u)
        assert np.array_equal(x, again), "the binary is not deterministic?"
        eng = pc.get_state()[1]
        np.savez_compressed(os.path.join(GOLD, f"citation_{name}.npz"), u=u, x=x, engine_final=eng)
        if "--report" in sys.argv:
            s = run_surrogate(u)
            w = slice(1001, None)                            # after the reset phase
            err = s[w] - x[w]
            report["scenarios"][name] = {
                "steps_after_reset": int(err.shape[0]),
                "rms_error": {nm: float(np.sqrt(np.mean(err[:, j] ** 2))) for j, nm in enumerate(STATE_NAMES)},
                "max_abs_error": {nm: float(np.abs(err[:, j]).max()) for j, nm in enumerate(STATE_NAMES)},
                "binary_excursion": {nm: float(np.ptp(x[w][:, j])) for j, nm in enumerate(STATE_NAMES)},
                "state_after_reset_binary": x[1000].tolist(), "state_after_reset_surrogate": s[1000].tolist()}
        print(f"citation_{name}.npz: {u.shape[0]} steps; after reset x = {np.array2string(x[1000], precision=5)}")
    if "--report" in sys.argv:
        path = os.path.join(ROOT, "profiles", "citation_fidelity_r02.json")
        json.dump(report, open(path, "w"), indent=1)
        for name, r in report["scenarios"].items():
            print(name, {k: f"{r['rms_error'][k]:.4g} / {r['binary_excursion'][k]:.4g}" for k in ("q", "alpha", "theta", "V", "h")})


if __name__ == "__main__":
    main()
