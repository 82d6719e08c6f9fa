"""TEST INFRASTRUCTURE ONLY -- the VERBATIM reference agent (objects.py's IDHPnonlin on the TensorFlow stand-in, verbatim
Ce500NonLinear wrapper) flown on the reference's REAL plant binary and on the calibrated surrogate, same initial weights
and noise stream, idhp_nonlin.py hyper-parameters.  Reports the tracking statistics of both (profiles/citation_closed_loop_r02.json).

    python oracle/pe_probe/closed_loop_compare.py [--steps 3000] [--seeds 2 3]

One plant per process (the binary keeps process-global state), so each run is a subprocess."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def one_run(plant, seed, steps):
    from oracle import make_golden, nl_c, ref_loader, sp_c

    O, tf = ref_loader.load_reference_objects()
    Env, stub = ref_loader.load_reference_nonlinear_env("ode5", plant=plant)
    th = nl_c.theta_reference()
    trim_input = np.array([-0.02855, 0, 0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.55, 0.55, 0])
    trim_state = np.array([0, 0, 0, 90, 0.0576, 0, 0, 0.0576, 0, 2000, 0, 0])
    env_config = {"state_dim": 4, "action_dim": 3, "trim_input": trim_input, "trim_state": trim_state, "dt": 0.01,
                  "t_end": steps * 0.01, "total_steps": steps, "fault_time": 60, "fault_scenario": "none",
                  "reference": {"tracked_state": ["phi", "theta", "psi"], "signal": [0 * th, th, 0 * th]}}
    idhp_config = {"gamma": 0.6, "multistep": 0, "lr_decay": 0.998, "lambda_h": 0.95, "lambda_l": 0.95, "kappa": [1, 2, 1],
                   "cooldown_time": 2.0, "sigma": 0.1, "warmup_time": 4.0, "error_thresh": 1, "tau": 0.02, "in_dims": 4,
                   "actor_config": {"layers": {10: "tanh", 1: "tanh"}, "eta_h": 35.0, "eta_l": 5.0, "elig": "accumulating"},
                   "critic_config": {"layers": {10: "tanh", 3: "linear"}, "eta_h": 1.4, "eta_l": 0.7, "elig": 1233},
                   "rls_config": {"state_dim": 3, "action_dim": 1, "rls_gamma": 1, "rls_cov": 10 ** 6}}     # idhp_nonlin.py:123-146
    noise = np.random.default_rng(seed).standard_normal(steps).astype(np.float32)
    tf.set_tanh(lambda v: sp_c.tanh_t13(np.asarray(v, dtype=np.float32)))
    tf.set_noise(noise[1:])
    env = Env(env_config)
    idhp = O.IDHPnonlin(env, idhp_config, verbose=False, seed=seed)
    w = nl_c.init_weights(1, seed)
    idhp.actor.set_weights([w["W1a"][0].reshape(4, 10), w["W2a"][0].reshape(10, 1)])
    idhp.critic.set_weights([w["W1c"][0].reshape(4, 10), w["W2c"][0].reshape(10, 3)])
    idhp.train()
    L = idhp.log
    x = np.asarray(L["x_full"], dtype=np.float64)
    e = np.asarray(L["e"], dtype=np.float64).ravel()
    ok = np.isfinite(e)
    n_ok = int(ok.sum())
    nz = x[:, 3] * x[:, 1] / 9.80665
    late = slice(min(1000, n_ok), n_ok)
    return {"plant": plant, "seed": seed, "steps": steps, "steps_flown": n_ok,
            "mean_abs_theta_error_deg": float(np.rad2deg(np.mean(np.abs(e[ok])))),
            "mean_abs_theta_error_after_10s_deg": float(np.rad2deg(np.mean(np.abs(e[late])))) if n_ok > 1000 else None,
            "peak_abs_nz_g": float(np.nanmax(np.abs(nz))), "peak_abs_q_deg_s": float(np.rad2deg(np.nanmax(np.abs(x[:, 1])))),
            "elevator_rms_deg": float(np.rad2deg(np.sqrt(np.nanmean(np.asarray(L["a_cmd"], dtype=np.float64) ** 2))))}


if __name__ == "__main__":
    if "--one" in sys.argv:
        i = sys.argv.index("--one")
        print("RESULT " + json.dumps(one_run(sys.argv[i + 1], int(sys.argv[i + 2]), int(sys.argv[i + 3]))))
        sys.exit(0)
    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 3000
    seeds = [int(s) for s in sys.argv[sys.argv.index("--seeds") + 1:]] if "--seeds" in sys.argv else [2, 3]
    procs = [(p, s, subprocess.Popen([sys.executable, os.path.abspath(__file__), "--one", p, str(s), str(steps)],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True))
             for s in seeds for p in ("binary", "surrogate")]
    rows = []
    for p, s, pr in procs:
        out = pr.communicate()[0]
        line = [ln for ln in out.splitlines() if ln.startswith("RESULT ")]
        rows.append(json.loads(line[0][7:]) if line else {"plant": p, "seed": s, "failed": True})
        print(rows[-1])
    json.dump({"what": "verbatim IDHPnonlin (idhp_nonlin.py hyper-parameters) on the reference's plant binary vs the calibrated surrogate",
               "runs": rows}, open(os.path.join(ROOT, "profiles", "citation_closed_loop_r02.json"), "w"), indent=1)
