/* TEST INFRASTRUCTURE ONLY -- see nl_oracle.h.  Build: make -C oracle. */
#include "nl_oracle.h"
#include "sp_oracle.h"          /* ORC_POLICY_*, ORC_TANH_* */
#include "sp_oracle_tanh.h"
#include <math.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* RLS.update for n = 3, m = 1 (objects.py:492-543).  numpy `@` orders measured for these shapes (header of
 * nl_oracle.h): params.T @ X -> fma(a0,b0,a1*b1) + fma(a2,b2,a3*b3); Cov @ X -> (p0+p2)+(p1+p3) with rounded products;
 * X.T @ (Cov X) and the norm -> in-order FMA chains. */
static void orc_rls3_core(double* th, double* cv, const double* Xr, const double* Y, double gamma, double* eps, double* eps_norm)
{
    double pred[3], CX[4], K[4];
    for (int i = 0; i < 3; ++i)
        pred[i] = fma(th[0 * 3 + i], Xr[0], th[1 * 3 + i] * Xr[1]) + fma(th[2 * 3 + i], Xr[2], th[3 * 3 + i] * Xr[3]);
    for (int i = 0; i < 3; ++i) eps[i] = Y[i] - pred[i];
    for (int i = 0; i < 4; ++i) {
        const double p0 = cv[i * 4 + 0] * Xr[0], p1 = cv[i * 4 + 1] * Xr[1], p2 = cv[i * 4 + 2] * Xr[2], p3 = cv[i * 4 + 3] * Xr[3];
        CX[i] = (p0 + p2) + (p1 + p3);
    }
    double xcx = Xr[0] * CX[0];
    for (int i = 1; i < 4; ++i) xcx = fma(Xr[i], CX[i], xcx);
    const double den = gamma + xcx;
    for (int i = 0; i < 4; ++i) K[i] = CX[i] / den;
    for (int t = 0; t < 4; ++t)
        for (int i = 0; i < 3; ++i) th[t * 3 + i] = th[t * 3 + i] + K[t] * eps[i];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) cv[i * 4 + j] = (cv[i * 4 + j] - K[i] * CX[j]) / gamma;
    *eps_norm = sqrt(fma(eps[2], eps[2], fma(eps[1], eps[1], eps[0] * eps[0])));
}

/* step-level form: theta [n][12], cov [n][16], dx0 [n][3], da0 [n], dx1 [n][3] -> eps [n][3], eps_norm [n] */
void orc_nl_rls_update(double gamma, double* theta, double* cov, const double* dx0, const double* da0, const double* dx1,
                       double* eps, double* eps_norm, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) {
        const double Xr[4] = {dx0[i * 3], dx0[i * 3 + 1], dx0[i * 3 + 2], da0[i]};
        orc_rls3_core(theta + i * 12, cov + i * 16, Xr, dx1 + i * 3, gamma, eps + i * 3, eps_norm + i);
    }
}

/* Optional external plant: `citation.initialize / step` of the reference's OWN model (the translated binary of
 * oracle/pe_probe, one instance per agent) in place of the surrogate header.  Single-threaded use only. */
typedef void (*orc_ext_plant_step_fn)(void* instance, const double* u, double* x_out);
typedef void (*orc_ext_plant_init_fn)(void* instance);
static orc_ext_plant_step_fn g_ext_step;
static orc_ext_plant_init_fn g_ext_init;
static void** g_ext_inst;
static int64_t g_cur_agent;
void orc_nl_set_external_plant(orc_ext_plant_step_fn step, orc_ext_plant_init_fn init, void** instances)
{
    g_ext_step = step; g_ext_init = init; g_ext_inst = instances;
}

/* Ce500NonLinear.step without the agent (envs/nonlinear/env.py:182-256): action scaling (:111-124), rate-limited
 * actuators (:161-180), saturation faults (:150-159), damping / c.g. / slow-actuator faults (:127-148), plant step (:210),
 * errors and the longitudinal reward (:215-220).  stepp is the env's counter BEFORE the step. */
typedef struct { double surf[3], u[11], e[3], reward, rg2, x_obs[12]; } orc_nl_envout;
static void orc_nl_env_core(const orc_nl_cfg* c, double theta_ref_k, const double* act, double* x_full, double* x_act,
                            int32_t stepp, orc_nl_envout* o)
{
    const double dt = c->dt;
    const int faulted = (c->fault_step >= 0 && stepp >= c->fault_step);           /* env.py:132,151 */
    double cmd[3];
    for (int i = 0; i < 3; ++i) {                                                 /* _scale_action env.py:111-124 */
        const double hi = c->limit_deg[i], lo = -c->limit_deg[i];
        double v = act[i] * (hi - lo) / 2.0;
        v = v + (hi + lo) / 2.0;
        cmd[i] = v * (M_PI / 180.0);
    }
    const double omega = (faulted && c->fault_damp == ORC_NL_SLOW_ALL && stepp > c->fault_step) ? c->omega_slow
                         : c->omega0;   /* slow_all takes effect from the step AFTER it is first engaged (env.py:145 runs after :205) */
    for (int i = 0; i < 3; ++i) {                                                 /* _propagate_surfaces_states env.py:161-180 */
        double d = cmd[i] - x_act[i];
        d = d * omega;
        d = d < -c->rate_limit ? -c->rate_limit : (d > c->rate_limit ? c->rate_limit : d);
        x_act[i] = x_act[i] + dt * d;
        o->surf[i] = x_act[i];
    }
    if (faulted && c->fault_sat != ORC_NL_SAT_NONE) {                             /* _saturate_surfaces env.py:150-159 */
        const int j = c->fault_sat - 1;
        const double L = c->sat_limit[j];
        o->surf[j] = o->surf[j] < -L ? -L : (o->surf[j] > L ? L : o->surf[j]);
    }
    double eff[11] = {0};
    for (int i = 0; i < 3; ++i) eff[i] = o->surf[i];
    if (faulted) {                                                                /* _engage_fault env.py:129-148 */
        switch (c->fault_damp) {
        case ORC_NL_DAMP_ELEVATOR: eff[0] *= c->damp_factor; break;
        case ORC_NL_DAMP_AILERON:  eff[1] *= c->damp_factor; break;
        case ORC_NL_DAMP_RUDDER:   eff[2] *= c->damp_factor; break;
        case ORC_NL_DAMP_ALL:      for (int i = 0; i < 3; ++i) eff[i] *= c->damp_factor; break;
        case ORC_NL_SHIFT_CG:      eff[10] = c->cg_shift; break;
        default: break;
        }
    }
    for (int i = 0; i < 11; ++i) o->u[i] = c->trim_input[i] + eff[i];             /* env.py:207-208 */
    /* env.py:210  x_full = model.step(input).  The reference's plant binary is an output-then-update block: step() RETURNS
     * the state before the step and then integrates (oracle/pe_probe/README.md), so the wrapper observes the aircraft one
     * sample late.  x_full (the carried state) is advanced; everything the wrapper computes uses the returned x_obs. */
    if (g_ext_step) {               /* external plant (the reference's own model): step(u) returns the state before the step */
        g_ext_step(g_ext_inst[g_cur_agent], o->u, o->x_obs);
        memcpy(x_full, o->x_obs, sizeof o->x_obs);
    } else {
        memcpy(o->x_obs, x_full, sizeof o->x_obs);
        if (c->integrator == RL4_CIT_INTEGRATOR_RK4) rl4_cit_step_rk4(&c->plant, x_full, o->u, dt);
        else rl4_cit_step_ode5(&c->plant, x_full, o->u, dt);
    }
    o->e[0] = o->x_obs[6] - 0.0; o->e[1] = o->x_obs[7] - theta_ref_k; o->e[2] = o->x_obs[8] - 0.0;   /* env.py:215 (state - ref) */
    o->reward = (-0.5 * c->Q_sym) * (o->e[1] * o->e[1]);                           /* env.py:218 */
    o->rg2 = (-c->Q_sym) * o->e[1];
}

/* test exports: one env step / one bare plant step / the plant's built-in initial state */
void orc_nl_env_step(const orc_nl_cfg* c, double theta_ref_k, const double* act, double* x_full, double* x_act, int32_t stepp,
                     double* surf, double* u, double* e, double* reward, double* x_obs)
{
    orc_nl_envout o;
    orc_nl_env_core(c, theta_ref_k, act, x_full, x_act, stepp, &o);
    memcpy(surf, o.surf, sizeof o.surf); memcpy(u, o.u, sizeof o.u); memcpy(e, o.e, sizeof o.e); *reward = o.reward;
    memcpy(x_obs, o.x_obs, sizeof o.x_obs);
}
void orc_cit_plant_step(const rl4_cit_params* P, double* x, const double* u, double dt, int integrator)
{
    if (integrator == RL4_CIT_INTEGRATOR_RK4) rl4_cit_step_rk4(P, x, u, dt); else rl4_cit_step_ode5(P, x, u, dt);
}

/* the one-step map of the surrogate for n samples (oracle/pe_probe/fit_surrogate.py compares it with the real binary's) */
void orc_cit_plant_step_batch(const rl4_cit_params* P, double* x, const double* u, double dt, int integrator, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) orc_cit_plant_step(P, x + 12 * i, u + 11 * i, dt, integrator);
}
void orc_cit_finalize(rl4_cit_params* P) { rl4_cit_finalize(P); }
void orc_cit_solve_trim(rl4_cit_params* P) { rl4_cit_solve_trim(P); }

#define TE double

#define TN float
#define SFX _mixed
#define N_FMA fmaf
#define N_SQRT sqrtf
#define N_T13 orc_t13_f32
#define N_LIBM_TANH tanhf
#include "nl_oracle_body.inc"
#undef TN
#undef SFX
#undef N_FMA
#undef N_SQRT
#undef N_T13
#undef N_LIBM_TANH

#define TN double
#define SFX _fp64
#define N_FMA fma
#define N_SQRT sqrt
#define N_T13 orc_t13_f64
#define N_LIBM_TANH tanh
#include "nl_oracle_body.inc"

int orc_nl_default_cfg(orc_nl_cfg* c)
{
    if (!c) return -1;
    memset(c, 0, sizeof(*c));
    rl4_cit_default_params(&c->plant);
    const double trim[11] = {-0.02855, 0, 0, 0, 0, 0, 0, 0, 0.55, 0.55, 0};      /* idhp_nonlin.py:53 */
    memcpy(c->trim_input, trim, sizeof(trim));
    c->dt = 0.01;                                                                   /* idhp_nonlin.py:36 */
    c->gamma = 0.6; c->gamma_sq = 0.36; c->tau = 0.02; c->lambda_h = 0.95; c->lambda_l = 0.95; c->lr_decay = 0.998;
    c->eta_a_h = 35.0; c->eta_a_l = 5.0; c->eta_c_h = 1.4; c->eta_c_l = 0.7;       /* idhp_nonlin.py:123-146 */
    c->rls_gamma = 1.0; c->rls_cov0 = 1e6; c->Q_sym = 2.0;
    c->lambda_t = 0.012; c->lambda_s = 0.001;                                       /* objects.py:1383 */
    c->noise_std[0] = 0.010; c->noise_std[1] = 0.010; c->noise_std[2] = 0.008; c->noise_std[3] = 0.003;
    c->omega0 = 13.0; c->omega_slow = 6.0; c->rate_limit = 19.7 * (M_PI / 180.0);
    c->limit_deg[0] = 15.0; c->limit_deg[1] = 37.0; c->limit_deg[2] = 22.0;
    c->damp_factor = 0.3; c->cg_shift = -0.5;
    c->sat_limit[0] = 5.0 * (M_PI / 180.0); c->sat_limit[1] = 18.0 * (M_PI / 180.0); c->sat_limit[2] = 10.0 * (M_PI / 180.0);
    c->multistep = 0; c->warmup_steps = 400; c->cooldown_steps = 200; c->fault_step = -1;
    c->elig_a = 1; c->fault_damp = 0; c->fault_sat = 0; c->integrator = RL4_CIT_INTEGRATOR_ODE5;
    c->flight_step = 5500;
    c->numpy2 = 1;                 /* NEP 50: the mode the verbatim agent was OBSERVED in (numpy 2.3); 0 = numpy 1.x, derived only */
    return 0;
}

int orc_nl_init(int policy, const orc_nl_cfg* cfgs, int cfg_stride, const double* W1a, const double* W2a,
                const double* W1c, const double* W2c, orc_nl_state* st, int64_t n)
{
    if (!cfgs || !W1a || !W2a || !W1c || !W2c || !st || n < 0) return -1;
    for (int64_t i = 0; i < n; ++i) {
        const orc_nl_cfg* c = cfgs + (cfg_stride ? i : 0);
        g_cur_agent = i;
        if (policy == ORC_POLICY_MIXED) nl_init_one_mixed(c, W1a + 40 * i, W2a + 10 * i, W1c + 40 * i, W2c + 30 * i, st + i);
        else if (policy == ORC_POLICY_FP64) nl_init_one_fp64(c, W1a + 40 * i, W2a + 10 * i, W1c + 40 * i, W2c + 30 * i, st + i);
        else return -1;
    }
    return 0;
}

int orc_nl_run(int policy, int tanh_mode, const orc_nl_cfg* cfgs, int cfg_stride, const double* theta_ref,
               const float* noise, int k0, int n_steps, orc_nl_state* st, int64_t n, orc_nl_logrow* log, int64_t n_log)
{
    if (!cfgs || !theta_ref || !noise || !st || n < 0 || n_steps < 0 || k0 < 0) return -1;
    if (policy == ORC_POLICY_MIXED) nl_run_mixed(tanh_mode, cfgs, cfg_stride, theta_ref, noise, k0, n_steps, st, n, log, n_log);
    else if (policy == ORC_POLICY_FP64) nl_run_fp64(tanh_mode, cfgs, cfg_stride, theta_ref, noise, k0, n_steps, st, n, log, n_log);
    else return -1;
    return 0;
}

int orc_nl_sizeof_cfg(void)    { return (int)sizeof(orc_nl_cfg); }
int orc_nl_sizeof_state(void)  { return (int)sizeof(orc_nl_state); }
/* the symmetric-flight variant of the plant step (what the CUDA kernel uses when it applies); returns 0 when the state is not symmetric */
int orc_cit_plant_step_lon(const rl4_cit_params* P, double* x, const double* u, double dt, int integrator)
{
    if (!rl4_cit_is_symmetric(x, u)) return 0;
    rl4_cit_step_auto(P, x, u, dt, integrator);
    return 1;
}
/* element-wise probes of the plant's deterministic elementary functions (tests) */
void orc_cit_sincos(const double* a, double* s, double* c, int64_t n) { for (int64_t i = 0; i < n; ++i) rl4_sincos(a[i], s + i, c + i); }
void orc_cit_air(const rl4_cit_params* P, const double* h, double* rho, double* lapse, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) { const rl4_cit_air a = rl4_cit_airdata(P, h[i]); rho[i] = a.rho; lapse[i] = a.thrust_lapse; }
}
int orc_nl_sizeof_logrow(void) { return (int)sizeof(orc_nl_logrow); }
