/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the nonlinear IDHP path.
 *
 * Plain-C restatement of the reference algorithm (wingos80/RL4AFCS):
 *   env wrapper            envs/nonlinear/env.py:60-311 (actuators, faults, reward, MDP state)
 *   Network / Critic_big / Actor_big     objects.py:39-140, 283-437
 *   RLS (n = 3, m = 1)     objects.py:439-549
 *   IDHPnonlin loop        objects.py:1006-1564 (_step_networks :1292, _update_networks :1350, _adapt_check :1212)
 * around the documented surrogate plant of include/rl4_citation_surrogate.h (the reference's plant is a
 * source-less Windows binary: with the surrogate header plant parity is calibrated, not pinned; `orc_nl_set_external_plant` runs
 * the wrapper + agent on the translated binary itself (oracle/pe_probe); TensorFlow parts unpinned as in sp_oracle.h).
 * Two policies: mixed (TF float32 nets + numpy float64 env/RLS, the reference's mix) and fp64.
 * numpy-side `@` orders measured in the build container (oracle/make_golden.py header):
 *   (4,3).T@(4,1) -> fma(a0,b0,a1*b1) + fma(a2,b2,a3*b3);  (4,4)@(4,1) -> (p0+p2)+(p1+p3), products rounded;
 *   (1,4)@(4,1), (1,3)@(3,3), ddot -> in-order FMA chain.
 */
#ifndef RL4_NL_ORACLE_H
#define RL4_NL_ORACLE_H
#include <stdint.h>
#include "../include/rl4_citation_surrogate.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_NL_DAMP_NONE = 0, ORC_NL_DAMP_ELEVATOR, ORC_NL_DAMP_AILERON, ORC_NL_DAMP_RUDDER, ORC_NL_DAMP_ALL,
       ORC_NL_SHIFT_CG, ORC_NL_SLOW_ALL };
enum { ORC_NL_SAT_NONE = 0, ORC_NL_SAT_ELEVATOR, ORC_NL_SAT_AILERON, ORC_NL_SAT_RUDDER };

typedef struct {
    rl4_cit_params plant;
    double trim_input[11];      /* idhp_nonlin.py:53 */
    double dt;
    double gamma, gamma_sq, tau, lambda_h, lambda_l, lr_decay;
    double eta_a_h, eta_a_l, eta_c_h, eta_c_l;
    double rls_gamma, rls_cov0;
    double Q_sym;               /* kappa[1], envs/nonlinear/env.py:64 */
    double lambda_t, lambda_s;  /* objects.py:1383 */
    double noise_std[4];        /* objects.py:1377 */
    double omega0, omega_slow;  /* 13 and 6 rad/s, envs/nonlinear/env.py:74,145 */
    double rate_limit;          /* deg2rad(19.7), envs/nonlinear/env.py:177 */
    double limit_deg[3];        /* 15, 37, 22 (symmetric), envs/nonlinear/env.py:107-109 */
    double damp_factor;         /* 0.3 */
    double cg_shift;            /* -0.5 */
    double sat_limit[3];        /* deg2rad(5), deg2rad(18), deg2rad(10), envs/nonlinear/env.py:154-158 */
    int32_t multistep, warmup_steps, cooldown_steps, fault_step;
    int32_t elig_a;             /* 0 none, 1 accumulating, 2 replacing */
    int32_t fault_damp, fault_sat, integrator;
    int32_t flight_step;        /* 5500: RSE split of functions.py:916-917,1038-1039 */
    int32_t numpy2;             /* 1 (default): NEP 50 (numpy >= 2) -- what the verbatim code was OBSERVED to do in this container,
                                 * the mode every golden fixture was generated in; 0: numpy 1.x value-based promotion in
                                 * _adapt_check, derived from the promotion rules, never observed (opt-in, unverified) */
} orc_nl_cfg;

typedef struct {
    double x_full[12];          /* plant state */
    double x_act[3];            /* actuator states, envs/nonlinear/env.py:71 */
    double s[4], s_prev[4];     /* MDP states (network dtype) */
    double a, a_prev;
    double x_lon[3], x_prev_lon[3];
    double W1a[40], W2a[10], W1c[40], W2c[30], W1t[40], W2t[30];
    double Ea[50];
    double theta[12], cov[16];
    double cgrad_prev[3];
    double M_prev[9];
    double eta_a, eta_c, lambdaa;   /* self.eta_a / eta_c / lambdaa (objects.py:1252-1263) */
    double lr_a, lr_c;              /* optimizer learning rates */
    double gl;                      /* gamma_lambda of actor / critic */
    double eps[3], eps_norm;
    double rse[2];                  /* cumulative RSE, objects.py:1503-1504 */
    double nz_peak;                 /* max |V*q/9.80665|, functions.py:774,1050,1055 */
    double rse_flight[2];           /* RSE over steps >= flight_step */
    int32_t cooldown, diverged_step, stepp;
    int32_t pyfloat_mask;           /* numpy2 mode: bit 0/1/2 set while eta_a / eta_c / lambdaa is still a python float */
} orc_nl_state;

/* per-step record for tests */
typedef struct {
    double x_full[12], s_next[4], a_next, reward, e_theta, lam[3], lam_t[3], td[3], dads[4], M[9], loss_grad, a_random;
    double surf[3], model_input[11];
    /* the quantities IDHPnonlin._log keeps (objects.py:1119-1165), as of the end of the step */
    double eta_a, rse_step[2], yref_theta, W1a[40], W2a[10], W1c[40], W2c[30], a_grad[50], c_grad[70], theta[12], cov[16],
           eps[3], eps_norm, wa_norm, wc_norm;
} orc_nl_logrow;

int orc_nl_default_cfg(orc_nl_cfg* c);
/* reset (envs/nonlinear/env.py:258-311: 1000 + 1 plant steps at trim input) + train() prologue (objects.py:1466-1488) */
int orc_nl_init(int policy, const orc_nl_cfg* cfgs, int cfg_stride, const double* W1a, const double* W2a,
                const double* W1c, const double* W2c, orc_nl_state* st, int64_t n);
/* steps [k0, k0+n_steps); theta_ref table of >= k0+n_steps samples (phi / psi references are zero as in
 * idhp_nonlin.py:116-117); noise[(k-k0)*n + i] = the N(0,1) draw of agent i at step k (objects.py:1375) */
int orc_nl_run(int policy, int tanh_mode, const orc_nl_cfg* cfgs, int cfg_stride, const double* theta_ref,
               const float* noise, int k0, int n_steps, orc_nl_state* st, int64_t n,
               orc_nl_logrow* log, int64_t n_log);
void orc_nl_rls_update(double gamma, double* theta, double* cov, const double* dx0, const double* da0, const double* dx1,
                       double* eps, double* eps_norm, int64_t n);
void orc_nl_env_step(const orc_nl_cfg* c, double theta_ref_k, const double* act, double* x_full, double* x_act, int32_t stepp,
                     double* surf, double* u, double* e, double* reward, double* x_obs);
void orc_cit_plant_step(const rl4_cit_params* P, double* x, const double* u, double dt, int integrator);
int orc_cit_plant_step_lon(const rl4_cit_params* P, double* x, const double* u, double dt, int integrator);
void orc_cit_sincos(const double* a, double* s, double* c, int64_t n);
void orc_cit_air(const rl4_cit_params* P, const double* h, double* rho, double* lapse, int64_t n);
int orc_nl_sizeof_cfg(void);
int orc_nl_sizeof_state(void);
int orc_nl_sizeof_logrow(void);

#ifdef __cplusplus
}
#endif
#endif
