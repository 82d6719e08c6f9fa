/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the short-period IDHP hot path.
 *
 * Plain-C restatement of the reference algorithm (wingos80/RL4AFCS):
 *   env step            envs/linear/env.py:156-220
 *   Network/Critic/Actor objects.py:39-281
 *   RLS                  objects.py:439-549
 *   IDHPsp loop          objects.py:551-1004
 * It is NOT part of the product: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.
 *
 * Parity pin: env + RLS are checked bit-for-bit against the verbatim
 * reference classes (tests/test_oracle_vs_reference.py, fixtures in tests/golden/).
 * The TensorFlow parts (actor/critic/update) have no runnable reference here
 * ("parity unpinned" at the TF boundary); they are validated against closed-form /
 * torch.autograd gradients.
 *
 * Arithmetic conventions (re-measured in tests):
 *   every numpy `@` is an in-order FMA chain  s = a0*b0; s = fma(a_t, b_t, s);
 *   every elementwise numpy expression is separately rounded (compile with
 *   -ffp-contract=off);  tanh is the one primitive that is not bit-reproducible
 *   across hosts, so it is selectable (libm | "t13" algorithm shared in WORDS, not in
 *   code, with the CUDA kernel -- see sp_oracle_tanh.h).
 */
#ifndef RL4_SP_ORACLE_H
#define RL4_SP_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_POLICY_FP64 = 0, ORC_POLICY_FP32 = 1, ORC_POLICY_MIXED = 2 };
enum { ORC_ELIG_NONE = 0, ORC_ELIG_ACCUMULATING = 1, ORC_ELIG_REPLACING = 2 };
enum { ORC_TANH_LIBM = 0, ORC_TANH_T13 = 1 };

/* Per-agent configuration (idhp_sp.py:45-52,150-173 + env construction). */
typedef struct {
    double A[4];            /* row-major 2x2, envs/linear/env.py:89-108 (computed by the host in Python) */
    double B[2];            /* envs/linear/env.py:110-119 */
    double A_fault[4];      /* plant after _engage_fault (envs/linear/env.py:127-154) */
    double B_fault[2];
    double dt;
    double gamma;           /* idhp_config['gamma'] */
    double gamma_sq;        /* python `self.gamma**2` (objects.py:887), computed by the host */
    double tau;
    double kappa;
    double lambda_h, lambda_l;
    double eta_a_h, eta_a_l, eta_c_h, eta_c_l;
    double rls_gamma, rls_cov0;
    double error_thresh_deg;
    double ref_amp;         /* reference = ref_amp * ref_base[k]  (idhp_sp.py:44: A*np.sin(...)) */
    int32_t multistep;      /* objects.py:560 */
    int32_t warmup_steps;   /* int(warmup_time/dt)   objects.py:793 */
    int32_t cooldown_steps; /* int(cooldown_time/dt) objects.py:569 */
    int32_t fault_step;     /* int(fault_time/dt), <0 = no fault   envs/linear/env.py:128 */
    int32_t elig_a, elig_c; /* ORC_ELIG_* */
    int32_t q3_alias;       /* x aliasing at k == 1 (reference behaviour = 1) */
    int32_t q7_numpy1;      /* 0 (default) = NEP 50 compare, as observed with the verbatim agent under numpy 2.3;
                             * 1 = numpy-1.x value-based compare of f32 lr vs python float (derived, unverified) */
    int32_t tracked_q;      /* 0: tracked_state 'alpha', 1: 'q'   envs/linear/env.py:180-184 */
    int32_t pad;
} orc_sp_cfg;

/* Loop-carried per-agent state.  All values are stored as doubles; under the fp32 /
 * mixed policies the float-typed members hold exactly representable floats. */
typedef struct {
    double x[2];            /* env.x == agent's x (x_k before step k) */
    double x_prev[2];
    double a, a_prev;       /* normalised actions, network dtype */
    double W1a[4], W2a[4];  /* actor 1-4-1 */
    double W1c[4], W2c[8];  /* critic 1-4-2, W2c row-major (4,2) */
    double W1t[4], W2t[8];  /* target critic */
    double Ea[8];           /* actor trace (1,8)    objects.py:236-254 */
    double Ec[24];          /* critic trace (2,12)  objects.py:161-188 */
    double theta[6];        /* RLS params (3,2) row-major  objects.py:461 */
    double cov[9];          /* RLS Cov (3,3) row-major */
    double cgrad_prev[2];   /* reward_grad of the previous step */
    double M_prev[4];       /* dx1dx0_prev (2,2) row-major */
    double eta_a, eta_c;    /* current SGD learning rates */
    double gl_a, gl_c;      /* gamma_lambda of actor / critic */
    double eps[2];          /* last RLS innovation */
    double eps_norm;
    double sum_c;           /* running sum of rewards (sequential order) */
    double sum_abs_e;       /* running sum |e| (nMAE numerator, an addition of this repo) */
    int32_t cooldown;
    int32_t changed;        /* RLS one-shot reset done  objects.py:838-841 */
    int32_t lr_init;        /* 1 until the first lr switch (eta still a python float, Q7) */
    int32_t diverged_step;  /* step at which isnan(c) broke the loop, -1 otherwise */
    int32_t conv_step;      /* last k with alpha error > 0.5 deg (utils.py:350-369), -1 if none */
    int32_t x_nan;          /* any NaN ever logged in x_hist (functions.py:162) */
} orc_sp_state;

/* One row of IDHPsp._log (objects.py:651-726). */
typedef struct {
    double t;
    double x[2];
    double a;
    double s[2];
    double c;
    double ref;
    double a_w1[4], a_w2[4], c_w1[4], c_w2[8];
    double a_e[8], c_e[24];
    double a_all_grad[8], c_all_grad[12];
    double a_grad_norm, c_grad_norm;
    double params[6], cov[9];
    double eps_norm, eps_abs[2];
    /* extras (not in the reference log, used by the parity tests) */
    double e, lam[2], lam_t[2], td[2], dadz, M[4], loss_grad;
} orc_sp_logrow;

/* Run steps [k0, k0+n_steps) for n_agents agents.
 *   cfgs      : n_agents configs if cfg_stride == 1, one shared config if cfg_stride == 0
 *   ref_base  : table of at least k0+n_steps reference samples
 *   states    : in/out
 *   log       : NULL or n_log_agents * n_steps rows (agents [0, n_log_agents), agent-major)
 * Returns 0, or -1 on bad arguments.  */
int orc_sp_run(int policy, int tanh_mode, const orc_sp_cfg* cfgs, int cfg_stride,
               const double* ref_base, int k0, int n_steps,
               orc_sp_state* states, int64_t n_agents,
               orc_sp_logrow* log, int64_t n_log_agents);

/* IDHPsp.__init__ + train() prologue (objects.py:921-948): builds the loop-entry state
 * from x0 and the initial weights (target <- critic, a0 = actor(0) = 0 BEFORE the sign
 * flip of W2a, RLS reset, zero traces). */
int orc_sp_init(int policy, const orc_sp_cfg* cfgs, int cfg_stride,
                const double* x0 /* [n][2] */, const double* W1a, const double* W2a,
                const double* W1c, const double* W2c /* [n][4],[n][4],[n][4],[n][8] */,
                orc_sp_state* states, int64_t n_agents);

/* Step-API forms of envs/linear/env.py:156-220 and objects.py:492-543 for n agents (AoS:
 * x [n][2], theta [n][6], cov [n][9], dx0/dx1/eps [n][2]); action_deg is the caller's 20*a. */
int orc_sp_env_step(int policy, const orc_sp_cfg* cfgs, int cfg_stride, const double* ref_base, int stepp,
                    double* x, const double* action_deg, double* reward, double* e, double* rg0, int64_t n);
int orc_sp_rls_update(int policy, const orc_sp_cfg* cfgs, int cfg_stride, double* theta, double* cov, const double* dx0,
                      const double* da0, const double* dx1, double* eps, double* eps_norm, int64_t n);

/* the oracle's tanh variants, exposed for accuracy tests */
void orc_tanh_t13_f64_array(const double* x, double* y, int64_t n);
void orc_tanh_t13_f32_array(const float* x, float* y, int64_t n);
double orc_tanh_t13_f64(double x);
float  orc_tanh_t13_f32(float x);

int orc_sizeof_cfg(void);
int orc_sizeof_state(void);
int orc_sizeof_logrow(void);

#ifdef __cplusplus
}
#endif
#endif
