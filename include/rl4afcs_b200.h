/* rl4afcs_b200 -- C ABI of the B200-native batched IDHP flight-control engine.
 *
 * The reference (wingos80/RL4AFCS) has no FFI on this path: it is a Python object API
 *   envs/linear/env.py:156,222        Ce500ShortPeriod.step / reset
 *   objects.py:142-281                Critic / Actor (call, get_weight_update, soft_update)
 *   objects.py:439-549                RLS (update, F, G, _reset)
 *   objects.py:551-1004               IDHPsp (train and its helpers)
 * so the drop-in boundary is the Python package `rl4afcs_b200` (same class / method /
 * attribute names with a leading batch dimension), and THIS header is the thin native
 * layer those classes call through ctypes.  Each entry point cites the reference
 * function(s) it replaces.  INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions
 *   - caller owns every buffer; nothing here allocates except rl4_sp_episode_host()'s
 *     context (rl4_ctx_create / rl4_ctx_destroy);
 *   - device buffers are structure-of-arrays planes  plane[field * stride + agent],
 *     agent index fastest, `stride` >= n_agents;
 *   - dtype policy: RL4_FP64 (all double), RL4_FP32 (all float), RL4_MIXED (the reference's
 *     own mix: network plane float, env/RLS/trace plane double);
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), returns
 *     0 on success, <0 for an argument error, >0 = cudaError_t; rl4_last_error() gives text;
 *   - sm_100a only.  There is no CPU fallback: without a B200 the calls fail.
 */
#ifndef RL4AFCS_B200_H
#define RL4AFCS_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RL4_ABI_VERSION 3      /* bumped whenever a struct, enum or signature below changes; _lib.load() checks it */

enum rl4_policy { RL4_FP64 = 0, RL4_FP32 = 1, RL4_MIXED = 2 };
enum rl4_elig   { RL4_ELIG_NONE = 0, RL4_ELIG_ACCUMULATING = 1, RL4_ELIG_REPLACING = 2 };
/* plant variants selectable per agent (envs/linear/env.py:127-154) */
enum rl4_fault  { RL4_FAULT_NONE = 0, RL4_FAULT_INVERT_ELEVATOR = 1, RL4_FAULT_DAMP_ELEVATOR = 2,
                  RL4_FAULT_SHIFT_CG = 3, RL4_FAULT_COUNT = 4 };

/* ---- state planes of the short-period agent (objects.py:551-1004 loop-carried values) ---- */
/* env plane (dtype TE): plant state, RLS model, traces, statistics */
enum rl4_sp_env_field {
    RL4_SPE_X = 0,          /* [2] env.x == agent x_k            envs/linear/env.py:43,193 */
    RL4_SPE_XPREV = 2,      /* [2] x_{k-1}                        objects.py:980 */
    RL4_SPE_THETA = 4,      /* [6] RLS params (3,2) row-major     objects.py:461 */
    RL4_SPE_COV = 10,       /* [9] RLS Cov (3,3) row-major        objects.py:470 */
    RL4_SPE_CGRAD_PREV = 19,/* [1] reward_grad[0] of step k-1     objects.py:984 */
    RL4_SPE_EPS = 20,       /* [2] last RLS innovation            objects.py:537 */
    RL4_SPE_EPS_NORM = 22,  /* [1]                                objects.py:539 */
    RL4_SPE_SUM_C = 23,     /* [1] running sum of rewards         functions.py:53 */
    RL4_SPE_SUM_ABS_E = 24, /* [1] running sum |e| (nMAE numerator; addition of this repo) */
    RL4_SPE_EA = 25,        /* [8] actor trace E (1,8)            objects.py:236-254 */
    RL4_SPE_EC_H = 33,      /* [4] critic trace E[0,0:4] == E[1,4:8]   objects.py:161-188 */
    RL4_SPE_EC_W1R0 = 37,   /* [4] critic trace E[0,8:12] */
    RL4_SPE_EC_W1R1 = 41,   /* [4] critic trace E[1,8:12] */
    RL4_SPE_COUNT = 45
};
/* net plane (dtype TN): actions, weights, learning rates */
enum rl4_sp_net_field {
    RL4_SPN_A = 0,          /* [1] a_k (normalised)               objects.py:933,981 */
    RL4_SPN_APREV = 1,      /* [1] */
    RL4_SPN_W1A = 2,        /* [4] actor W1 (1,4)                 objects.py:590 */
    RL4_SPN_W2A = 6,        /* [4] actor W2 (4,1) */
    RL4_SPN_W1C = 10,       /* [4] critic W1 (1,4)                objects.py:591 */
    RL4_SPN_W2C = 14,       /* [8] critic W2 (4,2) row-major */
    RL4_SPN_W1T = 22,       /* [4] target critic                  objects.py:592 */
    RL4_SPN_W2T = 26,       /* [8] */
    RL4_SPN_MPREV = 34,     /* [4] dx1dx0_prev (2,2) row-major    objects.py:985 */
    RL4_SPN_ETA_A = 38,     /* [1] actor SGD learning rate        objects.py:918 */
    RL4_SPN_ETA_C = 39,     /* [1] critic SGD learning rate       objects.py:919 */
    RL4_SPN_COUNT = 40
};
/* int plane (int32) */
enum rl4_sp_int_field {
    RL4_SPI_COOLDOWN = 0,   /* objects.py:568,809-835 */
    RL4_SPI_FLAGS = 1,      /* RL4_SPF_* */
    RL4_SPI_DIVERGED_STEP = 2, /* step whose reward was NaN (objects.py:991), -1 otherwise */
    RL4_SPI_CONV_STEP = 3,  /* last step with |alpha error| > 0.5 deg (utils.py:350-369), -1 if none */
    RL4_SPI_COUNT = 4
};
enum rl4_sp_flag {
    RL4_SPF_CHANGED = 1,    /* one-shot RLS reset done            objects.py:838-841 */
    RL4_SPF_LR_INIT = 2,    /* eta still the python float */
    RL4_SPF_LAMBDA_LOW = 4, /* gamma_lambda currently = lambda_l*gamma  objects.py:814-832 */
    RL4_SPF_X_NAN = 8       /* a NaN state was logged             functions.py:162 */
};

typedef struct rl4_sp_state {
    void*    env;           /* [RL4_SPE_COUNT][stride] of TE */
    void*    net;           /* [RL4_SPN_COUNT][stride] of TN */
    int32_t* ints;          /* [RL4_SPI_COUNT][stride] */
    int64_t  stride;
} rl4_sp_state;

/* per-agent hyper-parameter overrides (device pointers, NULL = use the shared scalar) */
enum rl4_sp_hp {
    RL4_HP_ETA_A_H = 0, RL4_HP_ETA_A_L, RL4_HP_ETA_C_H, RL4_HP_ETA_C_L,
    RL4_HP_LAMBDA_H, RL4_HP_LAMBDA_L, RL4_HP_GAMMA, RL4_HP_GAMMA_SQ, RL4_HP_TAU, RL4_HP_KAPPA,
    RL4_HP_RLS_GAMMA, RL4_HP_RLS_COV0, RL4_HP_ERROR_THRESH_DEG, RL4_HP_REF_AMP,
    RL4_HP_COUNT
};
enum rl4_sp_hpi {
    RL4_HPI_MULTISTEP = 0, RL4_HPI_WARMUP_STEPS, RL4_HPI_COOLDOWN_STEPS,
    RL4_HPI_FAULT_STEP, RL4_HPI_FAULT_KIND, RL4_HPI_ELIG_A, RL4_HPI_ELIG_C,
    RL4_HPI_TRACKED_Q,          /* 0: tracked_state 'alpha' (idhp_sp.py:175), 1: 'q' (envs/linear/env.py:180-184); the reward
                                 * gradient stays in the alpha slot either way (Q4) */
    RL4_HPI_COUNT
};

/* Shared configuration: idhp_sp.py:45-52,150-173 and the plant of envs/linear/env.py:66-154.
 * The plant matrices are computed by the host exactly as the reference does (Python floats). */
typedef struct rl4_sp_params {
    double A[RL4_FAULT_COUNT][4];   /* row-major 2x2 per plant variant; variant 0 = nominal */
    double B[RL4_FAULT_COUNT][2];
    double dt;
    double hp[RL4_HP_COUNT];        /* shared scalars, indexed by rl4_sp_hp */
    int32_t hpi[RL4_HPI_COUNT];     /* shared ints, indexed by rl4_sp_hpi (fault_step < 0: no fault) */
    int32_t q3_alias;               /* reproduce the x aliasing at k == 1 (reference: 1) */
    int32_t q7_numpy1;              /* 0 (default) = NEP 50 float32-vs-python-float compare, the behaviour OBSERVED when the
                                     * verbatim agent runs under numpy >= 2; 1 = numpy-1.x value-based compare (derived from the promotion
                                     * rules, never observed: opt-in, unverified) */
    const double*  hp_agent[RL4_HP_COUNT];    /* optional per-agent overrides, length n_agents */
    const int32_t* hpi_agent[RL4_HPI_COUNT];
} rl4_sp_params;

/* Optional trajectory log: rows for agents [0, n_agents_logged), steps k with
 * (k - k0) % every == 0, layout  buf[((row * n_fields) + field) * n_agents_logged + agent]. */
enum rl4_sp_log_level { RL4_LOG_NONE = 0, RL4_LOG_BASIC = 1, RL4_LOG_FULL = 2 };
enum rl4_sp_log_basic_field {   /* IDHPsp._log objects.py:682-688 */
    RL4_LB_X = 0 /* [2] */, RL4_LB_A = 2, RL4_LB_C = 3, RL4_LB_REF = 4, RL4_LB_E = 5, RL4_LB_COUNT = 6
};
enum rl4_sp_log_full_field {    /* the rest of objects.py:691-726 */
    RL4_LF_AW1 = 6 /* [4] */, RL4_LF_AW2 = 10 /* [4] */, RL4_LF_CW1 = 14 /* [4] */, RL4_LF_CW2 = 18 /* [8] */,
    RL4_LF_AE = 26 /* [8] */, RL4_LF_CE = 34 /* [12] h, w1 row0, w1 row1 */,
    RL4_LF_AGRAD = 46 /* [8] a_all_grad */, RL4_LF_CGRAD = 54 /* [12] c_all_grad */,
    RL4_LF_PARAMS = 66 /* [6] */, RL4_LF_COV = 72 /* [9] */, RL4_LF_EPS_NORM = 81, RL4_LF_EPS_ABS = 82 /* [2] */,
    RL4_LF_LAM = 84 /* [2] */, RL4_LF_LAM_T = 86 /* [2] */, RL4_LF_TD = 88 /* [2] */, RL4_LF_DADZ = 90,
    RL4_LF_M = 91 /* [4] */, RL4_LF_LOSS_GRAD = 95, RL4_LF_COUNT = 96
};
typedef struct rl4_sp_log {
    double* buf;                /* device, may be NULL when level == RL4_LOG_NONE */
    int32_t level;
    int32_t every;              /* >= 1 */
    int64_t n_agents_logged;
} rl4_sp_log;

/* ---- library ---- */
int         rl4_abi_version(void);
/* sizeof(rl4_sp_params) / sizeof(rl4_nl_params) as this library was compiled: the binding compares them with its own
 * struct layouts before the first by-value call */
int64_t     rl4_sizeof_sp_params(void);
int64_t     rl4_sizeof_nl_params(void);
const char* rl4_last_error(void);
/* 0 if `device` is an sm_100 GPU this library can run on */
int         rl4_device_check(int device);

/* ---- short-period path ---- */

/* IDHPsp.__init__/_setup_networks/_setup_optimizers/_invert_controller + train() prologue
 * (objects.py:552-615,843-851,911-948) and Ce500ShortPeriod.reset (envs/linear/env.py:222-258):
 * builds the loop-entry state from x0 [2][stride_in] and initial weights W1a[4], W2a[4],
 * W1c[4], W2c[8] (each [n][stride_in] planes of double, rounded to the policy's dtypes). */
int rl4_sp_init(int policy, const rl4_sp_params* p, const double* x0, const double* w1a, const double* w2a,
                const double* w1c, const double* w2c, int64_t stride_in,
                rl4_sp_state st, int64_t n_agents, void* stream);

/* IDHPsp.train() hot loop (objects.py:950-1004), steps [k0, k0 + n_steps) for every agent, fused
 * with env.step (envs/linear/env.py:156-220), Critic/Actor forward + traces (objects.py:151-193,
 * 226-259), _update_networks (:884-908), RLS.update (:492-543), _adapt_check (:783-841) and the
 * episode statistics of MC_run_seed (functions.py:39-60).  `ref_base` is a device table with at
 * least k0 + n_steps samples; the reference signal is hp[REF_AMP] * ref_base[k] (idhp_sp.py:44,174).
 * `use_traces` = 0 requires elig_a == elig_c == RL4_ELIG_NONE for every agent. */
int rl4_sp_run(int policy, const rl4_sp_params* p, const double* ref_base, int32_t k0, int32_t n_steps,
               rl4_sp_state st, int64_t n_agents, int32_t use_traces, rl4_sp_log log, void* stream);

/* Ce500ShortPeriod.step (envs/linear/env.py:156-220) for every agent: step-API form.
 *   x [2][stride] (TE, in/out), action_deg [stride] (TN; the caller's 20*a, degrees),
 *   out_reward, out_e, out_reward_grad0 [stride] (TE).  stepp = the env's step counter. */
int rl4_sp_env_step(int policy, const rl4_sp_params* p, const double* ref_base, int32_t stepp,
                    void* x, const void* action_deg, void* out_reward, void* out_e, void* out_reward_grad0,
                    int64_t stride, int64_t n_agents, void* stream);

/* RLS.update (objects.py:492-543): theta [6][stride], cov [9][stride] in/out (TE);
 * dx0 [2][stride], da0 [stride], dx1 [2][stride] (TE); out eps [2][stride], eps_norm [stride]. */
int rl4_sp_rls_update(int policy, const rl4_sp_params* p, void* theta, void* cov,
                      const void* dx0, const void* da0, const void* dx1, void* out_eps, void* out_eps_norm,
                      int64_t stride, int64_t n_agents, void* stream);

/* Critic.call (objects.py:151-193) / Actor.call (:226-259): forward + trace for every agent.
 *   z [stride] (TN), w1 [4][stride], w2 [4*n_out][stride] (TN), E planes (TE, in/out),
 *   out [n_out][stride] (TN), gamma_lambda shared, elig mode shared.
 *   critic: n_out = 2, E = [12][stride] (h, w1 row0, w1 row1);  actor: n_out = 1, E = [8][stride],
 *   out_dadz [stride] (TN) may be NULL (objects.py:876-878). */
int rl4_sp_critic_forward(int policy, const void* z, const void* w1, const void* w2, void* E, void* out_lambda,
                          double gamma_lambda, int32_t elig, int64_t stride, int64_t n_agents, void* stream);
int rl4_sp_actor_forward(int policy, const void* z, const void* w1, const void* w2, void* E, void* out_a,
                         void* out_dadz, double gamma_lambda, int32_t elig, int64_t stride, int64_t n_agents,
                         void* stream);

/* Critic.get_weight_update (objects.py:195-205): td [2][stride] (TN) times the critic trace E
 * [12][stride] (TE, cast to TN as TF does); out [12][stride] (TN) = W1_update (1,4) then W2_update
 * (4,2) row-major.  (Actor.get_weight_update is a plain product loss*E and needs no kernel.) */
int rl4_sp_critic_weight_update(int policy, const void* td, const void* E, void* out, int64_t stride,
                                int64_t n_agents, void* stream);


/* =====================================================================================
 * Nonlinear path: Ce500NonLinear (envs/nonlinear/env.py:11-319) + IDHPnonlin (objects.py:1006-1564)
 * around the documented surrogate plant of include/rl4_citation_surrogate.h (the reference's plant
 * is a source-less binary; plant parity unpinned).  Policies: RL4_MIXED and RL4_FP64.
 * ===================================================================================== */
#include "rl4_citation_surrogate.h"

enum rl4_nl_env_field {        /* env plane, double */
    RL4_NLE_XFULL = 0,         /* [12] plant state p q r V alpha beta phi theta psi h xe ye   env.py:22-26 */
    RL4_NLE_XACT = 12,         /* [3] actuator states                                           env.py:71 */
    RL4_NLE_XLON = 15,         /* [3] x_lon = [alpha theta q]                                   objects.py:1476 */
    RL4_NLE_XPREVLON = 18,     /* [3] */
    RL4_NLE_THETA = 21,        /* [12] RLS params (4,3) row-major                               objects.py:461 */
    RL4_NLE_COV = 33,          /* [16] RLS Cov (4,4) */
    RL4_NLE_CGRAD_PREV = 49,   /* [1] reward_grad_lon[2] of the previous step                   objects.py:1537 */
    RL4_NLE_EPS = 50,          /* [3] */
    RL4_NLE_EPS_NORM = 53,
    RL4_NLE_RSE = 54,          /* [2] cumulative RSE                                            objects.py:1503-1504 */
    RL4_NLE_NZ_PEAK = 56,      /* max |V q / g0|                                                functions.py:774,1050,1055 */
    RL4_NLE_ETA_A = 57,        /* self.eta_a / eta_c / lambdaa / gamma_lambda                   objects.py:1252-1281 */
    RL4_NLE_ETA_C = 58,
    RL4_NLE_LAMBDAA = 59,
    RL4_NLE_GL = 60,
    RL4_NLE_EA = 61,           /* [50] actor trace E (1,50)                                     objects.py:385-392 */
    RL4_NLE_RSE_FLIGHT = 111,  /* [2] RSE summed over steps >= RL4_NHPI_FLIGHT_STEP only            functions.py:916-917,1038-1039 */
    RL4_NLE_COUNT = 113
};
enum rl4_nl_net_field {        /* net plane, TN */
    RL4_NLN_S = 0,             /* [4] MDP state s                                               objects.py:1474 */
    RL4_NLN_SPREV = 4,         /* [4] */
    RL4_NLN_A = 8, RL4_NLN_APREV = 9,
    RL4_NLN_W1A = 10,          /* [40] actor W1 (4,10) row-major */
    RL4_NLN_W2A = 50,          /* [10] */
    RL4_NLN_W1C = 60,          /* [40] */
    RL4_NLN_W2C = 100,         /* [30] critic W2 (10,3) row-major */
    RL4_NLN_W1T = 130,         /* [40] */
    RL4_NLN_W2T = 170,         /* [30] */
    RL4_NLN_MPREV = 200,       /* [9] dx1dx0_prev (3,3) */
    RL4_NLN_LR_A = 209, RL4_NLN_LR_C = 210,   /* optimizer learning rates */
    RL4_NLN_COUNT = 211
};
enum rl4_nl_int_field { RL4_NLI_COOLDOWN = 0, RL4_NLI_DIVERGED_STEP = 1, RL4_NLI_STEPP = 2,
                        RL4_NLI_PYFLOAT_MASK = 3,   /* NUMPY2 mode: bit 0/1/2 set while eta_a / eta_c / lambdaa is a python float */
                        RL4_NLI_COUNT = 4 };

typedef struct rl4_nl_state {
    double*  env;            /* [RL4_NLE_COUNT][stride] */
    void*    net;            /* [RL4_NLN_COUNT][stride] of TN */
    int32_t* ints;           /* [RL4_NLI_COUNT][stride] */
    int64_t  stride;
} rl4_nl_state;

enum rl4_nl_hp {
    RL4_NHP_ETA_A_H = 0, RL4_NHP_ETA_A_L, RL4_NHP_ETA_C_H, RL4_NHP_ETA_C_L, RL4_NHP_LAMBDA_H, RL4_NHP_LAMBDA_L,
    RL4_NHP_GAMMA, RL4_NHP_GAMMA_SQ, RL4_NHP_TAU, RL4_NHP_LR_DECAY, RL4_NHP_RLS_GAMMA, RL4_NHP_RLS_COV0,
    RL4_NHP_Q_SYM, RL4_NHP_LAMBDA_T, RL4_NHP_LAMBDA_S, RL4_NHP_DAMP_FACTOR, RL4_NHP_CG_SHIFT, RL4_NHP_COUNT
};
enum rl4_nl_hpi {
    RL4_NHPI_MULTISTEP = 0, RL4_NHPI_WARMUP_STEPS, RL4_NHPI_COOLDOWN_STEPS, RL4_NHPI_FAULT_STEP,
    RL4_NHPI_FAULT_DAMP, RL4_NHPI_FAULT_SAT, RL4_NHPI_ELIG_A,
    RL4_NHPI_FLIGHT_STEP,      /* first step of the 'flight' phase of the RSE split (5500)      functions.py:916-917 */
    RL4_NHPI_NUMPY2,           /* _adapt_check promotion rules (objects.py:1235-1284): 1 (default) = NEP 50 / numpy >= 2 (python
                                * floats are weak: float32 arithmetic once eta / lambda have become float32 arrays) -- the mode
                                * observed against the verbatim agent; 0 = numpy 1.x value-based casting (float64 intermediates),
                                * derived from the promotion rules, never observed: opt-in, unverified */
    RL4_NHPI_COUNT
};
enum rl4_nl_fault_damp { RL4_NL_DAMP_NONE = 0, RL4_NL_DAMP_ELEVATOR, RL4_NL_DAMP_AILERON, RL4_NL_DAMP_RUDDER,
                         RL4_NL_DAMP_ALL, RL4_NL_SHIFT_CG, RL4_NL_SLOW_ALL };
enum rl4_nl_fault_sat  { RL4_NL_SAT_NONE = 0, RL4_NL_SAT_ELEVATOR, RL4_NL_SAT_AILERON, RL4_NL_SAT_RUDDER };

typedef struct rl4_nl_params {
    rl4_cit_params plant;
    double trim_input[11];          /* idhp_nonlin.py:53 */
    double dt;
    double hp[RL4_NHP_COUNT];
    double noise_std[4];            /* objects.py:1377 */
    double omega0, omega_slow;      /* envs/nonlinear/env.py:74,145 */
    double rate_limit;              /* deg2rad(19.7) */
    double limit_deg[3];            /* 15, 37, 22 */
    double sat_limit[3];            /* deg2rad(5, 18, 10) */
    int32_t hpi[RL4_NHPI_COUNT];
    int32_t integrator;             /* RL4_CIT_INTEGRATOR_* */
    const double*  hp_agent[RL4_NHP_COUNT];
    const int32_t* hpi_agent[RL4_NHPI_COUNT];
} rl4_nl_params;

/* log.level 1: compact row */
enum rl4_nl_log_field { RL4_NLL_XFULL = 0 /* [12] */, RL4_NLL_A = 12, RL4_NLL_E_THETA = 13, RL4_NLL_REWARD = 14,
                        RL4_NLL_SURF = 15 /* [3] action_commanded */, RL4_NLL_COUNT = 18 };
/* log.level 2: every quantity IDHPnonlin._log keeps (objects.py:1083-1176), values as of the END of step k
 * (after the weight / RLS / learning-rate updates, like the reference's call site objects.py:1541).  'a_elig' and
 * 'c_elig' of the reference are allocated but never written (all zero) and are not carried; 't' is dt (k + 1). */
enum rl4_nl_fulllog_field {
    RL4_NLF_ETA_A = 0,           /* log['eta_a']                                                  */
    RL4_NLF_XFULL = 1,           /* [12] log['x_full']                                            */
    RL4_NLF_RSE = 13,            /* [2] per-step RSE (env.py:251)                                 */
    RL4_NLF_X = 15,              /* [3] x_lon = x_full[4, 7, 1]                                   */
    RL4_NLF_A_CMD = 18,          /* action_commanded[0]: elevator position after saturation [rad] */
    RL4_NLF_A_EFF = 19,          /* action_effective[0]: model_input[0] = trim + effective        */
    RL4_NLF_S = 20,              /* info['s'][0] = alpha (the reference broadcasts it over the 4 columns of log['s']) */
    RL4_NLF_YREF = 21,           /* info['yref'][1] = theta_ref (broadcast the same way)          */
    RL4_NLF_E = 22,              /* info['e'][1] = theta error                                    */
    RL4_NLF_A_W1 = 23,           /* [40] actor W1 (4,10) row-major                                */
    RL4_NLF_A_W2 = 63,           /* [10]                                                          */
    RL4_NLF_C_W1 = 73,           /* [40]                                                          */
    RL4_NLF_C_W2 = 113,          /* [30] (10,3) row-major                                         */
    RL4_NLF_A_GRAD = 143,        /* [50] actor_loss_grad: W1 (4,10) then W2 (10,1); zero at k = 0 */
    RL4_NLF_C_GRAD = 193,        /* [70] critic_loss_grad: W1 (4,10) then W2 (10,3); zero at k = 0 */
    RL4_NLF_RLS_PARAMS = 263,    /* [12] (4,3) row-major                                          */
    RL4_NLF_RLS_COV = 275,       /* [16]                                                          */
    RL4_NLF_RLS_EPS = 291,       /* [3]                                                           */
    RL4_NLF_RLS_EPS_NORM = 294,
    RL4_NLF_A = 295,             /* a_next (not in the reference's log; convenience)              */
    RL4_NLF_REWARD = 296,        /* reward_lon                                                    */
    RL4_NLF_COUNT = 297
};

/* log.level 3: the per-step quantities MC_test_hparam keeps per repetition (functions.py:1040-1052), light enough
 * to log EVERY agent of a sweep (88 B per agent-step); n_z follows as V q / g0. */
enum rl4_nl_mclog_field {
    RL4_NLM_E = 0,               /* theta error [rad]                 */
    RL4_NLM_THETA = 1, RL4_NLM_ALPHA = 2, RL4_NLM_Q = 3, RL4_NLM_V = 4, RL4_NLM_H = 5,   /* x_full[7, 4, 1, 3, 9] */
    RL4_NLM_A_CMD = 6, RL4_NLM_A_EFF = 7,                                                 /* as in level 2         */
    RL4_NLM_WA_NORM = 8,         /* ||actor W1||_2 (functions.py:1028) */
    RL4_NLM_WC_NORM = 9,         /* ||critic W1||_2 (functions.py:1030) */
    RL4_NLM_RLS_EPS = 10,        /* rls eps_norm                       */
    RL4_NLM_COUNT = 11
};

/* Fills *p with the configuration of idhp_nonlin.py:36-54,107-146 and the default surrogate plant (host only). */
int rl4_nl_default_params(rl4_nl_params* p);
/* Host only: the plant state Ce500NonLinear.reset leaves behind (envs/nonlinear/env.py:278-291: initialize, 1000 + 1
 * model.step calls at the trim input).  out_state [12] = the state the plant carries into the episode (1001 integrations),
 * out_observed [12] = what the last model.step call returned = env.state after reset (the plant returns the state BEFORE
 * each step: 1000 integrations).  Either may be NULL. */
int rl4_nl_trim_state(const rl4_nl_params* p, double* out_state, double* out_observed);
/* Ce500NonLinear.reset (envs/nonlinear/env.py:258-311: 1000 + 1 plant steps at trim input) and the
 * IDHPnonlin.train() prologue (objects.py:1466-1488).  Weights: planes of double, W1a [40][stride_in]
 * ((4,10) row-major), W2a [10], W1c [40], W2c [30] ((10,3) row-major). */
int rl4_nl_init(int policy, const rl4_nl_params* p, const double* w1a, const double* w2a, const double* w1c,
                const double* w2c, int64_t stride_in, rl4_nl_state st, int64_t n_agents, void* stream);
/* IDHPnonlin.train() loop (objects.py:1491-1558) fused with Ce500NonLinear.step (envs/nonlinear/env.py:182-256),
 * the plant step, Critic_big / Actor_big (objects.py:283-437), RLS.update and _adapt_check (:1212-1290), for steps
 * [k0, k0+n_steps).  theta_ref: device table (phi / psi references are zero, idhp_nonlin.py:116-117);
 * noise: device float [n_steps][noise_stride], the N(0,1) draw of tf.random.normal (objects.py:1375) per agent and
 * step -- an explicit input like the weights (TensorFlow's stream is not reproducible);
 * log: buf NULL, or level 1 (RL4_NLL_COUNT fields) / 2 (RL4_NLF_COUNT) / 3 (RL4_NLM_COUNT) for the first n_agents_logged
 * agents, layout as rl4_sp_log; the row of the step whose state turned NaN and every later row are NaN (objects.py:1168-1175). */
int rl4_nl_run(int policy, const rl4_nl_params* p, const double* theta_ref, const float* noise, int64_t noise_stride,
               int32_t k0, int32_t n_steps, rl4_nl_state st, int64_t n_agents, rl4_sp_log log, void* stream);
/* RLS.update (objects.py:492-543) with the nonlinear task's dimensions (3 states + 1 action): theta [12][stride] ((4,3)
 * row-major), cov [16][stride] in/out; dx0 [3][stride], da0 [stride], dx1 [3][stride]; out eps [3][stride], eps_norm
 * [stride]; all double (both policies keep the RLS in float64).  Forgetting factor = p->hp[RL4_NHP_RLS_GAMMA] or its
 * per-agent override. */
int rl4_nl_rls_update(const rl4_nl_params* p, double* theta, double* cov, const double* dx0, const double* da0,
                      const double* dx1, double* out_eps, double* out_eps_norm, int64_t stride, int64_t n_agents, void* stream);
/* Critic_big.call (objects.py:294-339): s [4][stride], w1 [40][stride] ((4,10) row-major), w2 [30][stride] ((10,3)
 * row-major), all TN; out_lambda [3][stride]. */
int rl4_nl_critic_forward(int policy, const void* s, void* w1, void* w2, void* out_lambda, int64_t stride,
                          int64_t n_agents, void* stream);
/* Actor_big.call (objects.py:374-407) and tape.gradient(a, s) (objects.py:1323): w1 [40][stride], w2 [10][stride] (TN),
 * E [50][stride] (double, in/out when trace != 0), out_a [stride], out_dads [4][stride] (may be NULL). */
int rl4_nl_actor_forward(int policy, const void* s, void* w1, void* w2, double* E, void* out_a, void* out_dads,
                         double gamma_lambda, int32_t elig, int32_t trace, int64_t stride, int64_t n_agents, void* stream);
/* Ce500NonLinear.step alone (envs/nonlinear/env.py:182-256): action [3][stride] normalised commands (double),
 * x_full [12][stride] (the plant's carried state), x_act [3][stride] in/out; out_mdp [4][stride], out_reward
 * (longitudinal), out_e_theta [stride]; out_surf [3][stride] = info['action_commanded'] (surface positions after
 * saturation), out_eff [3][stride] = info['action_effective'] (model_input[:3]), out_x_obs [12][stride] =
 * info['x_full'] = what model.step returned -- the reference's plant is an output-then-update block, step() returns
 * the state BEFORE the step (oracle/pe_probe); the MDP state, errors and rewards are functions of it.  The three
 * optional outputs may be NULL. */
int rl4_nl_env_step(const rl4_nl_params* p, const double* theta_ref, int32_t stepp, double* x_full, double* x_act,
                    const double* action, double* out_mdp, double* out_reward, double* out_e_theta, double* out_surf,
                    double* out_eff, double* out_x_obs, int64_t stride, int64_t n_agents, void* stream);

/* ---- step-API arithmetic that the reference does in TensorFlow / numpy one-liners ---- */
/* Critic / Actor / Critic_big.soft_update (objects.py:207-215, 273-281, 353-361): target <- (1 - tau) target + tau source,
 * three separately rounded operations per weight in the network dtype TN; target, source: [n_rows][stride] planes. */
int rl4_soft_update(int policy, void* target, const void* source, double tau, int32_t n_rows, int64_t stride,
                    int64_t n_agents, void* stream);
/* Actor / Actor_big.get_weight_update (objects.py:261-271, 417-427): out[r] = loss * (TN) E[r]; loss [stride] (TN),
 * E [n_rows][stride] (TE: the trace dtype of the policy, double on the nonlinear path), out [n_rows][stride] (TN). */
int rl4_actor_weight_update(int policy, const void* loss, const void* E, void* out, int32_t n_rows, int64_t stride,
                            int64_t n_agents, void* stream);

/* ---- episode statistics ---- */
/* Per-agent numbers of MC_run_seed (functions.py:53,57; utils.py:350-369) from the state planes after n_steps steps,
 * out [RL4_SPS_COUNT][out_stride] doubles.  nMAE = mean|e| / (max ref - min ref) is an addition of this repo
 * (BASELINE.json asks for it; the reference has no such statistic): ref = hp[REF_AMP] * ref_base, so the caller passes
 * the extremes of ref_base over the episode. */
enum rl4_sp_stat_field {
    RL4_SPS_SUM_C = 0,      /* sum(c) / kappa                                  functions.py:53 */
    RL4_SPS_CONV_TIME,      /* dt * last step with |alpha error| > 0.5 deg     utils.py:350-369 */
    RL4_SPS_DIVERGED,       /* 1.0 if the run diverged (NaN state / reward)    functions.py:162; objects.py:991 */
    RL4_SPS_UNSTEADY,       /* 1.0 if CONV_TIME > 30 s                         functions.py:165 */
    RL4_SPS_MEAN_ABS_E,     /* sum|e| / n_steps */
    RL4_SPS_NMAE,           /* MEAN_ABS_E / |amp * ref_base_max - amp * ref_base_min| */
    RL4_SPS_COUNT
};
int rl4_sp_agent_stats(int policy, const rl4_sp_params* p, rl4_sp_state st, int64_t n_agents, int32_t n_steps,
                       double ref_base_min, double ref_base_max, double* out, int64_t out_stride, void* stream);
/* Per-agent numbers MC_test_hparam keeps (functions.py:916-917, 1036-1039, 1050-1055), out [RL4_NLS_COUNT][out_stride]. */
enum rl4_nl_stat_field { RL4_NLS_RSE_WARMUP = 0, RL4_NLS_RSE_FLIGHT, RL4_NLS_RSE_LAT, RL4_NLS_NZ_PEAK, RL4_NLS_DIVERGED,
                         RL4_NLS_COUNT };
int rl4_nl_agent_stats(rl4_nl_state st, int64_t n_agents, double* out, int64_t out_stride, void* stream);
/* MC_run's reduction over runs (functions.py:161-182): sums of n_fields (<= 8) per-agent planes [n_fields][stride] over
 * n_agents agents.  `exclude` (may be NULL): a plane; agents with exclude != 0 (diverged runs, functions.py:176-178)
 * are left out of the "kept" sums.  out [2 * n_fields + 2] (device) = for each field: sum over kept agents, sum over
 * all agents; then the number of kept and of excluded agents.  Deterministic (fixed summation order, no atomics).
 * `work`: device scratch of at least rl4_stats_reduce_work_doubles() doubles. */
int64_t rl4_stats_reduce_work_doubles(void);
int rl4_stats_reduce(const double* planes, int64_t stride, int32_t n_fields, const double* exclude, int64_t n_agents,
                     double* out, double* work, int64_t work_doubles, void* stream);

/* ---- host-buffer episodes (what a reference user calls: IDHPsp(...).train() / IDHPnonlin(...).train() for a batch) ----
 * Copy the inputs from host memory, run init + the fused loop for n_steps on the GPU, copy the selected results back.
 * Host buffers may be pageable or pinned (pinned: the agent chunks pipeline H2D / kernel / D2H over four streams). */
typedef struct rl4_ctx rl4_ctx;
/* `max_steps` bounds n_steps of the episodes run through this context; a context serves both paths (the nonlinear
 * buffers are allocated by the first rl4_nl_episode_host call). */
int rl4_ctx_create(int device, int policy, int64_t max_agents, int32_t max_steps, rl4_ctx** out);
int rl4_ctx_destroy(rl4_ctx* ctx);
/* which groups of final-state fields travel back (rows of the same full-plane host layout; unselected rows are not written) */
enum rl4_out_mask {
    RL4_OUT_STATS = 1,      /* SP: SUM_C, SUM_ABS_E + the int plane;  NL: RSE, NZ_PEAK, RSE_FLIGHT, ETA / LAMBDA + the int plane */
    RL4_OUT_WEIGHTS = 2,    /* actor and critic weights (IDHPsp.actor / critic.trainable_weights) */
    RL4_OUT_RLS = 4,        /* THETA, COV, EPS, EPS_NORM (IDHPsp.model) */
    RL4_OUT_STATE = 8,      /* plant state, actions, M_prev, previous reward gradient, learning rates: what a resume needs */
    RL4_OUT_TRACES = 16,    /* eligibility / Jacobian traces */
    RL4_OUT_TARGET = 32,    /* target-critic weights (IDHPsp.target_critic.trainable_weights) */
    RL4_OUT_ALL = 63
};
typedef struct rl4_sp_host_io {
    const double* x0;           /* [2][n] */
    const double* w1a;          /* [4][n] */
    const double* w2a;          /* [4][n] */
    const double* w1c;          /* [4][n] */
    const double* w2c;          /* [8][n] */
    const double* ref_base;     /* [n_steps] */
    void*    out_env;           /* [RL4_SPE_COUNT][n] of TE */
    void*    out_net;           /* [RL4_SPN_COUNT][n] of TN */
    int32_t* out_ints;          /* [RL4_SPI_COUNT][n] */
    int32_t  out_mask;          /* rl4_out_mask bits; 0 = RL4_OUT_ALL */
    int32_t  reserved;
} rl4_sp_host_io;
int rl4_sp_episode_host(rl4_ctx* ctx, const rl4_sp_params* p, const rl4_sp_host_io* io,
                        int64_t n_agents, int32_t n_steps, int32_t use_traces);

/* N(0,1) float32 draws for the nonlinear agent's policy-smoothing term (tf.random.normal, objects.py:1375): Philox4x32-10
 * keyed by `seed`, counter = (agent0 + i, step), Box-Muller; out[(k - k0) * stride + i] for k in [k0, k0 + n_steps),
 * i in [0, n_agents): a pure function of (seed, global agent index, step), so any split over chunks, launches or GPUs
 * draws the same numbers.  TensorFlow's stream is not reproducible, so this is a stream of this repo; tests read it back
 * and feed the oracle. */
int rl4_nl_noise_fill(uint64_t seed, int64_t agent0, int32_t k0, int32_t n_steps, int64_t n_agents, float* out, int64_t stride,
                      void* stream);
typedef struct rl4_nl_host_io {
    const double* w1a;          /* [40][n] (4,10) row-major */
    const double* w2a;          /* [10][n] */
    const double* w1c;          /* [40][n] */
    const double* w2c;          /* [30][n] (10,3) row-major */
    const double* theta_ref;    /* [n_steps] */
    const float*  noise;        /* [n_steps][n] host N(0,1) draws, or NULL: generated on the device by rl4_nl_noise_fill(noise_seed, ...) */
    uint64_t      noise_seed;
    int64_t       noise_agent0; /* global index of this batch's first agent in the device-drawn stream (multi-GPU shards) */
    double*  out_env;           /* [RL4_NLE_COUNT][n] */
    void*    out_net;           /* [RL4_NLN_COUNT][n] of TN */
    int32_t* out_ints;          /* [RL4_NLI_COUNT][n] */
    int32_t  out_mask;          /* rl4_out_mask bits; 0 = RL4_OUT_ALL */
    int32_t  reserved;
} rl4_nl_host_io;
/* IDHPnonlin(...).train() for a batch from host buffers (objects.py:1457-1564): rl4_nl_init + rl4_nl_run in chunks of
 * agents x steps.  The context's policy must be RL4_MIXED or RL4_FP64. */
int rl4_nl_episode_host(rl4_ctx* ctx, const rl4_nl_params* p, const rl4_nl_host_io* io, int64_t n_agents, int32_t n_steps);

/* ---- the reference's own nonlinear aircraft ("dasmat" plant) -----------------------------------------------------------
 * Replaces `_citation.initialize / step / terminate` (envs/nonlinear/citation.py:62-69; called at envs/nonlinear/env.py:210,
 * 288-291): the DASMAT Citation model the reference ships as a Windows binary, translated from that binary's machine code at
 * build time (rl4afcs_b200/tools/lift_plant.py -> rl4afcs_b200/csrc/_gen/, only where the reference tree exists) and compiled for
 * sm_100a, one aircraft per thread.  A library built without the binary exports the same symbols; they return -2 and
 * rl4_last_error() says so.  Buffers (all device memory, caller-owned):
 *   image        rl4_dasmat_image_bytes() bytes: the model's memory image after initialize(); shared, read-only afterwards
 *   plant_state  uint64 [rl4_dasmat_state_words()][stride]: the part of the model's memory that step() writes, per aircraft
 *                (continuous states at words rl4_dasmat_word_x() .. +12 and rl4_dasmat_word_engine() .. +4, as doubles)
 *   device_err   one int32, OR of error bits raised by any aircraft (1 untranslated path / bad indirect call, 2 access
 *                outside the model's memory, 4 store into the shared image); 0 after a valid run */
int     rl4_dasmat_available(void);
int64_t rl4_dasmat_image_bytes(void);
int32_t rl4_dasmat_state_words(void);
int32_t rl4_dasmat_word_x(void);
int32_t rl4_dasmat_word_engine(void);
/* citation.initialize(): fills `image` (synchronises the stream: the model's own initialisation code runs on the device) */
int rl4_dasmat_initialize(void* image, void* stream);
/* every aircraft <- the initialised model */
int rl4_dasmat_reset(const void* image, uint64_t* plant_state, int64_t stride, int64_t n_agents, void* stream);
/* n_steps calls of citation.step(u) per aircraft: u [11][u_stride] held over the launch; out [12][out_stride] = the last
 * call's return (may be NULL); out_all [n_steps][12][out_stride] = every call's return (may be NULL) */
int rl4_dasmat_step(void* image, uint64_t* plant_state, int64_t stride, int64_t n_agents, const double* u, int64_t u_stride,
                    int32_t n_steps, double* out, int64_t out_stride, double* out_all, int32_t* device_err, void* stream);
/* aircraft `src_index` of one state plane -> every aircraft of another (the trimmed state after Ce500NonLinear.reset) */
int rl4_dasmat_broadcast(const uint64_t* src, int64_t src_stride, int64_t src_index, uint64_t* dst, int64_t dst_stride, int64_t n_agents,
                         void* stream);
/* rl4_nl_env_step / rl4_nl_run with the dasmat plant in place of the surrogate (policy RL4_MIXED; integrator = the model's own) */
int rl4_nl_env_step_dasmat(const rl4_nl_params* p, const double* theta_ref, int32_t stepp, double* x_full, double* x_act,
                           const double* action, double* out_mdp, double* out_reward, double* out_e_theta, double* out_surf,
                           double* out_eff, double* out_x_obs, int64_t stride, int64_t n, void* image, uint64_t* plant_state,
                           int64_t plant_stride, int32_t* device_err, void* stream);
int rl4_nl_run_dasmat(int policy, const rl4_nl_params* p, const double* theta_ref, const float* noise, int64_t noise_stride,
                      int32_t k0, int32_t n_steps, rl4_nl_state st, int64_t n, rl4_sp_log lg, void* image, uint64_t* plant_state,
                      int64_t plant_stride, int32_t* device_err, void* stream);

/* ---- measurement helpers (bench.py roofline denominators) ---- */
/* Runs an FMA micro-kernel and returns achieved FLOP/s.  is_double: 0 = FFMA, 1 = DFMA with 16 independent chains per thread
 * whose multiplier and addend are shared (the pipe's peak, the roofline denominator); 2 / 3 = FFMA / DFMA whose three
 * operands are distinct registers (what the pipe sustains on the agent kernels' operand pattern; reported beside it). */
int rl4_peak_fma(int is_double, double* out_flops_per_s, void* stream);
/* Test hook: element-wise probe of the arithmetic primitives on device arrays.
 * op 0: tanh t13 (double)  1: tanh t13 (float)  2: shared-reciprocal division a/b (double; RLS gain / covariance)
 * 3: __ddiv_rn(a, b)  4 / 5: rl4_sincos sine / cosine  6 / 7: ISA density / thrust lapse at altitude a (b = device copy
 * of an rl4_cit_params).  Used by tests/test_gpu_math.py only. */
int rl4_test_math(int op, const void* a, const void* b, void* out, int64_t n, void* stream);
/* Test hook: exhaustive comparison of the float t13 quotient em/(em+2) with __fdiv_rn over the float
 * bit patterns [lo_bits, hi_bits); adds the number of mismatches to *device_mismatch_counter. */
int rl4_test_t13_div_f32(uint32_t lo_bits, uint32_t hi_bits, unsigned long long* device_mismatch_counter, void* stream);
/* Test hook: exhaustive comparison of the float32 reciprocal that seeds the double t13 quotient (MUFU.RCP + one Newton
 * step) with __frcp_rn (IEEE 1/d) over the float bit patterns [lo_bits, hi_bits); adds the mismatches to the counter. */
int rl4_test_rcp_f32(uint32_t lo_bits, uint32_t hi_bits, unsigned long long* device_mismatch_counter, void* stream);
/* number of kernel launches issued by this library since load (bench.py "gpu_launches") */
int64_t rl4_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
