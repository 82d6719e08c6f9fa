/* TEST INFRASTRUCTURE ONLY -- see sp_oracle.h.  Build: make -C oracle  (gcc, -ffp-contract=off). */
#include "sp_oracle.h"
#include "sp_oracle_tanh.h"
#include <math.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* numpy: npy_deg2radf(x) = x * (NPY_PIf / 180.0f) (npy_math_internal.h.src); checked in tests */
#define ORC_DEG2RAD_F (3.141592653589793238462643383279502884F / 180.0F)
#define ORC_DEG2RAD_D (M_PI / 180.0)
#define ORC_RAD2DEG_D (180.0 / M_PI)
#define ORC_RAD2DEG_F (180.0F / 3.141592653589793238462643383279502884F)

/* ---- mixed: the reference's own dtype mix (TF float32 nets, numpy float64 env/RLS) ---- */
#define TN float
#define TE double
#define SFX _mixed
#define N_FMA fmaf
#define N_SQRT sqrtf
#define N_T13 orc_t13_f32
#define N_LIBM_TANH tanhf
#define N_DEG2RAD ORC_DEG2RAD_F
#define E_FMA fma
#define E_SQRT sqrt
#define E_ABS fabs
#define E_DEG2RAD ORC_DEG2RAD_D
#define E_RAD2DEG ORC_RAD2DEG_D
#include "sp_oracle_body.inc"
#undef TN
#undef TE
#undef SFX
#undef N_FMA
#undef N_SQRT
#undef N_T13
#undef N_LIBM_TANH
#undef N_DEG2RAD
#undef E_FMA
#undef E_SQRT
#undef E_ABS
#undef E_DEG2RAD
#undef E_RAD2DEG

/* ---- fp64: everything double ---- */
#define TN double
#define TE double
#define SFX _fp64
#define N_FMA fma
#define N_SQRT sqrt
#define N_T13 orc_t13_f64
#define N_LIBM_TANH tanh
#define N_DEG2RAD ORC_DEG2RAD_D
#define E_FMA fma
#define E_SQRT sqrt
#define E_ABS fabs
#define E_DEG2RAD ORC_DEG2RAD_D
#define E_RAD2DEG ORC_RAD2DEG_D
#include "sp_oracle_body.inc"
#undef TN
#undef TE
#undef SFX
#undef N_FMA
#undef N_SQRT
#undef N_T13
#undef N_LIBM_TANH
#undef N_DEG2RAD
#undef E_FMA
#undef E_SQRT
#undef E_ABS
#undef E_DEG2RAD
#undef E_RAD2DEG

/* ---- fp32: everything float ---- */
#define TN float
#define TE float
#define SFX _fp32
#define N_FMA fmaf
#define N_SQRT sqrtf
#define N_T13 orc_t13_f32
#define N_LIBM_TANH tanhf
#define N_DEG2RAD ORC_DEG2RAD_F
#define E_FMA fmaf
#define E_SQRT sqrtf
#define E_ABS fabsf
#define E_DEG2RAD ORC_DEG2RAD_F
#define E_RAD2DEG ORC_RAD2DEG_F
#include "sp_oracle_body.inc"

int orc_sp_run(int policy, int tanh_mode, const orc_sp_cfg* cfgs, int cfg_stride,
               const double* ref_base, int k0, int n_steps,
               orc_sp_state* states, int64_t n_agents,
               orc_sp_logrow* log, int64_t n_log_agents)
{
    if (!cfgs || !ref_base || !states || n_agents < 0 || n_steps < 0 || k0 < 0) return -1;
    if (tanh_mode != ORC_TANH_LIBM && tanh_mode != ORC_TANH_T13) return -1;
    switch (policy) {
    case ORC_POLICY_FP64:  run_fp64(tanh_mode, cfgs, cfg_stride, ref_base, k0, n_steps, states, n_agents, log, n_log_agents); return 0;
    case ORC_POLICY_FP32:  run_fp32(tanh_mode, cfgs, cfg_stride, ref_base, k0, n_steps, states, n_agents, log, n_log_agents); return 0;
    case ORC_POLICY_MIXED: run_mixed(tanh_mode, cfgs, cfg_stride, ref_base, k0, n_steps, states, n_agents, log, n_log_agents); return 0;
    default: return -1;
    }
}

int orc_sp_init(int policy, const orc_sp_cfg* cfgs, int cfg_stride,
                const double* x0, const double* W1a, const double* W2a,
                const double* W1c, const double* W2c,
                orc_sp_state* states, int64_t n_agents)
{
    if (!cfgs || !x0 || !W1a || !W2a || !W1c || !W2c || !states || n_agents < 0) return -1;
    for (int64_t i = 0; i < n_agents; ++i) {
        const orc_sp_cfg* c = cfgs + (cfg_stride ? i : 0);
        switch (policy) {
        case ORC_POLICY_FP64:  init_one_fp64(c, x0 + 2 * i, W1a + 4 * i, W2a + 4 * i, W1c + 4 * i, W2c + 8 * i, states + i); break;
        case ORC_POLICY_FP32:  init_one_fp32(c, x0 + 2 * i, W1a + 4 * i, W2a + 4 * i, W1c + 4 * i, W2c + 8 * i, states + i); break;
        case ORC_POLICY_MIXED: init_one_mixed(c, x0 + 2 * i, W1a + 4 * i, W2a + 4 * i, W1c + 4 * i, W2c + 8 * i, states + i); break;
        default: return -1;
        }
    }
    return 0;
}

int orc_sp_env_step(int policy, const orc_sp_cfg* cfgs, int cfg_stride, const double* ref_base, int stepp,
                    double* x, const double* action_deg, double* reward, double* e, double* rg0, int64_t n)
{
    if (!cfgs || !ref_base || !x || !action_deg || !reward || !e || !rg0 || n < 0 || stepp < 0) return -1;
    switch (policy) {
    case ORC_POLICY_FP64:  env_step_batch_fp64(cfgs, cfg_stride, ref_base, stepp, x, action_deg, reward, e, rg0, n); return 0;
    case ORC_POLICY_FP32:  env_step_batch_fp32(cfgs, cfg_stride, ref_base, stepp, x, action_deg, reward, e, rg0, n); return 0;
    case ORC_POLICY_MIXED: env_step_batch_mixed(cfgs, cfg_stride, ref_base, stepp, x, action_deg, reward, e, rg0, n); return 0;
    default: return -1;
    }
}

int orc_sp_rls_update(int policy, const orc_sp_cfg* cfgs, int cfg_stride, double* theta, double* cov, const double* dx0,
                      const double* da0, const double* dx1, double* eps, double* eps_norm, int64_t n)
{
    if (!cfgs || !theta || !cov || !dx0 || !da0 || !dx1 || !eps || !eps_norm || n < 0) return -1;
    switch (policy) {
    case ORC_POLICY_FP64:  rls_update_batch_fp64(cfgs, cfg_stride, theta, cov, dx0, da0, dx1, eps, eps_norm, n); return 0;
    case ORC_POLICY_FP32:  rls_update_batch_fp32(cfgs, cfg_stride, theta, cov, dx0, da0, dx1, eps, eps_norm, n); return 0;
    case ORC_POLICY_MIXED: rls_update_batch_mixed(cfgs, cfg_stride, theta, cov, dx0, da0, dx1, eps, eps_norm, n); return 0;
    default: return -1;
    }
}

void orc_tanh_t13_f64_array(const double* x, double* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = orc_t13_f64(x[i]); }
void orc_tanh_t13_f32_array(const float* x, float* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = orc_t13_f32(x[i]); }
double orc_tanh_t13_f64(double x) { return orc_t13_f64(x); }
float  orc_tanh_t13_f32(float x)  { return orc_t13_f32(x); }
int orc_sizeof_cfg(void)    { return (int)sizeof(orc_sp_cfg); }
int orc_sizeof_state(void)  { return (int)sizeof(orc_sp_state); }
int orc_sizeof_logrow(void) { return (int)sizeof(orc_sp_logrow); }
