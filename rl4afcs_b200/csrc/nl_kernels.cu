// Nonlinear IDHP path (sm_100a): Ce500NonLinear wrapper + surrogate 6-DOF plant + IDHPnonlin agent,
// fused into one persistent kernel, one aircraft+agent per thread.
//
// The per-agent state (~1.7 KB mixed / 2.6 KB fp64: 4-10-{1,3} nets, traces, 4x4 RLS covariance, 12 plant states)
// exceeds the 255-register budget: the actor trace and the target critic live in shared memory, ptxas keeps the rest
// in registers + ~1.3 KB of local memory.  256 agents per SM fill the register file and 174 KB of shared memory;
// DESIGN.md section 9 has the capacity analysis and the measured alternatives (RL4_NL_SMEM_* / RL4_NL_BLOCK_* / RL4_NL_STEP_BARRIER below).
//
// Arithmetic: built with -fmad=false, FMAs explicit; numpy-side `@` orders as measured for these
// shapes, TensorFlow-side `@` in-order chains, tanh = t13.
#include "nl_core.cuh"
#include "nl_pipeline.cuh"
namespace rl4 {

// Ce500NonLinear.reset (envs/nonlinear/env.py:278-291): the model's built-in initial state, then 1000 + 1 steps at trim
// input.  Every agent shares the plant and the trim input, and the plant is IEEE-basic-operations only, so the 1001 steps
// are integrated ONCE on the host (same header, same bits as on the device) and broadcast by the init kernel.
// The plant is an output-then-update block (model.step returns the state before the step): the 1001 calls of reset leave
// the carried state at 1001 integrations (x) and return the state after 1000 of them (x_obs = env.state after reset).
struct NlTrim { double x[12]; double x_obs[12]; };
static NlTrim nl_trim_on_host(const rl4_nl_params* p)
{
    NlTrim t = {{0, 0, 0, 90, 0.0576, 0, 0, 0.0576, 0, 2000, 0, 0}, {0}};
    const int n_trim = (int)(10.0 / p->dt) + 1;
    for (int k = 0; k < n_trim; ++k) {
        memcpy(t.x_obs, t.x, sizeof t.x_obs);
        if (p->integrator == RL4_CIT_INTEGRATOR_RK4) rl4_cit_step_rk4(&p->plant, t.x, p->trim_input, p->dt);
        else rl4_cit_step_ode5(&p->plant, t.x, p->trim_input, p->dt);
    }
    return t;
}

// Ce500NonLinear.reset + IDHPnonlin prologue
template <typename TN>
__global__ void __launch_bounds__(128)
nl_init_kernel(const __grid_constant__ rl4_nl_params p, const NlTrim trim, const double* __restrict__ w1a, const double* __restrict__ w2a,
               const double* __restrict__ w1c, const double* __restrict__ w2c, int64_t stride_in, const rl4_nl_state st, int64_t n_agents)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_agents) return;
    const NlHp<true> hv{p, i};
    const int64_t S = st.stride;
    double* E = st.env + i;
    TN* Nn = (TN*)st.net + i;
    for (int f = 0; f < RL4_NLE_COUNT; ++f) E[(int64_t)f * S] = 0.0;
    for (int f = 0; f < RL4_NLN_COUNT; ++f) Nn[(int64_t)f * S] = TN(0);
    // reset (env.py:278-291): the trimmed state is the same for every agent and was integrated once on the host
    for (int j = 0; j < 12; ++j) E[(int64_t)(RL4_NLE_XFULL + j) * S] = trim.x[j];
    const double c0 = hv.hp(RL4_NHP_RLS_COV0);
    for (int d = 0; d < 4; ++d) E[(int64_t)(RL4_NLE_COV + d * 5) * S] = c0;
    E[(int64_t)RL4_NLE_ETA_A * S] = hv.hp(RL4_NHP_ETA_A_H);
    E[(int64_t)RL4_NLE_ETA_C * S] = hv.hp(RL4_NHP_ETA_C_H);
    E[(int64_t)RL4_NLE_LAMBDAA * S] = hv.hp(RL4_NHP_LAMBDA_H);
    E[(int64_t)RL4_NLE_GL * S] = hv.hp(RL4_NHP_GAMMA) * hv.hp(RL4_NHP_LAMBDA_H);
    for (int j = 0; j < 40; ++j) {
        const TN wa = (TN)w1a[j * stride_in + i], wc = (TN)w1c[j * stride_in + i];
        Nn[(int64_t)(RL4_NLN_W1A + j) * S] = wa; Nn[(int64_t)(RL4_NLN_W1C + j) * S] = wc; Nn[(int64_t)(RL4_NLN_W1T + j) * S] = wc;
    }
    for (int j = 0; j < 10; ++j) Nn[(int64_t)(RL4_NLN_W2A + j) * S] = (TN)w2a[j * stride_in + i];
    for (int j = 0; j < 30; ++j) { const TN wc = (TN)w2c[j * stride_in + i]; Nn[(int64_t)(RL4_NLN_W2C + j) * S] = wc; Nn[(int64_t)(RL4_NLN_W2T + j) * S] = wc; }
    Nn[(int64_t)RL4_NLN_LR_A * S] = (TN)hv.hp(RL4_NHP_ETA_A_H);
    Nn[(int64_t)RL4_NLN_LR_C * S] = (TN)hv.hp(RL4_NHP_ETA_C_H);
    st.ints[(int64_t)RL4_NLI_COOLDOWN * S + i] = 0;
    st.ints[(int64_t)RL4_NLI_DIVERGED_STEP * S + i] = -1;
    st.ints[(int64_t)RL4_NLI_STEPP * S + i] = 0;
    st.ints[(int64_t)RL4_NLI_PYFLOAT_MASK * S + i] = 7;
}

__global__ void __launch_bounds__(128)
nl_env_step_kernel(const __grid_constant__ rl4_nl_params p, const double* __restrict__ theta_ref, int stepp, double* __restrict__ x_full,
                   double* __restrict__ x_act_p, const double* __restrict__ action, double* __restrict__ out_mdp,
                   double* __restrict__ out_reward, double* __restrict__ out_e, double* __restrict__ out_surf,
                   double* __restrict__ out_eff, double* __restrict__ out_x_obs, int64_t S, int64_t n_agents)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_agents) return;
    const NlHp<true> hv{p, i};
    double x[12], xo[12], xa[3], act[3], surf[3], ueff[3], e_phi, e_th, e_psi, reward, rg2;
    for (int j = 0; j < 12; ++j) x[j] = x_full[j * S + i];
    for (int j = 0; j < 3; ++j) { xa[j] = x_act_p[j * S + i]; act[j] = action[j * S + i]; }
    if (p.integrator == RL4_CIT_INTEGRATOR_RK4)
        nl_env_step<true, RL4_CIT_INTEGRATOR_RK4>(p, hv, stepp, __ldg(theta_ref + stepp), act, x, xa, surf, e_phi, e_th, e_psi, reward, rg2, ueff, xo);
    else
        nl_env_step<true, RL4_CIT_INTEGRATOR_ODE5>(p, hv, stepp, __ldg(theta_ref + stepp), act, x, xa, surf, e_phi, e_th, e_psi, reward, rg2, ueff, xo);
    for (int j = 0; j < 12; ++j) x_full[j * S + i] = x[j];
    for (int j = 0; j < 3; ++j) x_act_p[j * S + i] = xa[j];
    if (out_x_obs) for (int j = 0; j < 12; ++j) out_x_obs[j * S + i] = xo[j];     // info['x_full'] = what model.step returned
    out_mdp[i] = xo[4]; out_mdp[S + i] = xo[7]; out_mdp[2 * S + i] = xo[1]; out_mdp[3 * S + i] = e_th;
    out_reward[i] = reward; out_e[i] = e_th;
    if (out_surf) for (int j = 0; j < 3; ++j) out_surf[j * S + i] = surf[j];          // action_commanded (env.py:244)
    if (out_eff) for (int j = 0; j < 3; ++j) out_eff[j * S + i] = ueff[j];            // action_effective = model_input[:3] (:245)
}

// RLS.update (objects.py:492-543) for n = 3, m = 1: step-API form, planes of double
__global__ void __launch_bounds__(128)
nl_rls_update_kernel(double rgam, const double* __restrict__ rgam_agent, double* theta, double* cov, const double* __restrict__ dx0,
                     const double* __restrict__ da0, const double* __restrict__ dx1, double* __restrict__ out_eps,
                     double* __restrict__ out_eps_norm, int64_t S, int64_t n_agents)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_agents) return;
    double th[12], cv[16], Xr[4], Y[3], eps[3], eps_norm;
    for (int j = 0; j < 12; ++j) th[j] = theta[j * S + i];
    for (int j = 0; j < 16; ++j) cv[j] = cov[j * S + i];
    for (int j = 0; j < 3; ++j) { Xr[j] = dx0[j * S + i]; Y[j] = dx1[j * S + i]; }
    Xr[3] = da0[i];
    nl_rls_update(th, cv, Xr, Y, rgam_agent ? __ldg(rgam_agent + i) : rgam, eps, eps_norm);
    for (int j = 0; j < 12; ++j) theta[j * S + i] = th[j];
    for (int j = 0; j < 16; ++j) cov[j * S + i] = cv[j];
    for (int j = 0; j < 3; ++j) out_eps[j * S + i] = eps[j];
    out_eps_norm[i] = eps_norm;
}

template <typename TN, int INTEG, bool LOG, bool PER_AGENT>
static int nl_launch_one(const rl4_nl_params* p, const double* theta_ref, const float* noise, int64_t noise_stride, int k0,
                         int n_steps, rl4_nl_state st, int64_t n, rl4_sp_log lg, unsigned grid, cudaStream_t s)
{
    constexpr int BLK = NlBlock<TN>::v;
    const size_t smem = (sizeof(double) * kNlSmemDoubles + sizeof(TN) * kNlSmemNet) * BLK;
    grid = (unsigned)((n + BLK - 1) / BLK);
    // per launch: the attribute is per device and per function, and setting it costs microseconds
    RL4_CUDA(cudaFuncSetAttribute(nl_run_kernel<TN, INTEG, LOG, PER_AGENT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nl_run_kernel<TN, INTEG, LOG, PER_AGENT><<<grid, BLK, smem, s>>>(*p, theta_ref, noise, noise_stride, k0, n_steps, st, n, lg);
    return 0;
}

#ifndef RL4_NL_PIPELINE
#define RL4_NL_PIPELINE 1       // log-free launches take the two-role pipeline kernel (nl_pipeline.cuh); 0: always the one-thread kernel
#endif

template <typename TN, int INTEG, bool PER_AGENT>
static int nl_launch_pipe(const rl4_nl_params* p, const double* theta_ref, const float* noise, int64_t noise_stride, int k0,
                          int n_steps, rl4_nl_state st, int64_t n, cudaStream_t s)
{
    const size_t smem = PipeSmem<TN>::bytes;
    RL4_CUDA(cudaFuncSetAttribute(nl_pipe_kernel<TN, INTEG, PER_AGENT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((n + kPipeAgents - 1) / kPipeAgents);
    nl_pipe_kernel<TN, INTEG, PER_AGENT><<<grid, kPipeThreads, smem, s>>>(*p, theta_ref, noise, noise_stride, k0, n_steps, st, n);
    return 0;
}

template <typename TN>
static int nl_launch(const rl4_nl_params* p, const double* theta_ref, const float* noise, int64_t noise_stride, int k0, int n_steps,
                     rl4_nl_state st, int64_t n, rl4_sp_log lg, bool log, bool per_agent, unsigned grid, cudaStream_t s)
{
    const bool rk4 = (p->integrator == RL4_CIT_INTEGRATOR_RK4);
#if RL4_NL_PIPELINE
    // the pipeline kernel covers log-free launches in the 'none' / 'accumulating' trace modes (the 'replacing' mode needs a
    // norm over all 50 trace elements in a fixed order; a per-agent mode array may contain it)
    if (!log && p->hpi[RL4_NHPI_ELIG_A] != RL4_ELIG_REPLACING && !p->hpi_agent[RL4_NHPI_ELIG_A]) {
        if (per_agent) return rk4 ? nl_launch_pipe<TN, RL4_CIT_INTEGRATOR_RK4, true>(p, theta_ref, noise, noise_stride, k0, n_steps, st, n, s)
                                  : nl_launch_pipe<TN, RL4_CIT_INTEGRATOR_ODE5, true>(p, theta_ref, noise, noise_stride, k0, n_steps, st, n, s);
        return rk4 ? nl_launch_pipe<TN, RL4_CIT_INTEGRATOR_RK4, false>(p, theta_ref, noise, noise_stride, k0, n_steps, st, n, s)
                   : nl_launch_pipe<TN, RL4_CIT_INTEGRATOR_ODE5, false>(p, theta_ref, noise, noise_stride, k0, n_steps, st, n, s);
    }
#endif
    // logging launches always take the general (per-agent capable) instantiation; the log-free hot path is specialised
    if (log) return rk4 ? nl_launch_one<TN, RL4_CIT_INTEGRATOR_RK4, true, true>(p, theta_ref, noise, noise_stride, k0, n_steps, st, n, lg, grid, s)
                        : nl_launch_one<TN, RL4_CIT_INTEGRATOR_ODE5, true, true>(p, theta_ref, noise, noise_stride, k0, n_steps, st, n, lg, grid, s);
    if (per_agent) return rk4 ? nl_launch_one<TN, RL4_CIT_INTEGRATOR_RK4, false, true>(p, theta_ref, noise, noise_stride, k0, n_steps, st, n, lg, grid, s)
                              : nl_launch_one<TN, RL4_CIT_INTEGRATOR_ODE5, false, true>(p, theta_ref, noise, noise_stride, k0, n_steps, st, n, lg, grid, s);
    return rk4 ? nl_launch_one<TN, RL4_CIT_INTEGRATOR_RK4, false, false>(p, theta_ref, noise, noise_stride, k0, n_steps, st, n, lg, grid, s)
               : nl_launch_one<TN, RL4_CIT_INTEGRATOR_ODE5, false, false>(p, theta_ref, noise, noise_stride, k0, n_steps, st, n, lg, grid, s);
}

// Critic_big.call (objects.py:294-339): forward of the 4-10-3 critic (its trace is never formed, :304-305)
template <typename TN>
__global__ void __launch_bounds__(128)
nl_critic_forward_kernel(const TN* __restrict__ s_in, TN* w1, TN* w2, TN* __restrict__ out_lambda, int64_t S, int64_t n_agents)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_agents) return;
    TN s[4], h[10];
    for (int j = 0; j < 4; ++j) s[j] = s_in[j * S + i];
    nl_hidden<TN>(s, PlaneCol<TN>{w1 + i, S}, h);
    const PlaneCol<TN> W2{w2 + i, S};
    for (int q = 0; q < 3; ++q) {
        TN acc = h[0] * W2[q];
        for (int j = 1; j < 10; ++j) acc = nfma<TN>(h[j], W2[j * 3 + q], acc);
        out_lambda[q * S + i] = acc;
    }
}

// Actor_big.call (objects.py:374-407) + tape.gradient(a, s) (objects.py:1323)
template <typename TN>
__global__ void __launch_bounds__(128)
nl_actor_forward_kernel(const TN* __restrict__ s_in, TN* w1, TN* w2, double* Eplane, TN* __restrict__ out_a, TN* __restrict__ out_dads,
                        double gl, int elig, int trace, int64_t S, int64_t n_agents)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_agents) return;
    TN s[4], h[10], ai1;
    for (int j = 0; j < 4; ++j) s[j] = s_in[j * S + i];
    const PlaneCol<TN> W1{w1 + i, S}, W2{w2 + i, S};
    double Etmp[50];
    TN a;
    if (trace) {
        a = nl_actor<TN>(s, W1, W2, PlaneCol<double>{Eplane + i, S}, elig, gl, h, ai1);
    } else {
        for (int j = 0; j < 50; ++j) Etmp[j] = 0.0;
        a = nl_actor<TN>(s, W1, W2, (double*)Etmp, RL4_ELIG_NONE, gl, h, ai1);
    }
    out_a[i] = a;
    if (out_dads) {
        const TN g_o = TN(1) * ai1;
        for (int ii = 0; ii < 4; ++ii) {
            TN acc = ((g_o * W2[0]) * (TN(1) - h[0] * h[0])) * W1[ii * 10];
            for (int j = 1; j < 10; ++j) acc = nfma<TN>((g_o * W2[j]) * (TN(1) - h[j] * h[j]), W1[ii * 10 + j], acc);
            out_dads[ii * S + i] = acc;
        }
    }
}

}  // namespace rl4

using namespace rl4;

extern "C" {

int rl4_nl_default_params(rl4_nl_params* p)
{
    RL4_REQUIRE(p != nullptr, "p is NULL");
    memset(p, 0, sizeof(*p));
    rl4_cit_default_params(&p->plant);
    const double trim[11] = {-0.02855, 0, 0, 0, 0, 0, 0, 0, 0.55, 0.55, 0};       // idhp_nonlin.py:53
    for (int i = 0; i < 11; ++i) p->trim_input[i] = trim[i];
    p->dt = 0.01;                                                                   // idhp_nonlin.py:36
    p->hp[RL4_NHP_ETA_A_H] = 35.0; p->hp[RL4_NHP_ETA_A_L] = 5.0; p->hp[RL4_NHP_ETA_C_H] = 1.4; p->hp[RL4_NHP_ETA_C_L] = 0.7;
    p->hp[RL4_NHP_LAMBDA_H] = 0.95; p->hp[RL4_NHP_LAMBDA_L] = 0.95; p->hp[RL4_NHP_GAMMA] = 0.6; p->hp[RL4_NHP_GAMMA_SQ] = 0.36;
    p->hp[RL4_NHP_TAU] = 0.02; p->hp[RL4_NHP_LR_DECAY] = 0.998; p->hp[RL4_NHP_RLS_GAMMA] = 1.0; p->hp[RL4_NHP_RLS_COV0] = 1e6;
    p->hp[RL4_NHP_Q_SYM] = 2.0; p->hp[RL4_NHP_LAMBDA_T] = 0.012; p->hp[RL4_NHP_LAMBDA_S] = 0.001;   // idhp_nonlin.py:123-146; objects.py:1383
    p->hp[RL4_NHP_DAMP_FACTOR] = 0.3; p->hp[RL4_NHP_CG_SHIFT] = -0.5;              // envs/nonlinear/env.py:135-143
    p->noise_std[0] = 0.010; p->noise_std[1] = 0.010; p->noise_std[2] = 0.008; p->noise_std[3] = 0.003;   // objects.py:1377
    const double d2r = 3.14159265358979323846 / 180.0;
    p->omega0 = 13.0; p->omega_slow = 6.0; p->rate_limit = 19.7 * d2r;             // envs/nonlinear/env.py:74,145,177
    p->limit_deg[0] = 15.0; p->limit_deg[1] = 37.0; p->limit_deg[2] = 22.0;        // envs/nonlinear/env.py:107-109
    p->sat_limit[0] = 5.0 * d2r; p->sat_limit[1] = 18.0 * d2r; p->sat_limit[2] = 10.0 * d2r;
    p->hpi[RL4_NHPI_MULTISTEP] = 0; p->hpi[RL4_NHPI_WARMUP_STEPS] = 400; p->hpi[RL4_NHPI_COOLDOWN_STEPS] = 200;
    p->hpi[RL4_NHPI_FAULT_STEP] = -1; p->hpi[RL4_NHPI_FAULT_DAMP] = 0; p->hpi[RL4_NHPI_FAULT_SAT] = 0;
    p->hpi[RL4_NHPI_ELIG_A] = RL4_ELIG_ACCUMULATING;
    p->hpi[RL4_NHPI_FLIGHT_STEP] = 5500;
    p->hpi[RL4_NHPI_NUMPY2] = 1;          // NEP 50: the mode observed against the verbatim agent (numpy 2.3); 0 = numpy 1.x, unverified
    p->integrator = RL4_CIT_INTEGRATOR_ODE5;
    return 0;
}

int rl4_nl_trim_state(const rl4_nl_params* p, double* out_state, double* out_observed)
{
    RL4_REQUIRE(p != nullptr, "p is NULL");
    const NlTrim t = nl_trim_on_host(p);
    if (out_state) memcpy(out_state, t.x, sizeof t.x);
    if (out_observed) memcpy(out_observed, t.x_obs, sizeof t.x_obs);
    return 0;
}

int rl4_nl_init(int policy, const rl4_nl_params* p, const double* w1a, const double* w2a, const double* w1c, const double* w2c,
                int64_t stride_in, rl4_nl_state st, int64_t n, void* stream)
{
    RL4_REQUIRE(p && w1a && w2a && w1c && w2c && st.env && st.net && st.ints, "NULL argument");
    RL4_REQUIRE(n >= 0 && stride_in >= n && st.stride >= n, "bad size");
    if (n == 0) return 0;
    const unsigned grid = (unsigned)((n + 127) / 128);
    cudaStream_t s = (cudaStream_t)stream;
    const NlTrim trim = nl_trim_on_host(p);
    if (policy == RL4_MIXED) nl_init_kernel<float><<<grid, 128, 0, s>>>(*p, trim, w1a, w2a, w1c, w2c, stride_in, st, n);
    else if (policy == RL4_FP64) nl_init_kernel<double><<<grid, 128, 0, s>>>(*p, trim, w1a, w2a, w1c, w2c, stride_in, st, n);
    else { set_error("rl4_nl_init: policy %d not supported on the nonlinear path (mixed or fp64)", policy); return -1; }
    return check_launch("nl_init_kernel");
}

int rl4_nl_run(int policy, const rl4_nl_params* p, const double* theta_ref, const float* noise, int64_t noise_stride,
               int32_t k0, int32_t n_steps, rl4_nl_state st, int64_t n, rl4_sp_log lg, void* stream)
{
    RL4_REQUIRE(p && theta_ref && noise && st.env && st.net && st.ints, "NULL argument");
    RL4_REQUIRE(n >= 0 && st.stride >= n && noise_stride >= n && k0 >= 0 && n_steps >= 0, "bad size");
    if (lg.level != RL4_LOG_NONE) RL4_REQUIRE(lg.level >= 1 && lg.level <= 3 && lg.buf && lg.every >= 1 && lg.n_agents_logged >= 0 && lg.n_agents_logged <= n, "bad log descriptor");
    if (n == 0 || n_steps == 0) return 0;
    bool per_agent = false;
    for (int j = 0; j < RL4_NHP_COUNT; ++j) per_agent |= (p->hp_agent[j] != nullptr);
    for (int j = 0; j < RL4_NHPI_COUNT; ++j) per_agent |= (p->hpi_agent[j] != nullptr);
    const unsigned grid = 0;                                  // set per network dtype in nl_launch_one
    cudaStream_t s = (cudaStream_t)stream;
    const bool log = lg.level != RL4_LOG_NONE;
    int rc = 0;
    if (policy == RL4_MIXED) rc = nl_launch<float>(p, theta_ref, noise, noise_stride, k0, n_steps, st, n, lg, log, per_agent, grid, s);
    else if (policy == RL4_FP64) rc = nl_launch<double>(p, theta_ref, noise, noise_stride, k0, n_steps, st, n, lg, log, per_agent, grid, s);
    else { set_error("rl4_nl_run: policy %d not supported on the nonlinear path (mixed or fp64)", policy); return -1; }
    if (rc) return rc;
    return check_launch("nl_run_kernel");
}

int rl4_nl_rls_update(const rl4_nl_params* p, double* theta, double* cov, const double* dx0, const double* da0, const double* dx1,
                      double* out_eps, double* out_eps_norm, int64_t stride, int64_t n, void* stream)
{
    RL4_REQUIRE(p && theta && cov && dx0 && da0 && dx1 && out_eps && out_eps_norm, "NULL argument");
    RL4_REQUIRE(n >= 0 && stride >= n, "bad size");
    if (n == 0) return 0;
    nl_rls_update_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        p->hp[RL4_NHP_RLS_GAMMA], p->hp_agent[RL4_NHP_RLS_GAMMA], theta, cov, dx0, da0, dx1, out_eps, out_eps_norm, stride, n);
    return check_launch("nl_rls_update_kernel");
}

int rl4_nl_critic_forward(int policy, const void* s, void* w1, void* w2, void* out_lambda, int64_t stride, int64_t n, void* stream)
{
    RL4_REQUIRE(s && w1 && w2 && out_lambda, "NULL argument");
    RL4_REQUIRE(n >= 0 && stride >= n, "bad size");
    if (n == 0) return 0;
    const unsigned grid = (unsigned)((n + 127) / 128);
    cudaStream_t st = (cudaStream_t)stream;
    if (policy == RL4_MIXED) nl_critic_forward_kernel<float><<<grid, 128, 0, st>>>((const float*)s, (float*)w1, (float*)w2, (float*)out_lambda, stride, n);
    else if (policy == RL4_FP64) nl_critic_forward_kernel<double><<<grid, 128, 0, st>>>((const double*)s, (double*)w1, (double*)w2, (double*)out_lambda, stride, n);
    else { set_error("rl4_nl_critic_forward: policy %d not supported on the nonlinear path", policy); return -1; }
    return check_launch("nl_critic_forward_kernel");
}

int rl4_nl_actor_forward(int policy, const void* s, void* w1, void* w2, double* E, void* out_a, void* out_dads, double gamma_lambda,
                         int32_t elig, int32_t trace, int64_t stride, int64_t n, void* stream)
{
    RL4_REQUIRE(s && w1 && w2 && out_a && (E || !trace), "NULL argument");
    RL4_REQUIRE(n >= 0 && stride >= n, "bad size");
    RL4_REQUIRE(elig >= RL4_ELIG_NONE && elig <= RL4_ELIG_REPLACING, "bad elig");
    if (n == 0) return 0;
    const unsigned grid = (unsigned)((n + 127) / 128);
    cudaStream_t st = (cudaStream_t)stream;
    if (policy == RL4_MIXED) nl_actor_forward_kernel<float><<<grid, 128, 0, st>>>((const float*)s, (float*)w1, (float*)w2, E, (float*)out_a, (float*)out_dads, gamma_lambda, elig, trace, stride, n);
    else if (policy == RL4_FP64) nl_actor_forward_kernel<double><<<grid, 128, 0, st>>>((const double*)s, (double*)w1, (double*)w2, E, (double*)out_a, (double*)out_dads, gamma_lambda, elig, trace, stride, n);
    else { set_error("rl4_nl_actor_forward: policy %d not supported on the nonlinear path", policy); return -1; }
    return check_launch("nl_actor_forward_kernel");
}

int rl4_nl_env_step(const rl4_nl_params* p, const double* theta_ref, int32_t stepp, double* x_full, double* x_act,
                    const double* action, double* out_mdp, double* out_reward, double* out_e_theta, double* out_surf,
                    double* out_eff, double* out_x_obs, int64_t stride, int64_t n, void* stream)
{
    RL4_REQUIRE(p && theta_ref && x_full && x_act && action && out_mdp && out_reward && out_e_theta, "NULL argument");
    RL4_REQUIRE(n >= 0 && stride >= n && stepp >= 0, "bad size");
    if (n == 0) return 0;
    nl_env_step_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*p, theta_ref, stepp, x_full, x_act, action,
                                                                                        out_mdp, out_reward, out_e_theta, out_surf, out_eff, out_x_obs, stride, n);
    return check_launch("nl_env_step_kernel");
}

}  // extern "C"
