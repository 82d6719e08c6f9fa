// Short-period IDHP kernels (sm_100a) and their C-ABI entry points.
//
//   sp_run_kernel      the fused persistent hot loop: env step + IDHP update + RLS + adapt +
//                      statistics for n_steps, one agent per thread, state in registers,
//                      loaded from / stored to the SoA planes once per launch.
//   sp_init_kernel     IDHPsp.__init__ / train() prologue.
//   step-API kernels   sp_env_step / sp_rls_update / sp_critic_forward / sp_actor_forward:
//                      the same device functions, one call each (HBM-bound).
#include "sp_core.cuh"
#include "rl4_runtime.h"

namespace rl4 {

constexpr int kBlock = RL4_SP_BLOCK;
// minimum resident CTAs per SM requested from ptxas, per dtype policy (tuned on B200, profiles/README.md)
#ifndef RL4_MINB_FP64
#define RL4_MINB_FP64 3
#endif
#ifndef RL4_MINB_MIXED
#define RL4_MINB_MIXED 3
#endif
#ifndef RL4_MINB_FP32
#define RL4_MINB_FP32 4
#endif
#ifndef RL4_SMEM_FP64
#define RL4_SMEM_FP64 1
#endif
#ifndef RL4_SMEM_MIXED
#define RL4_SMEM_MIXED 0
#endif
#ifndef RL4_SMEM_FP32
#define RL4_SMEM_FP32 0
#endif
template <typename TN, typename TE> struct UseSmem { static constexpr bool v = RL4_SMEM_FP64; };
template <> struct UseSmem<float, double> { static constexpr bool v = RL4_SMEM_MIXED; };
template <> struct UseSmem<float, float> { static constexpr bool v = RL4_SMEM_FP32; };
template <typename TN, typename TE> struct MinBlocks { static constexpr int v = RL4_MINB_FP64; };
template <> struct MinBlocks<float, double> { static constexpr int v = RL4_MINB_MIXED; };
template <> struct MinBlocks<float, float> { static constexpr int v = RL4_MINB_FP32; };

template <typename T> struct Plane {
    T* base;
    int64_t stride;
    __device__ __forceinline__ Rn<T> ld(int f, int64_t i) const { return Rn<T>(base[(int64_t)f * stride + i]); }
    __device__ __forceinline__ void st(int f, int64_t i, Rn<T> v) const { base[(int64_t)f * stride + i] = v.v; }
};

template <typename TN, typename TE, bool SM>
__device__ __forceinline__ void sp_load(SpAgent<TN, TE, SM>& s, const rl4_sp_state& st, int64_t i, bool traces)
{
    const Plane<TE> e{(TE*)st.env, st.stride};
    const Plane<TN> n{(TN*)st.net, st.stride};
#pragma unroll
    for (int j = 0; j < 2; ++j) { s.x[j] = e.ld(RL4_SPE_X + j, i); s.xp[j] = e.ld(RL4_SPE_XPREV + j, i); s.eps[j] = e.ld(RL4_SPE_EPS + j, i); }
#pragma unroll
    for (int j = 0; j < 6; ++j) s.th[j] = e.ld(RL4_SPE_THETA + j, i);
#pragma unroll
    for (int j = 0; j < 9; ++j) s.cv[j] = e.ld(RL4_SPE_COV + j, i);
    s.cgp = e.ld(RL4_SPE_CGRAD_PREV, i);
    s.epsn = e.ld(RL4_SPE_EPS_NORM, i);
    s.eps_fresh = false;
    s.sumc = e.ld(RL4_SPE_SUM_C, i);
    s.sumabse = e.ld(RL4_SPE_SUM_ABS_E, i);
    if (traces) {
#pragma unroll
        for (int j = 0; j < 8; ++j) s.Ea[j] = e.ld(RL4_SPE_EA + j, i);
#pragma unroll
        for (int j = 0; j < 4; ++j) { s.EcH[j] = e.ld(RL4_SPE_EC_H + j, i); s.EcR0[j] = e.ld(RL4_SPE_EC_W1R0 + j, i); s.EcR1[j] = e.ld(RL4_SPE_EC_W1R1 + j, i); }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) s.Ea[j] = Rn<TE>(TE(0));
#pragma unroll
        for (int j = 0; j < 4; ++j) { s.EcH[j] = Rn<TE>(TE(0)); s.EcR0[j] = Rn<TE>(TE(0)); s.EcR1[j] = Rn<TE>(TE(0)); }
    }
    s.a = n.ld(RL4_SPN_A, i); s.ap = n.ld(RL4_SPN_APREV, i);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        s.W1a[j] = n.ld(RL4_SPN_W1A + j, i); s.W2a[j] = n.ld(RL4_SPN_W2A + j, i);
        s.W1c[j] = n.ld(RL4_SPN_W1C + j, i); s.W1t[j] = n.ld(RL4_SPN_W1T + j, i);
        s.Mp[j] = n.ld(RL4_SPN_MPREV + j, i);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { s.W2c[j] = n.ld(RL4_SPN_W2C + j, i); s.W2t[j] = n.ld(RL4_SPN_W2T + j, i); }
    s.eta_a = n.ld(RL4_SPN_ETA_A, i); s.eta_c = n.ld(RL4_SPN_ETA_C, i);
    s.cooldown = st.ints[(int64_t)RL4_SPI_COOLDOWN * st.stride + i];
    s.flags = st.ints[(int64_t)RL4_SPI_FLAGS * st.stride + i];
    s.diverged_step = st.ints[(int64_t)RL4_SPI_DIVERGED_STEP * st.stride + i];
    s.conv_step = st.ints[(int64_t)RL4_SPI_CONV_STEP * st.stride + i];
}

// the Jacobian traces (actor E (1,8), critic E: 12 unique of 24)
template <typename TN, typename TE, bool SM>
__device__ __forceinline__ void sp_store_traces(const SpAgent<TN, TE, SM>& s, const rl4_sp_state& st, int64_t i)
{
    const Plane<TE> e{(TE*)st.env, st.stride};
#pragma unroll
    for (int j = 0; j < 8; ++j) e.st(RL4_SPE_EA + j, i, s.Ea[j]);
#pragma unroll
    for (int j = 0; j < 4; ++j) { e.st(RL4_SPE_EC_H + j, i, s.EcH[j]); e.st(RL4_SPE_EC_W1R0 + j, i, s.EcR0[j]); e.st(RL4_SPE_EC_W1R1 + j, i, s.EcR1[j]); }
}

template <typename TN, typename TE, bool SM, bool WITH_TRACES = true>
__device__ __forceinline__ void sp_store(const SpAgent<TN, TE, SM>& s, const rl4_sp_state& st, int64_t i)
{
    const Plane<TE> e{(TE*)st.env, st.stride};
    const Plane<TN> n{(TN*)st.net, st.stride};
#pragma unroll
    for (int j = 0; j < 2; ++j) { e.st(RL4_SPE_X + j, i, s.x[j]); e.st(RL4_SPE_XPREV + j, i, s.xp[j]); e.st(RL4_SPE_EPS + j, i, s.eps[j]); }
#pragma unroll
    for (int j = 0; j < 6; ++j) e.st(RL4_SPE_THETA + j, i, s.th[j]);
#pragma unroll
    for (int j = 0; j < 9; ++j) e.st(RL4_SPE_COV + j, i, s.cv[j]);
    e.st(RL4_SPE_CGRAD_PREV, i, s.cgp);
    e.st(RL4_SPE_EPS_NORM, i, s.epsn);
    e.st(RL4_SPE_SUM_C, i, s.sumc);
    e.st(RL4_SPE_SUM_ABS_E, i, s.sumabse);
    if (WITH_TRACES) sp_store_traces<TN, TE, SM>(s, st, i);
    n.st(RL4_SPN_A, i, s.a); n.st(RL4_SPN_APREV, i, s.ap);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        n.st(RL4_SPN_W1A + j, i, s.W1a[j]); n.st(RL4_SPN_W2A + j, i, s.W2a[j]);
        n.st(RL4_SPN_W1C + j, i, s.W1c[j]); n.st(RL4_SPN_W1T + j, i, s.W1t[j]);
        n.st(RL4_SPN_MPREV + j, i, s.Mp[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { n.st(RL4_SPN_W2C + j, i, s.W2c[j]); n.st(RL4_SPN_W2T + j, i, s.W2t[j]); }
    n.st(RL4_SPN_ETA_A, i, s.eta_a); n.st(RL4_SPN_ETA_C, i, s.eta_c);
    st.ints[(int64_t)RL4_SPI_COOLDOWN * st.stride + i] = s.cooldown;
    st.ints[(int64_t)RL4_SPI_FLAGS * st.stride + i] = s.flags;
    st.ints[(int64_t)RL4_SPI_DIVERGED_STEP * st.stride + i] = s.diverged_step;
    st.ints[(int64_t)RL4_SPI_CONV_STEP * st.stride + i] = s.conv_step;
}

// One row of IDHPsp._log (objects.py:651-726), SoA over the logged agents.
template <typename TN, typename TE, int LOG, bool SM>
__device__ __forceinline__ void sp_write_log(const rl4_sp_log& lg, int64_t row, int64_t i, int k,
                                             const SpAgent<TN, TE, SM>& s, const SpStepOut<TN, TE>& o)
{
    const int nf = (LOG == RL4_LOG_FULL) ? RL4_LF_COUNT : RL4_LB_COUNT;
    double* b = lg.buf + (row * nf) * lg.n_agents_logged + i;
    const int64_t L = lg.n_agents_logged;
    auto put = [&](int f, double v) { b[(int64_t)f * L] = v; };
    put(RL4_LB_X + 0, (double)s.x[0].v); put(RL4_LB_X + 1, (double)s.x[1].v);
    put(RL4_LB_A, (double)s.a.v); put(RL4_LB_C, (double)o.cost.v);
    put(RL4_LB_REF, (double)o.ref.v); put(RL4_LB_E, (double)o.e.v);
    if (LOG == RL4_LOG_FULL) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            put(RL4_LF_AW1 + j, (double)s.W1a[j].v); put(RL4_LF_AW2 + j, (double)s.W2a[j].v);
            put(RL4_LF_CW1 + j, (double)s.W1c[j].v);
            put(RL4_LF_CE + j, (double)s.EcH[j].v); put(RL4_LF_CE + 4 + j, (double)s.EcR0[j].v); put(RL4_LF_CE + 8 + j, (double)s.EcR1[j].v);
            put(RL4_LF_M + j, (double)o.M[j].v);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { put(RL4_LF_CW2 + j, (double)s.W2c[j].v); put(RL4_LF_AE + j, (double)s.Ea[j].v); }
        const bool g = k > 1;                       // objects.py:704
#pragma unroll
        for (int j = 0; j < 4; ++j) {               // a_all_grad = [W1_update, W2_update] (objects.py:711-714)
            put(RL4_LF_AGRAD + j, g ? (double)o.dWa[4 + j].v : 0.0);
            put(RL4_LF_AGRAD + 4 + j, g ? (double)o.dWa[j].v : 0.0);
            put(RL4_LF_CGRAD + j, g ? (double)o.dWc[8 + j].v : 0.0);            // objects.py:717-720
            put(RL4_LF_CGRAD + 4 + j * 2 + 0, g ? (double)o.dWc[0 + j].v : 0.0);
            put(RL4_LF_CGRAD + 4 + j * 2 + 1, g ? (double)o.dWc[4 + j].v : 0.0);
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) put(RL4_LF_PARAMS + j, g ? (double)s.th[j].v : 0.0);
#pragma unroll
        for (int j = 0; j < 9; ++j) put(RL4_LF_COV + j, g ? (double)s.cv[j].v : 0.0);
        put(RL4_LF_EPS_NORM, g ? (double)s.epsn.v : 0.0);
        put(RL4_LF_EPS_ABS + 0, g ? (double)fabs((double)s.eps[0].v) : 0.0);
        put(RL4_LF_EPS_ABS + 1, g ? (double)fabs((double)s.eps[1].v) : 0.0);
#pragma unroll
        for (int j = 0; j < 2; ++j) { put(RL4_LF_LAM + j, (double)o.lam[j].v); put(RL4_LF_LAM_T + j, (double)o.lt[j].v); put(RL4_LF_TD + j, (double)o.td[j].v); }
        put(RL4_LF_DADZ, (double)o.dadz.v); put(RL4_LF_LOSS_GRAD, (double)o.loss_grad.v);
    }
}

template <typename TN, typename TE, bool TRACES, int LOG, bool PER_AGENT>
__global__ void __launch_bounds__(kBlock, (MinBlocks<TN, TE>::v))
sp_run_kernel(const __grid_constant__ rl4_sp_params p, const double* __restrict__ ref_base, int k0, int n_steps,
              const rl4_sp_state st, int64_t n_agents, const rl4_sp_log lg)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    constexpr bool SM = UseSmem<TN, TE>::v && (LOG == RL4_LOG_NONE);
    __shared__ Rn<TE> sm_e[SM ? 15 * kBlock : 1];
    __shared__ Rn<TN> sm_n[SM ? 12 * kBlock : 1];
    if (i >= n_agents) return;
    SpAgent<TN, TE, SM> s;
    if constexpr (SM) {
        s.th.p = sm_e + threadIdx.x; s.cv.p = sm_e + 6 * kBlock + threadIdx.x;
        s.W1t.p = sm_n + threadIdx.x; s.W2t.p = sm_n + 4 * kBlock + threadIdx.x;
    }
    sp_load<TN, TE, SM>(s, st, i, TRACES);
    const HpView<PER_AGENT> hv{p, i};
    SpStepOut<TN, TE> o;
    const bool logged = (LOG != RL4_LOG_NONE) && i < lg.n_agents_logged;
    int k = k0;
    for (; k < k0 + n_steps; ++k) {
        if (s.diverged_step >= 0) break;            // the reference left its loop (objects.py:991)
        sp_agent_step<TN, TE, TRACES, PER_AGENT, SM>(s, p, hv, k, __ldg(ref_base + k), o);
        if (!TRACES) {
            // Without eligibility traces E is this step's Jacobian, overwritten by every step and read only inside it: it
            // goes to the state planes from the last step an agent executes and is NOT carried across iterations
            // (20 values = 40 registers that would otherwise stay live for the store after the loop).
            if (k == k0 + n_steps - 1 || s.diverged_step >= 0) sp_store_traces<TN, TE, SM>(s, st, i);
        }
        if (LOG != RL4_LOG_NONE) {
            if (logged && (k - k0) % lg.every == 0) { s.refresh_epsn(); sp_write_log<TN, TE, LOG, SM>(lg, (k - k0) / lg.every, i, k, s, o); }
        }
    }
    if (LOG != RL4_LOG_NONE) {                      // NaN-fill the rows after a divergence (objects.py:656-679)
        if (logged) {
            const int nf = (LOG == RL4_LOG_FULL) ? RL4_LF_COUNT : RL4_LB_COUNT;
            for (; k < k0 + n_steps; ++k) {
                if ((k - k0) % lg.every) continue;
                double* b = lg.buf + ((int64_t)((k - k0) / lg.every) * nf) * lg.n_agents_logged + i;
                for (int f = 0; f < nf; ++f) b[(int64_t)f * lg.n_agents_logged] = __longlong_as_double(0x7ff8000000000000LL);
            }
        }
    }
    s.refresh_epsn();
    sp_store<TN, TE, SM, TRACES>(s, st, i);
}

// IDHPsp.__init__ + train() prologue (objects.py:552-615, 843-851, 911-948); env.reset (env.py:222-258)
template <typename TN, typename TE>
__global__ void __launch_bounds__(256)
sp_init_kernel(const __grid_constant__ rl4_sp_params p, const double* __restrict__ x0, const double* __restrict__ w1a,
               const double* __restrict__ w2a, const double* __restrict__ w1c, const double* __restrict__ w2c,
               int64_t stride_in, const rl4_sp_state st, int64_t n_agents)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_agents) return;
    const HpView<true> hv{p, i};
    SpAgent<TN, TE> s;
    using N = Rn<TN>;
    using E = Rn<TE>;
    const E ze = E(TE(0));
    const N zn = N(TN(0));
#pragma unroll
    for (int j = 0; j < 2; ++j) { s.x[j] = E(TE(x0[j * stride_in + i])); s.xp[j] = ze; s.eps[j] = ze; }
#pragma unroll
    for (int j = 0; j < 6; ++j) s.th[j] = ze;                            // objects.py:461
    const E c0 = E(TE(hv.hp(RL4_HP_RLS_COV0)));
#pragma unroll
    for (int j = 0; j < 9; ++j) s.cv[j] = (j % 4 == 0) ? c0 : ze;        // objects.py:470
    s.cgp = ze; s.epsn = ze; s.eps_fresh = false; s.sumc = ze; s.sumabse = ze;
#pragma unroll
    for (int j = 0; j < 8; ++j) s.Ea[j] = ze;
#pragma unroll
    for (int j = 0; j < 4; ++j) { s.EcH[j] = ze; s.EcR0[j] = ze; s.EcR1[j] = ze; }
    s.a = zn; s.ap = zn;                                                 // a0 = actor([[0]]) = 0 (objects.py:933)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        s.W1a[j] = N(TN(w1a[j * stride_in + i]));
        s.W2a[j] = -N(TN(w2a[j * stride_in + i]));                       // _invert_controller (objects.py:850; Q6)
        s.W1c[j] = N(TN(w1c[j * stride_in + i]));
        s.W1t[j] = s.W1c[j];                                             // soft_update(tau=1) (objects.py:926)
        s.Mp[j] = zn;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { s.W2c[j] = N(TN(w2c[j * stride_in + i])); s.W2t[j] = s.W2c[j]; }
    s.eta_a = N(TN(hv.hp(RL4_HP_ETA_A_H)));                              // objects.py:915-919
    s.eta_c = N(TN(hv.hp(RL4_HP_ETA_C_H)));
    s.cooldown = 0; s.flags = RL4_SPF_LR_INIT; s.diverged_step = -1; s.conv_step = -1;
    sp_store<TN, TE, false>(s, st, i);
}

// ---- step-API kernels ----------------------------------------------------------------
template <typename TN, typename TE>
__global__ void __launch_bounds__(256)
sp_env_step_kernel(const __grid_constant__ rl4_sp_params p, const double* __restrict__ ref_base, int stepp,
                   TE* __restrict__ x, const TN* __restrict__ action_deg, TE* __restrict__ out_reward,
                   TE* __restrict__ out_e, TE* __restrict__ out_rg0, int64_t stride, int64_t n_agents)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_agents) return;
    const HpView<true> hv{p, i};
    using E = Rn<TE>;
    const int fault_step = hv.hpi(RL4_HPI_FAULT_STEP);
    const int variant = (fault_step >= 0 && stepp >= fault_step) ? hv.hpi(RL4_HPI_FAULT_KIND) : 0;
    E xs[2] = {E(x[i]), E(x[stride + i])}, xn[2], e, cost, rg0;
    const E ref = E(TE(hv.hp(RL4_HP_REF_AMP))) * E(TE(__ldg(ref_base + stepp)));
    sp_env_step<TN, TE>(xs, Rn<TN>(action_deg[i]), ref, E(TE(hv.hp(RL4_HP_KAPPA))), E(TE(p.dt)), p.A[variant], p.B[variant],
                        hv.hpi(RL4_HPI_TRACKED_Q) != 0, e, cost, rg0, xn);
    x[i] = xn[0].v; x[stride + i] = xn[1].v;
    out_reward[i] = cost.v; out_e[i] = e.v; out_rg0[i] = rg0.v;
}

template <typename TE>
__global__ void __launch_bounds__(256)
sp_rls_update_kernel(const __grid_constant__ rl4_sp_params p, TE* __restrict__ theta, TE* __restrict__ cov,
                     const TE* __restrict__ dx0, const TE* __restrict__ da0, const TE* __restrict__ dx1,
                     TE* __restrict__ out_eps, TE* __restrict__ out_eps_norm, int64_t stride, int64_t n_agents)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_agents) return;
    const HpView<true> hv{p, i};
    using E = Rn<TE>;
    E th[6], cv[9], X[3], Y[2], eps[2];
#pragma unroll
    for (int j = 0; j < 6; ++j) th[j] = E(theta[j * stride + i]);
#pragma unroll
    for (int j = 0; j < 9; ++j) cv[j] = E(cov[j * stride + i]);
    X[0] = E(dx0[i]); X[1] = E(dx0[stride + i]); X[2] = E(da0[i]);
    Y[0] = E(dx1[i]); Y[1] = E(dx1[stride + i]);
    sp_rls_update<TE>(th, cv, X, Y, E(TE(hv.hp(RL4_HP_RLS_GAMMA))), eps);
    const E epsn = sqrt_rn(sp_eps_sq<TE>(eps));                         // objects.py:539
#pragma unroll
    for (int j = 0; j < 6; ++j) theta[j * stride + i] = th[j].v;
#pragma unroll
    for (int j = 0; j < 9; ++j) cov[j * stride + i] = cv[j].v;
    out_eps[i] = eps[0].v; out_eps[stride + i] = eps[1].v; out_eps_norm[i] = epsn.v;
}

template <typename TN, typename TE>
__global__ void __launch_bounds__(256)
sp_critic_forward_kernel(const TN* __restrict__ z, const TN* __restrict__ w1, const TN* __restrict__ w2, TE* __restrict__ Eplane,
                         TN* __restrict__ out_lambda, double gamma_lambda, int elig, int64_t stride, int64_t n_agents)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_agents) return;
    using N = Rn<TN>;
    using E = Rn<TE>;
    N W1[4], W2[8], lam[2];
    E EcH[4], EcR0[4], EcR1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { W1[j] = N(w1[j * stride + i]); EcH[j] = E(Eplane[j * stride + i]); EcR0[j] = E(Eplane[(4 + j) * stride + i]); EcR1[j] = E(Eplane[(8 + j) * stride + i]); }
#pragma unroll
    for (int j = 0; j < 8; ++j) W2[j] = N(w2[j * stride + i]);
    N h[4];
    sp_hidden(N(z[i]), W1, h);
    sp_critic_forward<TN, TE>(N(z[i]), h, W2, EcH, EcR0, EcR1, elig, E(TE(gamma_lambda)), lam);
#pragma unroll
    for (int j = 0; j < 4; ++j) { Eplane[j * stride + i] = EcH[j].v; Eplane[(4 + j) * stride + i] = EcR0[j].v; Eplane[(8 + j) * stride + i] = EcR1[j].v; }
    out_lambda[i] = lam[0].v; out_lambda[stride + i] = lam[1].v;
}

template <typename TN, typename TE>
__global__ void __launch_bounds__(256)
sp_actor_forward_kernel(const TN* __restrict__ z, const TN* __restrict__ w1, const TN* __restrict__ w2, TE* __restrict__ Eplane,
                        TN* __restrict__ out_a, TN* __restrict__ out_dadz, double gamma_lambda, int elig, int64_t stride,
                        int64_t n_agents)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_agents) return;
    using N = Rn<TN>;
    using E = Rn<TE>;
    N W1[4], W2[4], a, dadz;
    E Ea[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) { W1[j] = N(w1[j * stride + i]); W2[j] = N(w2[j * stride + i]); }
#pragma unroll
    for (int j = 0; j < 8; ++j) Ea[j] = E(Eplane[j * stride + i]);
    N h[4];
    sp_hidden(N(z[i]), W1, h);
    sp_actor_forward<TN, TE>(N(z[i]), h, W1, W2, Ea, elig, E(TE(gamma_lambda)), a, dadz);
#pragma unroll
    for (int j = 0; j < 8; ++j) Eplane[j * stride + i] = Ea[j].v;
    out_a[i] = a.v;
    if (out_dadz) out_dadz[i] = dadz.v;
}

// Critic.get_weight_update (objects.py:195-205): td @ E in the tensor dtype; out = [dW1 (4), dW2 (4,2) row-major]
template <typename TN, typename TE>
__global__ void __launch_bounds__(256)
sp_critic_weight_update_kernel(const TN* __restrict__ td, const TE* __restrict__ Eplane, TN* __restrict__ out,
                               int64_t stride, int64_t n_agents)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_agents) return;
    using N = Rn<TN>;
    const N td0 = N(td[i]), td1 = N(td[stride + i]), zero = N(TN(0));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const N eh = cvt<TN>(Rn<TE>(Eplane[j * stride + i]));
        const N e0 = cvt<TN>(Rn<TE>(Eplane[(4 + j) * stride + i])), e1 = cvt<TN>(Rn<TE>(Eplane[(8 + j) * stride + i]));
        out[j * stride + i] = fma(td1, e1, td0 * e0).v;                       // W1_update[0][j] = crit_grad[8+j]
        out[(4 + j * 2 + 0) * stride + i] = fma(td1, zero, td0 * eh).v;       // W2_update[j][0] = crit_grad[j]
        out[(4 + j * 2 + 1) * stride + i] = fma(td1, eh, td0 * zero).v;       // W2_update[j][1] = crit_grad[4+j]
    }
}

// ---- host-side dispatch ------------------------------------------------------------------
static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

template <typename TN, typename TE>
static int launch_run(const rl4_sp_params* p, const double* ref_base, int k0, int n_steps, rl4_sp_state st, int64_t n,
                      int use_traces, rl4_sp_log lg, cudaStream_t stream)
{
    bool per_agent = false;
    for (int j = 0; j < RL4_HP_COUNT; ++j) per_agent |= (p->hp_agent[j] != nullptr);
    for (int j = 0; j < RL4_HPI_COUNT; ++j) per_agent |= (p->hpi_agent[j] != nullptr);
    const unsigned grid = grid_for(n, kBlock);
    if (lg.level == RL4_LOG_BASIC)
        sp_run_kernel<TN, TE, true, RL4_LOG_BASIC, true><<<grid, kBlock, 0, stream>>>(*p, ref_base, k0, n_steps, st, n, lg);
    else if (lg.level == RL4_LOG_FULL)
        sp_run_kernel<TN, TE, true, RL4_LOG_FULL, true><<<grid, kBlock, 0, stream>>>(*p, ref_base, k0, n_steps, st, n, lg);
    else if (use_traces && per_agent)
        sp_run_kernel<TN, TE, true, RL4_LOG_NONE, true><<<grid, kBlock, 0, stream>>>(*p, ref_base, k0, n_steps, st, n, lg);
    else if (use_traces)
        sp_run_kernel<TN, TE, true, RL4_LOG_NONE, false><<<grid, kBlock, 0, stream>>>(*p, ref_base, k0, n_steps, st, n, lg);
    else if (per_agent)
        sp_run_kernel<TN, TE, false, RL4_LOG_NONE, true><<<grid, kBlock, 0, stream>>>(*p, ref_base, k0, n_steps, st, n, lg);
    else
        sp_run_kernel<TN, TE, false, RL4_LOG_NONE, false><<<grid, kBlock, 0, stream>>>(*p, ref_base, k0, n_steps, st, n, lg);
    return check_launch("sp_run_kernel");
}

static int check_state(const rl4_sp_state& st, int64_t n)
{
    if (!st.env || !st.net || !st.ints) { set_error("state planes must be non-NULL"); return -1; }
    if (st.stride < n) { set_error("state stride %lld < n_agents %lld", (long long)st.stride, (long long)n); return -1; }
    return 0;
}

}  // namespace rl4

using namespace rl4;

extern "C" {

int rl4_sp_init(int policy, const rl4_sp_params* p, const double* x0, const double* w1a, const double* w2a,
                const double* w1c, const double* w2c, int64_t stride_in, rl4_sp_state st, int64_t n, void* stream)
{
    RL4_REQUIRE(p && x0 && w1a && w2a && w1c && w2c, "NULL argument");
    RL4_REQUIRE(n >= 0 && stride_in >= n, "bad n_agents / stride_in");
    if (check_state(st, n)) return -1;
    if (n == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = grid_for(n, 256);
    switch (policy) {
    case RL4_FP64:  sp_init_kernel<double, double><<<grid, 256, 0, s>>>(*p, x0, w1a, w2a, w1c, w2c, stride_in, st, n); break;
    case RL4_FP32:  sp_init_kernel<float, float><<<grid, 256, 0, s>>>(*p, x0, w1a, w2a, w1c, w2c, stride_in, st, n); break;
    case RL4_MIXED: sp_init_kernel<float, double><<<grid, 256, 0, s>>>(*p, x0, w1a, w2a, w1c, w2c, stride_in, st, n); break;
    default: set_error("rl4_sp_init: unknown policy %d", policy); return -1;
    }
    return check_launch("sp_init_kernel");
}

int rl4_sp_run(int policy, const rl4_sp_params* p, const double* ref_base, int32_t k0, int32_t n_steps,
               rl4_sp_state st, int64_t n, int32_t use_traces, rl4_sp_log lg, void* stream)
{
    RL4_REQUIRE(p && ref_base, "NULL argument");
    RL4_REQUIRE(n >= 0 && k0 >= 0 && n_steps >= 0, "negative size");
    if (check_state(st, n)) return -1;
    if (lg.level != RL4_LOG_NONE) {
        RL4_REQUIRE(lg.level == RL4_LOG_BASIC || lg.level == RL4_LOG_FULL, "bad log level");
        RL4_REQUIRE(lg.buf && lg.every >= 1 && lg.n_agents_logged >= 0 && lg.n_agents_logged <= n, "bad log descriptor");
    }
    if (!use_traces && lg.level == RL4_LOG_NONE) {
        RL4_REQUIRE(p->hpi[RL4_HPI_ELIG_A] == RL4_ELIG_NONE && p->hpi[RL4_HPI_ELIG_C] == RL4_ELIG_NONE &&
                    !p->hpi_agent[RL4_HPI_ELIG_A] && !p->hpi_agent[RL4_HPI_ELIG_C],
                    "use_traces == 0 needs elig_a == elig_c == none");
    }
    if (n == 0 || n_steps == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    switch (policy) {
    case RL4_FP64:  return launch_run<double, double>(p, ref_base, k0, n_steps, st, n, use_traces, lg, s);
    case RL4_FP32:  return launch_run<float, float>(p, ref_base, k0, n_steps, st, n, use_traces, lg, s);
    case RL4_MIXED: return launch_run<float, double>(p, ref_base, k0, n_steps, st, n, use_traces, lg, s);
    default: set_error("rl4_sp_run: unknown policy %d", policy); return -1;
    }
}

int rl4_sp_env_step(int policy, const rl4_sp_params* p, const double* ref_base, int32_t stepp, void* x, const void* action_deg,
                    void* out_reward, void* out_e, void* out_rg0, int64_t stride, int64_t n, void* stream)
{
    RL4_REQUIRE(p && ref_base && x && action_deg && out_reward && out_e && out_rg0, "NULL argument");
    RL4_REQUIRE(n >= 0 && stride >= n && stepp >= 0, "bad size");
    if (n == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = grid_for(n, 256);
    switch (policy) {
    case RL4_FP64:  sp_env_step_kernel<double, double><<<grid, 256, 0, s>>>(*p, ref_base, stepp, (double*)x, (const double*)action_deg, (double*)out_reward, (double*)out_e, (double*)out_rg0, stride, n); break;
    case RL4_FP32:  sp_env_step_kernel<float, float><<<grid, 256, 0, s>>>(*p, ref_base, stepp, (float*)x, (const float*)action_deg, (float*)out_reward, (float*)out_e, (float*)out_rg0, stride, n); break;
    case RL4_MIXED: sp_env_step_kernel<float, double><<<grid, 256, 0, s>>>(*p, ref_base, stepp, (double*)x, (const float*)action_deg, (double*)out_reward, (double*)out_e, (double*)out_rg0, stride, n); break;
    default: set_error("rl4_sp_env_step: unknown policy %d", policy); return -1;
    }
    return check_launch("sp_env_step_kernel");
}

int rl4_sp_rls_update(int policy, const rl4_sp_params* p, void* theta, void* cov, const void* dx0, const void* da0,
                      const void* dx1, void* out_eps, void* out_eps_norm, int64_t stride, int64_t n, void* stream)
{
    RL4_REQUIRE(p && theta && cov && dx0 && da0 && dx1 && out_eps && out_eps_norm, "NULL argument");
    RL4_REQUIRE(n >= 0 && stride >= n, "bad size");
    if (n == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = grid_for(n, 256);
    if (policy == RL4_FP32)
        sp_rls_update_kernel<float><<<grid, 256, 0, s>>>(*p, (float*)theta, (float*)cov, (const float*)dx0, (const float*)da0, (const float*)dx1, (float*)out_eps, (float*)out_eps_norm, stride, n);
    else if (policy == RL4_FP64 || policy == RL4_MIXED)
        sp_rls_update_kernel<double><<<grid, 256, 0, s>>>(*p, (double*)theta, (double*)cov, (const double*)dx0, (const double*)da0, (const double*)dx1, (double*)out_eps, (double*)out_eps_norm, stride, n);
    else { set_error("rl4_sp_rls_update: unknown policy %d", policy); return -1; }
    return check_launch("sp_rls_update_kernel");
}

int rl4_sp_critic_forward(int policy, const void* z, const void* w1, const void* w2, void* E, void* out_lambda,
                          double gamma_lambda, int32_t elig, int64_t stride, int64_t n, void* stream)
{
    RL4_REQUIRE(z && w1 && w2 && E && out_lambda, "NULL argument");
    RL4_REQUIRE(n >= 0 && stride >= n, "bad size");
    RL4_REQUIRE(elig >= RL4_ELIG_NONE && elig <= RL4_ELIG_REPLACING, "bad elig");
    if (n == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = grid_for(n, 256);
    switch (policy) {
    case RL4_FP64:  sp_critic_forward_kernel<double, double><<<grid, 256, 0, s>>>((const double*)z, (const double*)w1, (const double*)w2, (double*)E, (double*)out_lambda, gamma_lambda, elig, stride, n); break;
    case RL4_FP32:  sp_critic_forward_kernel<float, float><<<grid, 256, 0, s>>>((const float*)z, (const float*)w1, (const float*)w2, (float*)E, (float*)out_lambda, gamma_lambda, elig, stride, n); break;
    case RL4_MIXED: sp_critic_forward_kernel<float, double><<<grid, 256, 0, s>>>((const float*)z, (const float*)w1, (const float*)w2, (double*)E, (float*)out_lambda, gamma_lambda, elig, stride, n); break;
    default: set_error("rl4_sp_critic_forward: unknown policy %d", policy); return -1;
    }
    return check_launch("sp_critic_forward_kernel");
}

int rl4_sp_actor_forward(int policy, const void* z, const void* w1, const void* w2, void* E, void* out_a, void* out_dadz,
                         double gamma_lambda, int32_t elig, int64_t stride, int64_t n, void* stream)
{
    RL4_REQUIRE(z && w1 && w2 && E && out_a, "NULL argument");
    RL4_REQUIRE(n >= 0 && stride >= n, "bad size");
    RL4_REQUIRE(elig >= RL4_ELIG_NONE && elig <= RL4_ELIG_REPLACING, "bad elig");
    if (n == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = grid_for(n, 256);
    switch (policy) {
    case RL4_FP64:  sp_actor_forward_kernel<double, double><<<grid, 256, 0, s>>>((const double*)z, (const double*)w1, (const double*)w2, (double*)E, (double*)out_a, (double*)out_dadz, gamma_lambda, elig, stride, n); break;
    case RL4_FP32:  sp_actor_forward_kernel<float, float><<<grid, 256, 0, s>>>((const float*)z, (const float*)w1, (const float*)w2, (float*)E, (float*)out_a, (float*)out_dadz, gamma_lambda, elig, stride, n); break;
    case RL4_MIXED: sp_actor_forward_kernel<float, double><<<grid, 256, 0, s>>>((const float*)z, (const float*)w1, (const float*)w2, (double*)E, (float*)out_a, (float*)out_dadz, gamma_lambda, elig, stride, n); break;
    default: set_error("rl4_sp_actor_forward: unknown policy %d", policy); return -1;
    }
    return check_launch("sp_actor_forward_kernel");
}

int rl4_sp_critic_weight_update(int policy, const void* td, const void* E, void* out, int64_t stride, int64_t n, void* stream)
{
    RL4_REQUIRE(td && E && out, "NULL argument");
    RL4_REQUIRE(n >= 0 && stride >= n, "bad size");
    if (n == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = grid_for(n, 256);
    switch (policy) {
    case RL4_FP64:  sp_critic_weight_update_kernel<double, double><<<grid, 256, 0, s>>>((const double*)td, (const double*)E, (double*)out, stride, n); break;
    case RL4_FP32:  sp_critic_weight_update_kernel<float, float><<<grid, 256, 0, s>>>((const float*)td, (const float*)E, (float*)out, stride, n); break;
    case RL4_MIXED: sp_critic_weight_update_kernel<float, double><<<grid, 256, 0, s>>>((const float*)td, (const double*)E, (float*)out, stride, n); break;
    default: set_error("rl4_sp_critic_weight_update: unknown policy %d", policy); return -1;
    }
    return check_launch("sp_critic_weight_update_kernel");
}

}  // extern "C"
