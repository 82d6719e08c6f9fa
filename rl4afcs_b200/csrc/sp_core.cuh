// Per-agent device functions of the short-period IDHP path (sm_100a).
//
// One thread owns one aircraft+agent instance; everything below works on registers.
// Each function states the reference code it replaces (wingos80/RL4AFCS file:line) and
// follows the arithmetic contract of DESIGN.md: numpy-side `@` = FMA chain in the term
// order numpy executed (A@x: 1,0; Cov@X: 1,0,2; others in order), TensorFlow-side `@` =
// in-order FMA chain, every other operation individually rounded (Rn<T>).
#pragma once
#include "rl4_math.cuh"
#include "../../include/rl4afcs_b200.h"

namespace rl4 {

// ---- hyper-parameter access: shared scalars straight from the kernel-parameter constant
// bank, or per-agent override arrays (hyper-parameter sweeps / Monte-Carlo fault studies).
template <bool PER_AGENT>
struct HpView {
    const rl4_sp_params& p;
    int64_t i;
    __device__ __forceinline__ double hp(int idx) const {
        if (PER_AGENT) { const double* a = p.hp_agent[idx]; if (a) return __ldg(a + i); }
        return p.hp[idx];
    }
    __device__ __forceinline__ int hpi(int idx) const {
        if (PER_AGENT) { const int32_t* a = p.hpi_agent[idx]; if (a) return __ldg(a + i); }
        return p.hpi[idx];
    }
};

#ifndef RL4_SP_BLOCK
#define RL4_SP_BLOCK 128
#endif

// Small per-agent array that lives either in registers or in shared memory laid out [element][thread]
// (conflict-free, static element index -> immediate offset).  The fp64 fused kernel keeps the parts of the state
// that are touched only once or twice per step (target-critic weights, RLS parameters and covariance) in shared
// memory so that 3 CTAs per SM fit the register file.
template <typename T, int N, bool SM> struct Arr;
template <typename T, int N> struct Arr<T, N, false> {
    Rn<T> a[N];
    __device__ __forceinline__ Rn<T>& operator[](int j) { return a[j]; }
    __device__ __forceinline__ const Rn<T>& operator[](int j) const { return a[j]; }
};
template <typename T, int N> struct Arr<T, N, true> {
    Rn<T>* p;
    __device__ __forceinline__ Rn<T>& operator[](int j) const { return p[j * RL4_SP_BLOCK]; }
};

// ||eps||^2 as np.linalg.norm forms it before the square root (ddot: in-order FMA chain; objects.py:539)
template <typename TE>
__device__ __forceinline__ Rn<TE> sp_eps_sq(const Rn<TE> (&eps)[2]) { return fma(eps[1], eps[1], eps[0] * eps[0]); }

template <typename TN, typename TE, bool SM = false>
struct SpAgent {
    using N = Rn<TN>;
    using E = Rn<TE>;
    // env plane
    E x[2], xp[2], cgp, eps[2], epsn, sumc, sumabse;
    Arr<TE, 6, SM> th;
    Arr<TE, 9, SM> cv;
    E Ea[8], EcH[4], EcR0[4], EcR1[4];
    // net plane
    N a, ap, W1a[4], W2a[4], W1c[4], W2c[8], Mp[4], eta_a, eta_c;
    Arr<TN, 4, SM> W1t;
    Arr<TN, 8, SM> W2t;
    // int plane
    int cooldown, flags, diverged_step, conv_step;
    bool eps_fresh;                 // eps was updated since epsn was last formed (the square root is taken lazily)
    __device__ __forceinline__ void refresh_epsn() { if (eps_fresh) { epsn = sqrt_rn(sp_eps_sq<TE>(eps)); eps_fresh = false; } }
};

// per-step scratch that the log and the step-API kernels want to see
template <typename TN, typename TE>
struct SpStepOut {
    Rn<TE> e, cost, ref, rg0;
    Rn<TN> lam[2], lt[2], td[2], dadz, M[4], loss_grad, dWa[8], dWc[12];
};

// ---------------------------------------------------------------------------------------
// Ce500ShortPeriod.step  envs/linear/env.py:176-199
//   u = deg2rad(action) in the action dtype (float32 in the reference, Q5), e = ref - alpha
//   BEFORE integration (Q12), reward -0.5 kappa e^2, reward_grad = kappa*[-2e, 0] (Q4),
//   forward Euler  x += (A@x + B*u)*dt.
// ---------------------------------------------------------------------------------------
template <typename TN, typename TE>
__device__ __forceinline__ void sp_env_step(Rn<TE> (&x)[2], Rn<TN> action_deg, Rn<TE> ref, Rn<TE> kappa, Rn<TE> dt,
                                            const double* __restrict__ A, const double* __restrict__ B, bool tracked_q,
                                            Rn<TE>& e, Rn<TE>& cost, Rn<TE>& rg0, Rn<TE> (&xn)[2])
{
    using E = Rn<TE>;
    const Rn<TN> u_n = action_deg * Rn<TN>(Consts<TN>::deg2rad());   // env.py:176
    const E u = cvt<TE>(u_n);
    // env.py:179  y = C@x + D*action with C = I, D = 0; env.py:180-184: error = ref - y[0] ('alpha') or ref - y[1] ('q')
    const E xt = tracked_q ? x[1] : x[0], xo = tracked_q ? x[0] : x[1];   // tracked / other state: one expression, two selects
    const E y = fma(E(TE(0)), xo, xt) + E(TE(0)) * u;                 // 1 * xt == xt bit for bit (the product by one is exact)
    e = ref - y;
    cost = (E(TE(-0.5)) * kappa) * (e * e);                           // env.py:187
    rg0 = kappa * (E(TE(-2)) * e);                                    // env.py:189
    const E A00 = E(TE(A[0])), A01 = E(TE(A[1])), A10 = E(TE(A[2])), A11 = E(TE(A[3]));
    const E B0 = E(TE(B[0])), B1 = E(TE(B[1]));
    const E xd0 = fma(A00, x[0], A01 * x[1]) + B0 * u;                // env.py:192 (numpy order 1,0)
    const E xd1 = fma(A10, x[0], A11 * x[1]) + B1 * u;
    xn[0] = x[0] + xd0 * dt;                                          // env.py:193
    xn[1] = x[1] + xd1 * dt;
}

// Network.base_call hidden layers (objects.py:111-139) of critic, target critic and actor for the
// scalar input z (Q1): h_j = tanh(z*W1_j), ai0_j = 1 - h_j^2.  The 12 tanh are one group so that
// their dependent chains interleave (the three nets are independent).
template <typename TN, typename W1T>
__device__ __forceinline__ void sp_hidden3(Rn<TN> z, const Rn<TN> (&W1c)[4], const W1T& W1t, const Rn<TN> (&W1a)[4],
                                           Rn<TN> (&hc)[4], Rn<TN> (&ht)[4], Rn<TN> (&ha)[4])
{
#ifndef RL4_SP_TANH_GROUP
#define RL4_SP_TANH_GROUP 12
#endif
    Rn<TN> pre[12], h[12];
#pragma unroll
    for (int j = 0; j < 4; ++j) { pre[j] = z * W1c[j]; pre[4 + j] = z * W1t[j]; pre[8 + j] = z * W1a[j]; }
#if RL4_SP_TANH_GROUP == 12
    tanh_t13_n<12>(pre, h);
#else
    // smaller groups: fewer interleaved chains, fewer live temporaries
#pragma unroll
    for (int g = 0; g < 12; g += RL4_SP_TANH_GROUP) {
        Rn<TN> pg[RL4_SP_TANH_GROUP], hg[RL4_SP_TANH_GROUP];
#pragma unroll
        for (int j = 0; j < RL4_SP_TANH_GROUP; ++j) pg[j] = pre[g + j];
        tanh_t13_n<RL4_SP_TANH_GROUP>(pg, hg);
#pragma unroll
        for (int j = 0; j < RL4_SP_TANH_GROUP; ++j) h[g + j] = hg[j];
    }
#endif
#pragma unroll
    for (int j = 0; j < 4; ++j) { hc[j] = h[j]; ht[j] = h[4 + j]; ha[j] = h[8 + j]; }
}
template <typename TN>
__device__ __forceinline__ void sp_hidden(Rn<TN> z, const Rn<TN> (&W1)[4], Rn<TN> (&h)[4])
{
    Rn<TN> pre[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) pre[j] = z * W1[j];
    tanh_t13_n<4>(pre, h);
}
template <typename TN>
__device__ __forceinline__ void sp_ai0(const Rn<TN> (&h)[4], Rn<TN> (&ai0)[4])
{
#pragma unroll
    for (int j = 0; j < 4; ++j) ai0[j] = Rn<TN>(TN(1)) - h[j] * h[j];
}

// (1,4)@(4,2) linear output layer, in-order FMA chain
template <typename TN, typename W2T>
__device__ __forceinline__ void sp_out2(const Rn<TN> (&h)[4], const W2T& W2, Rn<TN> (&out)[2])
{
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        Rn<TN> acc = h[0] * W2[q];
#pragma unroll
        for (int j = 1; j < 4; ++j) acc = fma(h[j], W2[j * 2 + q], acc);
        out[q] = acc;
    }
}

// trace update rule shared by actor and critic (objects.py:161-188, 236-254)
//   None: E <- g;  accumulating: E <- gl*E + g;  replacing handled by the callers (needs norms)
template <typename TE>
__device__ __forceinline__ Rn<TE> trace_none_or_acc(int elig, Rn<TE> Eold, Rn<TE> g, Rn<TE> gl)
{
    if (elig == RL4_ELIG_ACCUMULATING) return Eold * gl + g;
    return g;
}

// Critic.call (objects.py:151-193): forward + Jacobian trace.
// E layout: EcH = E[0,0:4] (== E[1,4:8]), EcR0 = E[0,8:12], EcR1 = E[1,8:12]; other slots are 0.
template <typename TN, typename TE>
__device__ __forceinline__ void sp_critic_forward(Rn<TN> z, const Rn<TN> (&h)[4], const Rn<TN> (&W2)[8],
                                                  Rn<TE> (&EcH)[4], Rn<TE> (&EcR0)[4], Rn<TE> (&EcR1)[4],
                                                  int elig, Rn<TE> gl, Rn<TN> (&lam)[2])
{
    using E = Rn<TE>;
    Rn<TN> ai0[4];
    sp_ai0(h, ai0);
    sp_out2(h, W2, lam);
    E gH[4], g0[4], g1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        gH[j] = cvt<TE>(h[j]);                                   // objects.py:162-165
        g0[j] = cvt<TE>((W2[j * 2 + 0] * ai0[j]) * z);           // objects.py:164
        g1[j] = cvt<TE>((W2[j * 2 + 1] * ai0[j]) * z);           // objects.py:166
    }
    if (elig == RL4_ELIG_REPLACING) {                            // objects.py:177-188
        E ng = E(TE(0)), ne = E(TE(0));
        // np.linalg.norm over the raveled (2,12) matrix; zero slots contribute fma(0,0,acc) = acc
#pragma unroll
        for (int j = 0; j < 4; ++j) { ng = fma(gH[j], gH[j], ng); ne = fma(EcH[j], EcH[j], ne); }
#pragma unroll
        for (int j = 0; j < 4; ++j) { ng = fma(g0[j], g0[j], ng); ne = fma(EcR0[j], EcR0[j], ne); }
#pragma unroll
        for (int j = 0; j < 4; ++j) { ng = fma(gH[j], gH[j], ng); ne = fma(EcH[j], EcH[j], ne); }
#pragma unroll
        for (int j = 0; j < 4; ++j) { ng = fma(g1[j], g1[j], ng); ne = fma(EcR1[j], EcR1[j], ne); }
        const bool take = sqrt_rn(ng).v > sqrt_rn(ne).v;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            EcH[j]  = take ? gH[j] : EcH[j] * gl;
            EcR0[j] = take ? g0[j] : EcR0[j] * gl;
            EcR1[j] = take ? g1[j] : EcR1[j] * gl;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            EcH[j]  = trace_none_or_acc(elig, EcH[j], gH[j], gl);
            EcR0[j] = trace_none_or_acc(elig, EcR0[j], g0[j], gl);
            EcR1[j] = trace_none_or_acc(elig, EcR1[j], g1[j], gl);
        }
    }
}

// Actor.call (objects.py:226-259) + d a / d z by reverse-mode autodiff in TF's order
// (objects.py:876-878): TanhGrad dy*(1-y*y), MatMul grad, TanhGrad, MatMul grad (chain).
template <typename TN, typename TE>
__device__ __forceinline__ void sp_actor_forward(Rn<TN> z, const Rn<TN> (&h)[4], const Rn<TN> (&W1)[4], const Rn<TN> (&W2)[4],
                                                 Rn<TE> (&Ea)[8], int elig, Rn<TE> gl, Rn<TN>& a_next, Rn<TN>& dadz)
{
    using N = Rn<TN>;
    using E = Rn<TE>;
    N ai0[4];
    sp_ai0(h, ai0);
    N o = h[0] * W2[0];
#pragma unroll
    for (int j = 1; j < 4; ++j) o = fma(h[j], W2[j], o);
    a_next = tanh_t13(o);
    const N ai1 = N(TN(1)) - a_next * a_next;
    E g[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        g[j]     = cvt<TE>(ai1 * h[j]);                          // objects.py:237
        g[4 + j] = cvt<TE>(((ai1 * W2[j]) * ai0[j]) * z);        // objects.py:238
    }
    if (elig == RL4_ELIG_REPLACING) {                            // objects.py:246-254
        E ng = E(TE(0)), ne = E(TE(0));
#pragma unroll
        for (int j = 0; j < 8; ++j) { ng = fma(g[j], g[j], ng); ne = fma(Ea[j], Ea[j], ne); }
        const bool take = sqrt_rn(ng).v > sqrt_rn(ne).v;
#pragma unroll
        for (int j = 0; j < 8; ++j) Ea[j] = take ? g[j] : Ea[j] * gl;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) Ea[j] = trace_none_or_acc(elig, Ea[j], g[j], gl);
    }
    const N g_o = ai1;                                            // dy = 1: 1 * ai1 == ai1 bit for bit
    N acc = ((g_o * W2[0]) * ai0[0]) * W1[0];
#pragma unroll
    for (int j = 1; j < 4; ++j) acc = fma((g_o * W2[j]) * ai0[j], W1[j], acc);
    dadz = acc;
}

// RLS.update (objects.py:492-543).  X = [dx0; da0], Y = dx1.
template <typename TE, typename TH, typename CV>
__device__ __forceinline__ void sp_rls_update(TH& th, CV& cv, const Rn<TE> (&X)[3], const Rn<TE> (&Y)[2],
                                              Rn<TE> rls_gamma, Rn<TE> (&eps)[2])
{
    using E = Rn<TE>;
    E CX[3], K[3];
#pragma unroll
    for (int i = 0; i < 2; ++i) {                                // params.T @ X  (objects.py:515; in order)
        E acc = th[0 * 2 + i] * X[0];
        acc = fma(th[1 * 2 + i], X[1], acc);
        acc = fma(th[2 * 2 + i], X[2], acc);
        eps[i] = Y[i] - acc;                                     // objects.py:516
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {                                // Cov @ X (objects.py:519; numpy order 1,0,2)
        E acc = cv[i * 3 + 1] * X[1];
        acc = fma(cv[i * 3 + 0], X[0], acc);
        acc = fma(cv[i * 3 + 2], X[2], acc);
        CX[i] = acc;
    }
    E xcx = X[0] * CX[0];                                        // objects.py:520
    xcx = fma(X[1], CX[1], xcx);
    xcx = fma(X[2], CX[2], xcx);
    const E den = rls_gamma + xcx;
    div_group<3>(CX, den, K);                                    // objects.py:521
#pragma unroll
    for (int j = 0; j < 3; ++j) {                                // objects.py:522
        th[j * 2 + 0] = th[j * 2 + 0] + K[j] * eps[0];
        th[j * 2 + 1] = th[j * 2 + 1] + K[j] * eps[1];
    }
    if (rls_gamma.v == TE(1)) {                                  // x / 1 == x exactly: skip the nine divisions
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) cv[i * 3 + j] = cv[i * 3 + j] - K[i] * CX[j];
    } else {
        E num[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) num[i * 3 + j] = cv[i * 3 + j] - K[i] * CX[j];
        E quo[9];
        div_group<9>(num, rls_gamma, quo);                       // objects.py:529-530
#pragma unroll
        for (int j = 0; j < 9; ++j) cv[j] = quo[j];
    }
    // objects.py:539: eps_norm = sqrt(sp_eps_sq(eps)) is formed by the callers, only where its value is needed
}

// Predicate `sqrt(v) > c` (c > 0) without the square root: sqrt is monotone and rounds once (relative error 2^-53 /
// 2^-24), so outside a guard band of 1e-12 (double) / 1e-5 (float) around c^2 the comparison of v with c^2 decides;
// inside the band the literal expression is evaluated.  Same truth value as the literal predicate for every input.
template <typename TE> struct Guard;
template <> struct Guard<double> { static __device__ __forceinline__ double lo() { return 1.0 - 1e-12; } static __device__ __forceinline__ double hi() { return 1.0 + 1e-12; } };
template <> struct Guard<float>  { static __device__ __forceinline__ float lo() { return 1.0f - 1e-5f; } static __device__ __forceinline__ float hi() { return 1.0f + 1e-5f; } };
template <typename TE>
__device__ __forceinline__ bool sqrt_gt(Rn<TE> v, TE c)
{
    const TE c2 = c * c;
    if (v.v > c2 * Guard<TE>::hi()) return true;
    if (!(v.v >= c2 * Guard<TE>::lo())) return false;        // below the band, or NaN (NaN > c is false)
    return sqrt_rn(v).v > c;
}

// ---------------------------------------------------------------------------------------
// One iteration of IDHPsp.train()'s loop body (objects.py:950-992) for one agent.
// ---------------------------------------------------------------------------------------
template <typename TN, typename TE, bool TRACES, bool PER_AGENT, bool SM>
__device__ __forceinline__ void sp_agent_step(SpAgent<TN, TE, SM>& s, const rl4_sp_params& p, const HpView<PER_AGENT>& hv,
                                              int k, double ref_base_k, SpStepOut<TN, TE>& o)
{
    using N = Rn<TN>;
    using E = Rn<TE>;
    const E kappa = E(TE(hv.hp(RL4_HP_KAPPA)));
    const E dt = E(TE(p.dt));

    // plant variant: the fault engages in the call where stepp becomes fault_step, i.e. it
    // first shapes the dynamics of step k == fault_step (env.py:128,198-199; Q11)
    const int fault_step = hv.hpi(RL4_HPI_FAULT_STEP);
    const int variant = (fault_step >= 0 && k >= fault_step) ? hv.hpi(RL4_HPI_FAULT_KIND) : 0;

    // ---- env.step(20*a)  (objects.py:955)
    const N a_k = s.a;
    o.ref = E(TE(hv.hp(RL4_HP_REF_AMP))) * E(TE(ref_base_k));           // idhp_sp.py:44,174
    E xn[2];
    sp_env_step<TN, TE>(s.x, N(TN(20)) * a_k, o.ref, kappa, dt, p.A[variant], p.B[variant], hv.hpi(RL4_HPI_TRACKED_Q) != 0,
                        o.e, o.cost, o.rg0, xn);
    const E rg1 = kappa * E(TE(0));

    // ---- _step_networks (objects.py:853-882): critic, target critic, actor on z = [[e]]
    const N z = cvt<TN>(o.e);                                           // objects.py:769
    const int elig_c = TRACES ? hv.hpi(RL4_HPI_ELIG_C) : RL4_ELIG_NONE;
    const int elig_a = TRACES ? hv.hpi(RL4_HPI_ELIG_A) : RL4_ELIG_NONE;
    E gl = E(TE(0));
    if (TRACES) {
        const double lam = (s.flags & RL4_SPF_LAMBDA_LOW) ? hv.hp(RL4_HP_LAMBDA_L) : hv.hp(RL4_HP_LAMBDA_H);
        gl = E(TE(__dmul_rn(lam, hv.hp(RL4_HP_GAMMA))));                // objects.py:602,814-817
    }
    N hc[4], ht[4], ha[4];
    sp_hidden3(z, s.W1c, s.W1t, s.W1a, hc, ht, ha);
    sp_critic_forward<TN, TE>(z, hc, s.W2c, s.EcH, s.EcR0, s.EcR1, elig_c, gl, o.lam);
    sp_out2(ht, s.W2t, o.lt);                                           // objects.py:867 (trace unused, Q16)
    N a_next;
    sp_actor_forward<TN, TE>(z, ha, s.W1a, s.W2a, s.Ea, elig_a, gl, a_next, o.dadz);

    // F, G of the RLS model BEFORE this step's update (objects.py:874; Q18), in the tensor dtype;
    // dx1dx0 = F + G@dadx with the (2,1) product broadcast over both columns (objects.py:963-964; Q2)
    N Gn[2];
    Gn[0] = cvt<TN>(s.th[4]); Gn[1] = cvt<TN>(s.th[5]);
    {
        const N gd0 = Gn[0] * o.dadz, gd1 = Gn[1] * o.dadz;
        o.M[0] = cvt<TN>(s.th[0]) + gd0;   // F00 = theta[0][0]
        o.M[1] = cvt<TN>(s.th[2]) + gd0;   // F01 = theta[1][0]
        o.M[2] = cvt<TN>(s.th[1]) + gd1;   // F10 = theta[0][1]
        o.M[3] = cvt<TN>(s.th[3]) + gd1;   // F11 = theta[1][1]
    }

    o.td[0] = N(TN(0)); o.td[1] = N(TN(0)); o.loss_grad = N(TN(0));
#pragma unroll
    for (int j = 0; j < 12; ++j) o.dWc[j] = N(TN(0));
#pragma unroll
    for (int j = 0; j < 8; ++j) o.dWa[j] = N(TN(0));

    if (k > 0) {
        // ---- _update_networks (objects.py:884-908)
        const double gamma_d = hv.hp(RL4_HP_GAMMA);
        const N gam = N(TN(gamma_d));
        if (hv.hpi(RL4_HPI_MULTISTEP)) {                                // objects.py:887 (Q14)
            const N gc0 = cvt<TN>(E(TE(gamma_d)) * o.rg0), gc1 = cvt<TN>(E(TE(gamma_d)) * rg1);
            const N T1_0 = fma(gc1, s.Mp[2], gc0 * s.Mp[0]), T1_1 = fma(gc1, s.Mp[3], gc0 * s.Mp[1]);
            const N g2 = N(TN(hv.hp(RL4_HP_GAMMA_SQ)));
            const N gl0 = g2 * o.lt[0], gl1 = g2 * o.lt[1];
            const N V0 = fma(gl1, s.Mp[2], gl0 * s.Mp[0]), V1 = fma(gl1, s.Mp[3], gl0 * s.Mp[1]);
            const N T2_0 = fma(V1, o.M[2], V0 * o.M[0]), T2_1 = fma(V1, o.M[3], V0 * o.M[1]);
            o.td[0] = ((o.lam[0] - cvt<TN>(s.cgp)) - T1_0) - T2_0;
            o.td[1] = ((o.lam[1] - cvt<TN>(rg1)) - T1_1) - T2_1;        // reward_grad[1] is kappa*0 every step
        } else {                                                        // objects.py:889
            const N gl0 = gam * o.lt[0], gl1 = gam * o.lt[1];
            const N T0 = fma(gl1, o.M[2], gl0 * o.M[0]), T1 = fma(gl1, o.M[3], gl0 * o.M[1]);
            o.td[0] = (o.lam[0] - cvt<TN>(o.rg0)) - T0;
            o.td[1] = (o.lam[1] - cvt<TN>(rg1)) - T1;
        }
        // critic.get_weight_update: td @ E with E cast to the tensor dtype (objects.py:201-205; Q15);
        // the structurally zero slots of E are kept in the chain so non-finite td behaves identically
        const N zero = N(TN(0));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const N eh = cvt<TN>(s.EcH[j]);
            o.dWc[0 + j] = fma(o.td[1], zero, o.td[0] * eh);
            o.dWc[4 + j] = fma(o.td[1], eh, o.td[0] * zero);
            o.dWc[8 + j] = fma(o.td[1], cvt<TN>(s.EcR1[j]), o.td[0] * cvt<TN>(s.EcR0[j]));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {                                   // SGD w -= lr*g (objects.py:892)
            s.W2c[j * 2 + 0] = s.W2c[j * 2 + 0] - s.eta_c * o.dWc[0 + j];
            s.W2c[j * 2 + 1] = s.W2c[j * 2 + 1] - s.eta_c * o.dWc[4 + j];
            s.W1c[j]         = s.W1c[j]         - s.eta_c * o.dWc[8 + j];
        }
        {   // target soft update with the just-updated critic (objects.py:895,207-215; Q17)
            const double tau_d = hv.hp(RL4_HP_TAU);
            const N omt = N(TN(__dsub_rn(1.0, tau_d))), tt = N(TN(tau_d));
#pragma unroll
            for (int j = 0; j < 4; ++j) s.W1t[j] = omt * s.W1t[j] + tt * s.W1c[j];
#pragma unroll
            for (int j = 0; j < 8; ++j) s.W2t[j] = omt * s.W2t[j] + tt * s.W2c[j];
        }
        {   // actor (objects.py:904-908): stale lambda', pre-update G, no minus sign (Q13)
            const N v0 = cvt<TN>(o.rg0) + gam * o.lt[0], v1 = cvt<TN>(rg1) + gam * o.lt[1];
            o.loss_grad = fma(v1, Gn[1], v0 * Gn[0]);
#pragma unroll
            for (int n = 0; n < 8; ++n) o.dWa[n] = o.loss_grad * cvt<TN>(s.Ea[n]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s.W2a[j] = s.W2a[j] - s.eta_a * o.dWa[j];
                s.W1a[j] = s.W1a[j] - s.eta_a * o.dWa[4 + j];
            }
        }
        // ---- RLS update (objects.py:972-975)
        {
            E X[3], Y[2];
            if (k == 1 && p.q3_alias) {     // Q3: x_prev aliases the live env array at k == 1
                X[0] = s.x[0] - xn[0]; X[1] = s.x[1] - xn[1];
            } else {
                X[0] = s.x[0] - s.xp[0]; X[1] = s.x[1] - s.xp[1];
            }
            X[2] = cvt<TE>(a_k - s.ap);                                 // objects.py:973 (tensor dtype)
            Y[0] = xn[0] - s.x[0]; Y[1] = xn[1] - s.x[1];
            sp_rls_update<TE>(s.th, s.cv, X, Y, E(TE(hv.hp(RL4_HP_RLS_GAMMA))), s.eps);
            s.eps_fresh = true;                                         // s.epsn is re-formed from s.eps when it is read
        }
        // ---- _adapt_check (objects.py:783-841)
        {
            const E thr = E(TE(hv.hp(RL4_HP_ERROR_THRESH_DEG))) * E(Consts<TE>::deg2rad());
            const bool cond1 = k < hv.hpi(RL4_HPI_WARMUP_STEPS);
            const bool cond2 = abs_rn(o.e).v > thr.v;
            const bool cond3 = sqrt_gt<TE>(sp_eps_sq<TE>(s.eps), TE(5e-5));    // rls.eps_norm > 5e-5 (objects.py:539,797)
            if (s.cooldown > 0) s.cooldown -= 1;
            const bool high = cond1 || cond2;
            const double eta_a_h = hv.hp(RL4_HP_ETA_A_H);
            const N des_a = N(TN(high ? eta_a_h : hv.hp(RL4_HP_ETA_A_L)));
            bool neq;
            if ((s.flags & RL4_SPF_LR_INIT) && p.q7_numpy1) neq = ((double)des_a.v != eta_a_h);   // Q7
            else neq = (des_a.v != s.eta_a.v);
            if (neq && s.cooldown == 0) {
                s.eta_a = des_a;
                s.eta_c = N(TN(high ? hv.hp(RL4_HP_ETA_C_H) : hv.hp(RL4_HP_ETA_C_L)));
                s.flags = (s.flags & ~(RL4_SPF_LAMBDA_LOW | RL4_SPF_LR_INIT)) | (high ? 0 : RL4_SPF_LAMBDA_LOW);
                s.cooldown = hv.hpi(RL4_HPI_COOLDOWN_STEPS);
            }
            if (cond3 && !(s.flags & RL4_SPF_CHANGED) && !cond1) {      // one-shot RLS reset (Q8)
                const E c0 = E(TE(hv.hp(RL4_HP_RLS_COV0)));
#pragma unroll
                for (int j = 0; j < 6; ++j) s.th[j] = E(TE(0));
#pragma unroll
                for (int j = 0; j < 9; ++j) s.cv[j] = (j % 4 == 0) ? c0 : E(TE(0));
                s.flags |= RL4_SPF_CHANGED;
            }
        }
    }

    // ---- shift (objects.py:980-985)
    s.xp[0] = s.x[0]; s.xp[1] = s.x[1];
    s.x[0] = xn[0]; s.x[1] = xn[1];
    s.ap = a_k; s.a = a_next;
    s.cgp = o.rg0;
#pragma unroll
    for (int j = 0; j < 4; ++j) s.Mp[j] = o.M[j];

    // ---- episode statistics (functions.py:39-60, utils.py:350-369)
    s.sumc = s.sumc + o.cost;
    s.sumabse = s.sumabse + abs_rn(o.e);
    {
        // utils.py:350-369 on this step's reward: sqrt(-2 c / kappa) * rad2deg > 0.5.  The literal expression is |e| * rad2deg
        // up to five roundings (c = (-0.5 kappa)(e e)), so outside a guard band around 0.5 deg the comparison of |e| with
        // the threshold in radians decides; inside the band (and for a kappa outside the range in which that error
        // analysis holds) the literal expression is evaluated.  Same truth value as the literal predicate.
        const TE ae = abs_rn(o.e).v;
        const TE thr = TE(0.5) / Consts<TE>::rad2deg();
        const TE kk = fabs(kappa.v);
        bool late;
        if (kk >= TE(1e-3) && kk <= TE(1e9) && (ae > thr * Guard<TE>::hi() || ae < thr * Guard<TE>::lo())) {
            late = ae > thr;
        } else {
            E cq[1] = {o.cost}, cdk[1];
            div_group<1>(cq, kappa, cdk);
            late = (sqrt_rn(E(TE(-2)) * cdk[0]) * E(Consts<TE>::rad2deg())).v > TE(0.5);
        }
        if (late) s.conv_step = k;
    }
    if (is_nan(xn[0]) || is_nan(xn[1])) s.flags |= RL4_SPF_X_NAN;
    if (is_nan(o.cost)) s.diverged_step = k;                            // objects.py:991 (Q19)
}

}  // namespace rl4
