// Device core of the nonlinear IDHP path: Ce500NonLinear wrapper + plant + IDHPnonlin agent, one aircraft+agent per thread
// (nl_run_kernel).  Included by nl_kernels.cu (PLANT = 0: the surrogate plant of include/rl4_citation_surrogate.h) and by
// dasmat_plant.cu (PLANT = 1: the reference's own aircraft model, translated from its binary; that unit defines
// RL4_NL_WITH_DASMAT and the DasmatThread type before including this file).
#pragma once
#define RL4_SLOWPATH_OUT_OF_LINE 1
#include "rl4_math.cuh"
#include "rl4_runtime.h"
#include "../../include/rl4afcs_b200.h"
#include <cstring>

namespace rl4 {

template <bool PER_AGENT>
struct NlHp {
    const rl4_nl_params& p;
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

template <typename TN> __device__ __forceinline__ TN nfma(TN a, TN b, TN c);
template <> __device__ __forceinline__ float nfma<float>(float a, float b, float c) { return __fmaf_rn(a, b, c); }
template <> __device__ __forceinline__ double nfma<double>(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float nsqrt(float a) { return sqrt_rn(Rn<float>(a)).v; }
__device__ __forceinline__ double nsqrt(double a) { return sqrt_rn(Rn<double>(a)).v; }
// sqrt(d * d) as the reference writes |d| (objects.py:1379-1380, envs/nonlinear/env.py:251).  In binary floating point with
// round-to-nearest sqrt(fl(d*d)) == |d| whenever d*d neither underflows nor overflows (checked exhaustively for every
// float32 in [2^-60, 2^60) and by sampling for float64); outside that range the literal expression is evaluated.
__device__ __forceinline__ float sqrt_of_square(float d)
{
    const float a = fabsf(d);
    return (a < 1.152921504606846976e18f && (a >= 8.673617379884035e-19f || a == 0.0f)) ? a : nsqrt(d * d);
}
__device__ __forceinline__ double sqrt_of_square(double d)
{
    const double a = fabs(d);
    return (a < 3.273390607896142e150 && (a >= 3.054936363499605e-151 || a == 0.0)) ? a : nsqrt(d * d);   // 2^+-500
}

// CTA size: 256 threads (one CTA per SM: 255 registers x 256 threads fill the register file, 174 KB of shared memory) for
// float32 networks, so that ONE barrier per step re-aligns all eight resident warps; 224 for float64 networks (their
// shared-memory share, 960 B per thread, allows seven warps per SM: 0.68e9 -> 0.94e9 agent-steps/s against 128 threads)
#ifndef RL4_NL_BLOCK_F32
#define RL4_NL_BLOCK_F32 256
#endif
#ifndef RL4_NL_BLOCK_F64
#define RL4_NL_BLOCK_F64 224
#endif
template <typename TN> struct NlBlock { static constexpr int v = RL4_NL_BLOCK_F64; };
template <> struct NlBlock<float> { static constexpr int v = RL4_NL_BLOCK_F32; };
// with the translated plant (PLANT = 1) the step is ~430 000 instructions of code far larger than the instruction cache and
// it waits on thread-local memory: one CTA per SM whose warps walk the code together, as many warps as the register file
// allows (measured on the plant alone, final translation: 512 x 1 -> 2.02e7, 768 x 1 -> 2.13e7, 1024 x 1 -> 2.33e7 plant steps/s)
#ifndef RL4_NL_BLOCK_DASMAT
#define RL4_NL_BLOCK_DASMAT 1024
#endif
constexpr int kNlBlockDasmat = RL4_NL_BLOCK_DASMAT;
#ifndef RL4_NL_MINB
#define RL4_NL_MINB 1
#endif
#ifndef RL4_NL_STEP_BARRIER
#define RL4_NL_STEP_BARRIER 1
#endif
#ifndef RL4_NL_SMEM_RLS
#define RL4_NL_SMEM_RLS 0       // RLS parameters (12) and covariance (16) in shared memory instead of registers
#endif
#ifndef RL4_NL_SMEM_ACTOR
#define RL4_NL_SMEM_ACTOR 0     // actor weights (50) in shared memory instead of registers
#endif
constexpr int kNlSmemDoubles = 50 + (RL4_NL_SMEM_RLS ? 28 : 0);
constexpr int kNlSmemNet = 70 + (RL4_NL_SMEM_ACTOR ? 50 : 0);

// per-thread arrays kept in shared memory, laid out [element][thread] (conflict-free, no indexing cost):
// the actor trace E (50 doubles) and the target-critic weights (70 values) are touched once or twice per
// step, so they are the cheapest state to move out of the 255-register budget
template <typename T, int BLOCK> struct Strided {
    T* p;
    __device__ __forceinline__ T& operator[](int j) const { return p[j * BLOCK]; }
};

// element j of a per-agent array stored in a global SoA plane (step-API kernels)
template <typename T> struct PlaneCol {
    T* p;
    int64_t stride;
    __device__ __forceinline__ T& operator[](int j) const { return p[(int64_t)j * stride]; }
};

// picks the local-array form (PlaneCol with stride 1) or the shared-memory form of a per-thread array
template <bool LOCAL, typename L, typename S> struct DasmatPick;
template <typename L, typename S> struct DasmatPick<true, L, S> {
    using type = L;
    template <typename T> static __device__ __forceinline__ L make(T* local, T*) { return L{local, 1}; }
};
template <typename L, typename S> struct DasmatPick<false, L, S> {
    using type = S;
    template <typename T> static __device__ __forceinline__ S make(T*, T* shared) { return S{shared}; }
};

// hidden layer of a 4-10-k net (Network.base_call, objects.py:111-139)
template <typename TN, typename WA>
__device__ __forceinline__ void nl_hidden(const TN (&s)[4], const WA W1, TN (&h)[10])
{
    Rn<TN> pre[10], out[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) {
        TN acc = s[0] * W1[j];
#pragma unroll
        for (int i = 1; i < 4; ++i) acc = nfma<TN>(s[i], W1[i * 10 + j], acc);
        pre[j] = Rn<TN>(acc);
    }
    tanh_t13_n<10>(pre, out);
#pragma unroll
    for (int j = 0; j < 10; ++j) h[j] = out[j].v;
}

// Actor_big.call (objects.py:374-407): forward + trace update
template <typename TN, typename WA, typename EA>
__device__ __forceinline__ TN nl_actor(const TN (&s)[4], const WA W1, const WA W2, const EA Ea,
                                       int elig, double gl, TN (&h)[10], TN& ai1)
{
    nl_hidden<TN>(s, W1, h);
    TN o = h[0] * W2[0];
#pragma unroll
    for (int j = 1; j < 10; ++j) o = nfma<TN>(h[j], W2[j], o);
    const TN a = tanh_t13(Rn<TN>(o)).v;
    ai1 = TN(1) - a * a;
    if (elig == RL4_ELIG_REPLACING) {
        double ng = 0.0, ne = 0.0;
        for (int j = 0; j < 10; ++j) { const double g = (double)(ai1 * h[j]); ng = __fma_rn(g, g, ng); ne = __fma_rn(Ea[j], Ea[j], ne); }
        for (int j = 0; j < 10; ++j) {
            const TN v = (ai1 * W2[j]) * (TN(1) - h[j] * h[j]);
            for (int i = 0; i < 4; ++i) { const double g = (double)(v * s[i]); ng = __fma_rn(g, g, ng); ne = __fma_rn(Ea[10 + j * 4 + i], Ea[10 + j * 4 + i], ne); }
        }
        const bool take = sqrt_rn(Rn<double>(ng)).v > sqrt_rn(Rn<double>(ne)).v;
        for (int j = 0; j < 10; ++j) {
            Ea[j] = take ? (double)(ai1 * h[j]) : Ea[j] * gl;
            const TN v = (ai1 * W2[j]) * (TN(1) - h[j] * h[j]);
            for (int i = 0; i < 4; ++i) Ea[10 + j * 4 + i] = take ? (double)(v * s[i]) : Ea[10 + j * 4 + i] * gl;
        }
    } else {
        const bool acc = (elig == RL4_ELIG_ACCUMULATING);
#pragma unroll
        for (int j = 0; j < 10; ++j) {
            const double g = (double)(ai1 * h[j]);                            // objects.py:385
            Ea[j] = acc ? (Ea[j] * gl + g) : g;
            const TN v = (ai1 * W2[j]) * (TN(1) - h[j] * h[j]);               // objects.py:386
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double gi = (double)(v * s[i]);
                Ea[10 + j * 4 + i] = acc ? (Ea[10 + j * 4 + i] * gl + gi) : gi;
            }
        }
    }
    return a;
}

// RLS.update for n = 3, m = 1 (objects.py:492-543); numpy `@` orders as measured for these shapes:
// params.T @ X -> fma(a0,b0,a1*b1) + fma(a2,b2,a3*b3); Cov @ X -> (p0+p2)+(p1+p3) with rounded products
template <typename TH, typename CV>
__device__ __forceinline__ void nl_rls_update(TH& th, CV& cv, const double (&Xr)[4], const double (&Y)[3], double rgam,
                                              double (&eps)[3], double& eps_norm)
{
#pragma unroll
    for (int ii = 0; ii < 3; ++ii) {
        const double pred = __dadd_rn(__fma_rn(th[ii], Xr[0], __dmul_rn(th[3 + ii], Xr[1])),
                                      __fma_rn(th[6 + ii], Xr[2], __dmul_rn(th[9 + ii], Xr[3])));
        eps[ii] = Y[ii] - pred;
    }
    Rn<double> CX[4], K[4];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
        const double p0 = cv[ii * 4] * Xr[0], p1 = cv[ii * 4 + 1] * Xr[1], p2 = cv[ii * 4 + 2] * Xr[2], p3 = cv[ii * 4 + 3] * Xr[3];
        CX[ii] = Rn<double>((p0 + p2) + (p1 + p3));
    }
    double xcx = Xr[0] * CX[0].v;
#pragma unroll
    for (int ii = 1; ii < 4; ++ii) xcx = __fma_rn(Xr[ii], CX[ii].v, xcx);
    div_group<4>(CX, Rn<double>(rgam + xcx), K);
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int ii = 0; ii < 3; ++ii) th[t * 3 + ii] = th[t * 3 + ii] + K[t].v * eps[ii];
    if (rgam == 1.0) {
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
#pragma unroll
            for (int j = 0; j < 4; ++j) cv[ii * 4 + j] = cv[ii * 4 + j] - K[ii].v * CX[j].v;
    } else {
        Rn<double> num[16], out[16];
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
#pragma unroll
            for (int j = 0; j < 4; ++j) num[ii * 4 + j] = Rn<double>(cv[ii * 4 + j] - K[ii].v * CX[j].v);
        div_group<16>(num, Rn<double>(rgam), out);
#pragma unroll
        for (int j = 0; j < 16; ++j) cv[j] = out[j].v;
    }
    eps_norm = nsqrt(__fma_rn(eps[2], eps[2], __fma_rn(eps[1], eps[1], eps[0] * eps[0])));
}

// Ce500NonLinear.step without the agent (envs/nonlinear/env.py:182-256)
#ifndef RL4_NL_WITH_DASMAT
struct DasmatThread;                        // only dasmat_plant.cu has the translated model
#endif
struct DasmatIo { uint8_t* image; uint64_t* state; int64_t stride; int32_t* err; };

template <bool PER_AGENT, int INTEG, int PLANT = 0>
__device__ __forceinline__ void nl_env_step(const rl4_nl_params& p, const NlHp<PER_AGENT>& hv, int stepp, double theta_ref_k,
                                            const double (&act)[3], double (&x)[12], double (&x_act)[3], double (&surf)[3],
                                            double& e_phi, double& e_th, double& e_psi, double& reward, double& rg2, double (&ueff)[3],
                                            double (&xo)[12], DasmatThread* dz = nullptr)
{
    const int fault_step = hv.hpi(RL4_NHPI_FAULT_STEP);
    const bool faulted = (fault_step >= 0 && stepp >= fault_step);                 // env.py:132,151
    const int damp = hv.hpi(RL4_NHPI_FAULT_DAMP), sat = hv.hpi(RL4_NHPI_FAULT_SAT);
    const double omega = (faulted && damp == RL4_NL_SLOW_ALL && stepp > fault_step) ? p.omega_slow : p.omega0;
    double eff[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) eff[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double hi = p.limit_deg[i], lo = -p.limit_deg[i];
        double v = act[i] * (hi - lo) / 2.0;                                       // _scale_action env.py:111-124
        v = v + (hi + lo) / 2.0;
        const double cmd = v * (3.14159265358979323846 / 180.0);
        double d = cmd - x_act[i];                                                 // _propagate_surfaces_states env.py:161-180
        d = d * omega;
        d = d < -p.rate_limit ? -p.rate_limit : (d > p.rate_limit ? p.rate_limit : d);
        x_act[i] = x_act[i] + p.dt * d;
        surf[i] = x_act[i];
    }
    if (faulted && sat != RL4_NL_SAT_NONE) {                                       // _saturate_surfaces env.py:150-159
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if (sat - 1 == j) { const double L = p.sat_limit[j]; surf[j] = surf[j] < -L ? -L : (surf[j] > L ? L : surf[j]); }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) eff[i] = surf[i];
    if (faulted) {                                                                 // _engage_fault env.py:129-148
        const double f = hv.hp(RL4_NHP_DAMP_FACTOR);
        if (damp == RL4_NL_DAMP_ELEVATOR || damp == RL4_NL_DAMP_ALL) eff[0] *= f;
        if (damp == RL4_NL_DAMP_AILERON || damp == RL4_NL_DAMP_ALL) eff[1] *= f;
        if (damp == RL4_NL_DAMP_RUDDER || damp == RL4_NL_DAMP_ALL) eff[2] *= f;
        if (damp == RL4_NL_SHIFT_CG) eff[10] = hv.hp(RL4_NHP_CG_SHIFT);
    }
    double u[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) u[i] = p.trim_input[i] + eff[i];                  // env.py:207-208
    ueff[0] = u[0]; ueff[1] = u[1]; ueff[2] = u[2];
    // env.py:210  x_full = model.step(input).  The reference's plant is an output-then-update block: step() returns the state
    // BEFORE the step and then integrates (found by running the binary in-process: DESIGN.md section 9), so the wrapper observes the
    // aircraft one sample late: xo is what model.step returned, x the carried state.
    // Symmetric flight (elevator-only commands from a trimmed start: always, in IDHPnonlin's task) takes the
    // longitudinal form of the same equations -- identical values, ~40 % less work and half the stage storage; INTEG is
    // a compile-time constant, so one integrator's code per kernel
#ifdef RL4_NL_WITH_DASMAT
    if (PLANT == 1) {
        // the reference's own model: step(u) returns the state before the step and integrates (the binary's behaviour itself)
        dasmat_thread_step(dz, u, xo);
#pragma unroll
        for (int j = 0; j < 12; ++j) x[j] = xo[j];
    } else
#endif
    {
#pragma unroll
        for (int j = 0; j < 12; ++j) xo[j] = x[j];
        rl4_cit_step_auto(&p.plant, x, u, p.dt, INTEG);
    }
    const double Q = hv.hp(RL4_NHP_Q_SYM);
    e_phi = xo[6] - 0.0; e_th = xo[7] - theta_ref_k; e_psi = xo[8] - 0.0;         // env.py:215 (state - ref)
    reward = (-0.5 * Q) * (e_th * e_th);                                           // env.py:218
    rg2 = (-Q) * e_th;                                                             // env.py:219-220 (q slot)
}

// (one out-of-line copy each: three call sites per step, four IEEE divisions per call)
static __device__ __noinline__ double nl_decay(double a, double b, double c, bool f32)
{   // objects.py:1235-1243, numpy-1.x promotion: float64 intermediates, network-dtype result
    double r = b / a;
    r = c + (1.0 - c) * r;
    const double a2 = a * (0.998 + (1.0 - 0.998) * b / a);
    const double v = a2 * r;
    return f32 ? (double)(float)v : v;
}
// the same under NEP 50 (numpy >= 2): once `a` is a 0-d float32 array the python-float operands are weak -> float32
// arithmetic with the constants rounded to float32 first; while `a` is still a python float it is float64 arithmetic
static __device__ __noinline__ double nl_decay_np2(double a, double b, double c, bool a_is_pyfloat)
{
    if (a_is_pyfloat) {
        double r = b / a;
        r = c + (1.0 - c) * r;
        const double a2 = a * (0.998 + (1.0 - 0.998) * b / a);
        return (double)(float)(a2 * r);
    }
    const float af = (float)a;
    float r = __fdiv_rn((float)b, af);
    r = __fadd_rn((float)c, __fmul_rn((float)(1.0 - c), r));
    const float q = __fdiv_rn((float)((1.0 - 0.998) * b), af);
    const float a2 = __fmul_rn(af, __fadd_rn((float)0.998, q));
    return (double)__fmul_rn(a2, r);
}
__device__ __forceinline__ bool nl_isclose(double a, double b) { return fabs(a - b) <= (1e-8 + 1e-5 * fabs(b)); }

template <typename TN, int INTEG, bool LOG, bool PER_AGENT, int PLANT = 0>
__global__ void __launch_bounds__((PLANT == 1 ? kNlBlockDasmat : NlBlock<TN>::v), RL4_NL_MINB)
nl_run_kernel(const __grid_constant__ rl4_nl_params p, const double* __restrict__ theta_ref, const float* __restrict__ noise,
              int64_t noise_stride, int k0, int n_steps, const rl4_nl_state st, int64_t n_agents, const rl4_sp_log lg,
              const DasmatIo dio = DasmatIo{})
{
    // PER_AGENT = false (no per-agent override array at all, the common case): every hyper-parameter read is a constant-
    // bank operand; true: a pointer test + load per read (about 150 instructions and the reloads of the spilled agent
    // index per step -- 10 % of the step time, so the uniform case gets its own instantiation)
    // Tail threads of the last CTA do not exit: every thread of the CTA must reach the per-step barrier.  They shadow the
    // last agent (loads only), are frozen like a diverged agent, and skip the store.
    const int64_t i_raw = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i_raw < n_agents;
    const int64_t i = active ? i_raw : n_agents - 1;
    const NlHp<PER_AGENT> hv{p, i};
    const int64_t S = st.stride;
    double* __restrict__ E = st.env + i;
    TN* __restrict__ Nn = (TN*)st.net + i;
    int32_t* __restrict__ I = st.ints + i;
#define EF(f) E[(int64_t)(f) * S]
#define NF(f) Nn[(int64_t)(f) * S]

    // ---- load ----
    extern __shared__ __align__(16) unsigned char nl_smem[];
    // layout: doubles first ([Ea 50][th 12 + cv 16 when RL4_NL_SMEM_RLS]), then TN ([W1t 40][W2t 30][W1a 40 + W2a 10 when RL4_NL_SMEM_ACTOR])
    double* const sm_d = reinterpret_cast<double*>(nl_smem) + threadIdx.x;
    constexpr int BLK = PLANT == 1 ? kNlBlockDasmat : NlBlock<TN>::v;
    TN* const sm_n = reinterpret_cast<TN*>(nl_smem + sizeof(double) * kNlSmemDoubles * BLK) + threadIdx.x;
#ifdef RL4_NL_WITH_DASMAT
    // PLANT = 1: 512-thread CTAs would need 348 KB of shared memory for these; the plant is >99 % of the step there, so the
    // trace and the target critic simply live in thread-local memory (same element order, same arithmetic)
    double ea_loc[PLANT == 1 ? 50 : 1];
    TN w1t_loc[PLANT == 1 ? 40 : 1], w2t_loc[PLANT == 1 ? 30 : 1];
    using EaT = typename DasmatPick<PLANT == 1, PlaneCol<double>, Strided<double, BLK>>::type;
    using WtT = typename DasmatPick<PLANT == 1, PlaneCol<TN>, Strided<TN, BLK>>::type;
    const EaT Ea = DasmatPick<PLANT == 1, PlaneCol<double>, Strided<double, BLK>>::make(ea_loc, sm_d);
    const WtT W1t = DasmatPick<PLANT == 1, PlaneCol<TN>, Strided<TN, BLK>>::make(w1t_loc, sm_n);
    const WtT W2t = DasmatPick<PLANT == 1, PlaneCol<TN>, Strided<TN, BLK>>::make(w2t_loc, sm_n + 40 * BLK);
#else
    const Strided<double, BLK> Ea{sm_d};
    const Strided<TN, BLK> W1t{sm_n};
    const Strided<TN, BLK> W2t{sm_n + 40 * BLK};
#endif
#if RL4_NL_SMEM_RLS
    const Strided<double, BLK> th{sm_d + 50 * BLK};
    const Strided<double, BLK> cv{sm_d + 62 * BLK};
    double x[12], x_act[3], x_lon[3], x_prev_lon[3], eps[3];
#else
    double x[12], x_act[3], x_lon[3], x_prev_lon[3], th[12], cv[16], eps[3];
#endif
#if RL4_NL_SMEM_ACTOR
    const Strided<TN, BLK> W1a{sm_n + 70 * BLK};
    const Strided<TN, BLK> W2a{sm_n + 110 * BLK};
    TN s[4], s_prev[4], W1c[40], W2c[30], Mp[9];
#else
    TN s[4], s_prev[4], W1a[40], W2a[10], W1c[40], W2c[30], Mp[9];
#endif
    for (int j = 0; j < 12; ++j) { x[j] = EF(RL4_NLE_XFULL + j); th[j] = EF(RL4_NLE_THETA + j); }
    for (int j = 0; j < 3; ++j) { x_act[j] = EF(RL4_NLE_XACT + j); x_lon[j] = EF(RL4_NLE_XLON + j); x_prev_lon[j] = EF(RL4_NLE_XPREVLON + j); eps[j] = EF(RL4_NLE_EPS + j); }
    for (int j = 0; j < 16; ++j) cv[j] = EF(RL4_NLE_COV + j);
    for (int j = 0; j < 50; ++j) Ea[j] = EF(RL4_NLE_EA + j);
    double cgp2 = EF(RL4_NLE_CGRAD_PREV), eps_norm = EF(RL4_NLE_EPS_NORM), rse0 = EF(RL4_NLE_RSE), rse1 = EF(RL4_NLE_RSE + 1);
    double rse_f0 = EF(RL4_NLE_RSE_FLIGHT), rse_f1 = EF(RL4_NLE_RSE_FLIGHT + 1);
    const int flight_step = hv.hpi(RL4_NHPI_FLIGHT_STEP);
    double nz_peak = EF(RL4_NLE_NZ_PEAK), eta_a = EF(RL4_NLE_ETA_A), eta_c = EF(RL4_NLE_ETA_C), lambdaa = EF(RL4_NLE_LAMBDAA), gl = EF(RL4_NLE_GL);
    for (int j = 0; j < 4; ++j) { s[j] = NF(RL4_NLN_S + j); s_prev[j] = NF(RL4_NLN_SPREV + j); }
    TN a = NF(RL4_NLN_A), a_prev = NF(RL4_NLN_APREV), lr_a = NF(RL4_NLN_LR_A), lr_c = NF(RL4_NLN_LR_C);
    for (int j = 0; j < 40; ++j) { W1a[j] = NF(RL4_NLN_W1A + j); W1c[j] = NF(RL4_NLN_W1C + j); W1t[j] = NF(RL4_NLN_W1T + j); }
    for (int j = 0; j < 10; ++j) W2a[j] = NF(RL4_NLN_W2A + j);
    for (int j = 0; j < 30; ++j) { W2c[j] = NF(RL4_NLN_W2C + j); W2t[j] = NF(RL4_NLN_W2T + j); }
    for (int j = 0; j < 9; ++j) Mp[j] = NF(RL4_NLN_MPREV + j);
    int cooldown = I[(int64_t)RL4_NLI_COOLDOWN * S], diverged_step = I[(int64_t)RL4_NLI_DIVERGED_STEP * S], stepp = I[(int64_t)RL4_NLI_STEPP * S];
    int pyfloat_mask = I[(int64_t)RL4_NLI_PYFLOAT_MASK * S];
    if (!active) diverged_step = 0;

    const bool logged = LOG && active && i < lg.n_agents_logged;
    const bool f32 = sizeof(TN) == 4;
#ifdef RL4_NL_WITH_DASMAT
    DasmatThread dz_storage;
    DasmatThread* const dzp = PLANT == 1 ? &dz_storage : nullptr;
    if (PLANT == 1) dasmat_thread_load(dzp, dio, i);
#else
    DasmatThread* const dzp = nullptr;
#endif
    int k = k0;
    for (; k < k0 + n_steps; ++k) {
#if RL4_NL_STEP_BARRIER
        // Re-align the warps of the CTA once per step: the step is ~8 000 straight-line instructions (~128 KB), as large
        // as the instruction cache; warps that drift apart each stream the whole loop through it, warps that walk it
        // together share the fetched lines.  Frozen (diverged) agents keep arriving at the barrier and skip the body.
        __syncthreads();
        // PLANT = 1: the translated model synchronises the CTA at every step() entry (it is far larger than the instruction cache),
        // so EVERY thread of the CTA must take the plant step, through the one call site below: a frozen (diverged) agent flies
        // it too -- on whatever its model holds, results discarded, its own state restored -- and leaves afterwards.  (Disabling
        // the barrier for CTAs with a diverged agent instead cost 40 % of the throughput of a 90 s episode: with 1 % of the
        // agents diverged nearly every 1024-agent CTA holds one.)
        const bool frozen = diverged_step >= 0;
        if (PLANT != 1 && frozen) {                                                // objects.py:1557 (break) + :1168-1175 (NaN rows)
            if (LOG) {
                if (logged && (k - k0) % lg.every == 0) {
                    const int nf = lg.level >= 3 ? RL4_NLM_COUNT : (lg.level == 2 ? RL4_NLF_COUNT : RL4_NLL_COUNT);
                    double* b = lg.buf + ((int64_t)((k - k0) / lg.every) * nf) * lg.n_agents_logged + i;
                    for (int f = 0; f < nf; ++f) b[(int64_t)f * lg.n_agents_logged] = __longlong_as_double(0x7ff8000000000000LL);
                }
            }
            continue;
        }
#else
        if (diverged_step >= 0) break;                                             // objects.py:1557
#endif
        const TN a_k = a;
        // ---- env.step(self._get_action(a))  (objects.py:1497, 1448-1455)
        const double act[3] = {(double)a_k, 0.0, 0.0};
        double surf[3], ueff[3], e_phi, e_th, e_psi, reward, rg2;
        const double yref_k = __ldg(theta_ref + k);
        double xo[12];                                                             // what model.step returned (the state before this step)
#ifdef RL4_NL_WITH_DASMAT
        double x_keep[PLANT == 1 ? 12 : 1], xa_keep[PLANT == 1 ? 3 : 1];
        if (PLANT == 1 && frozen) {
#pragma unroll
            for (int j = 0; j < 12; ++j) x_keep[j] = x[j];
#pragma unroll
            for (int j = 0; j < 3; ++j) xa_keep[j] = x_act[j];
        }
#endif
        nl_env_step<PER_AGENT, INTEG, PLANT>(p, hv, stepp, yref_k, act, x, x_act, surf, e_phi, e_th, e_psi, reward, rg2, ueff, xo, dzp);
#ifdef RL4_NL_WITH_DASMAT
        if (PLANT == 1 && frozen) {
#pragma unroll
            for (int j = 0; j < 12; ++j) x[j] = x_keep[j];
#pragma unroll
            for (int j = 0; j < 3; ++j) x_act[j] = xa_keep[j];
            if (LOG) {
                if (logged && (k - k0) % lg.every == 0) {
                    const int nf = lg.level >= 3 ? RL4_NLM_COUNT : (lg.level == 2 ? RL4_NLF_COUNT : RL4_NLL_COUNT);
                    double* b = lg.buf + ((int64_t)((k - k0) / lg.every) * nf) * lg.n_agents_logged + i;
                    for (int f = 0; f < nf; ++f) b[(int64_t)f * lg.n_agents_logged] = __longlong_as_double(0x7ff8000000000000LL);
                }
            }
            continue;
        }
#endif
        stepp += 1;
        const double x_next_lon[3] = {xo[4], xo[7], xo[1]};                           // env.py:231
        bool nans = false;
#pragma unroll
        for (int j = 0; j < 12; ++j) nans |= (xo[j] != xo[j]);
        const double rse_k0 = sqrt_of_square(e_th), rse_k1 = nsqrt(e_phi * e_phi + e_psi * e_psi);   // env.py:251
        rse0 += rse_k0;                                                            // objects.py:1503-1504
        rse1 += rse_k1;
        if (k >= flight_step) { rse_f0 += rse_k0; rse_f1 += rse_k1; }              // functions.py:917,1039
        // full-log row of this step (level 2): the two loss gradients are written where they are formed
        double* const fb = (LOG && logged && lg.level == 2 && (k - k0) % lg.every == 0)
                               ? lg.buf + ((int64_t)((k - k0) / lg.every) * RL4_NLF_COUNT) * lg.n_agents_logged + i : nullptr;
        { const double nz = fabs(xo[3] * xo[1] / 9.80665); if (nz > nz_peak) nz_peak = nz; }   // functions.py:774,1055
        TN s_next[4] = {(TN)xo[4], (TN)xo[7], (TN)xo[1], (TN)e_th};                   // env.py:236-238; objects.py:1499

        // ---- _step_networks (objects.py:1292-1348)
        TN hc[10], ht[10], lam[3], lt[3];
        nl_hidden<TN>(s_prev, (const TN*)W1c, hc);                                 // critic(s_prev)
        nl_hidden<TN>(s_next, W1t, ht);                                            // target_critic(s_next)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            TN acc = hc[0] * W2c[q], acc2 = ht[0] * W2t[q];
#pragma unroll
            for (int j = 1; j < 10; ++j) { acc = nfma<TN>(hc[j], W2c[j * 3 + q], acc); acc2 = nfma<TN>(ht[j], W2t[j * 3 + q], acc2); }
            lam[q] = acc; lt[q] = acc2;
        }
        // actor(s_prev) with trace pass 1 (objects.py:1310) and, for k > 0, actor(s_random) with trace pass 2
        // (objects.py:1375-1378).  The second pass only needs the actor weights and the trace, which nothing touches in
        // between, so both passes run back to back through ONE copy of the code (a two-trip loop, not unrolled).
        const int elig_a = hv.hpi(RL4_NHPI_ELIG_A);
        TN ha[10], ai1 = TN(0), a_next = TN(0), a_random = TN(0);
        {
            const TN nz = (TN)__ldg(noise + (int64_t)(k - k0) * noise_stride + i);
            const int n_pass = (k > 0) ? 2 : 1;
#pragma unroll 1
            for (int pass = 0; pass < n_pass; ++pass) {
                TN sin[4], hh[10], aa1;
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) sin[ii] = pass ? (nz * (TN)p.noise_std[ii] + s_prev[ii]) : s_prev[ii];
                const TN aout = nl_actor<TN>(sin, W1a, W2a, Ea, elig_a, gl, hh, aa1);
                if (pass == 0) {
                    a_next = aout; ai1 = aa1;
#pragma unroll
                    for (int j = 0; j < 10; ++j) ha[j] = hh[j];
                } else {
                    a_random = aout;
                }
            }
        }
        TN dads[4];                                                                // tape.gradient(a, s_prev) (objects.py:1323)
        {
            const TN g_o = TN(1) * ai1;
            TN gp[10];
#pragma unroll
            for (int j = 0; j < 10; ++j) gp[j] = (g_o * W2a[j]) * (TN(1) - ha[j] * ha[j]);
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                TN acc = gp[0] * W1a[ii * 10];
#pragma unroll
                for (int j = 1; j < 10; ++j) acc = nfma<TN>(gp[j], W1a[ii * 10 + j], acc);
                dads[ii] = acc;
            }
        }
        TN M[9], Gn[3];                                                            // objects.py:1327-1329
#pragma unroll
        for (int ii = 0; ii < 3; ++ii) Gn[ii] = (TN)th[9 + ii];
#pragma unroll
        for (int ii = 0; ii < 3; ++ii)
#pragma unroll
            for (int j = 0; j < 3; ++j) M[ii * 3 + j] = (TN)th[j * 3 + ii] + Gn[ii] * dads[j];

        if (k > 0) {
            // ---- _update_networks (objects.py:1350-1399)
            const double gamma_d = hv.hp(RL4_NHP_GAMMA);
            const TN gam = (TN)gamma_d;
            const double rg[3] = {0.0, 0.0, rg2};
            TN td[3];
            if (hv.hpi(RL4_NHPI_MULTISTEP)) {                                      // objects.py:1360 (Q14)
                const double gc[3] = {gamma_d * rg[0], gamma_d * rg[1], gamma_d * rg[2]};
                const TN g2 = (TN)hv.hp(RL4_NHP_GAMMA_SQ);
                const TN gl3[3] = {g2 * lt[0], g2 * lt[1], g2 * lt[2]};
                TN V[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) { TN acc = gl3[0] * M[j]; acc = nfma<TN>(gl3[1], M[3 + j], acc); acc = nfma<TN>(gl3[2], M[6 + j], acc); V[j] = acc; }
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    double t1 = gc[0] * (double)Mp[j];                             // numpy f64 (1,3)@(3,3)
                    t1 = __fma_rn(gc[1], (double)Mp[3 + j], t1);
                    t1 = __fma_rn(gc[2], (double)Mp[6 + j], t1);
                    TN t2 = V[0] * Mp[j]; t2 = nfma<TN>(V[1], Mp[3 + j], t2); t2 = nfma<TN>(V[2], Mp[6 + j], t2);
                    const TN cgp = (j == 2) ? (TN)cgp2 : TN(0);
                    td[j] = ((lam[j] - cgp) - (TN)t1) - t2;
                }
            } else {                                                               // objects.py:1362
                const TN gl3[3] = {gam * lt[0], gam * lt[1], gam * lt[2]};
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    TN acc = gl3[0] * M[j]; acc = nfma<TN>(gl3[1], M[3 + j], acc); acc = nfma<TN>(gl3[2], M[6 + j], acc);
                    td[j] = (lam[j] - (TN)rg[j]) - acc;
                }
            }
            {   // critic VJP (tape.gradient with output_gradients = td, objects.py:1365) + SGD (:1368)
                TN dpre[10];
#pragma unroll
                for (int j = 0; j < 10; ++j) {
                    TN dh = td[0] * W2c[j * 3];
                    dh = nfma<TN>(td[1], W2c[j * 3 + 1], dh);
                    dh = nfma<TN>(td[2], W2c[j * 3 + 2], dh);
                    dpre[j] = dh * (TN(1) - hc[j] * hc[j]);
                }
#pragma unroll
                for (int j = 0; j < 10; ++j)
#pragma unroll
                    for (int q = 0; q < 3; ++q) W2c[j * 3 + q] = W2c[j * 3 + q] - lr_c * (hc[j] * td[q]);
#pragma unroll
                for (int ii = 0; ii < 4; ++ii)
#pragma unroll
                    for (int j = 0; j < 10; ++j) W1c[ii * 10 + j] = W1c[ii * 10 + j] - lr_c * (s_prev[ii] * dpre[j]);
                if (LOG) {
                    if (fb) {                                                      // critic_loss_grad (objects.py:1365,1151-1152)
                        const int64_t L = lg.n_agents_logged;
                        for (int ii = 0; ii < 4; ++ii)
                            for (int j = 0; j < 10; ++j) fb[(int64_t)(RL4_NLF_C_GRAD + ii * 10 + j) * L] = (double)(s_prev[ii] * dpre[j]);
                        for (int j = 0; j < 10; ++j)
                            for (int q = 0; q < 3; ++q) fb[(int64_t)(RL4_NLF_C_GRAD + 40 + j * 3 + q) * L] = (double)(hc[j] * td[q]);
                    }
                }
            }
            {   // target soft update (objects.py:1371)
                const double tau_d = hv.hp(RL4_NHP_TAU);
                const TN omt = (TN)(1.0 - tau_d), tt = (TN)tau_d;
#pragma unroll
                for (int j = 0; j < 40; ++j) W1t[j] = omt * W1t[j] + tt * W1c[j];
#pragma unroll
                for (int j = 0; j < 30; ++j) W2t[j] = omt * W2t[j] + tt * W2c[j];
            }
            {   // actor (objects.py:1379-1388): smoothness terms, loss, update (the s_random pass ran above)
                const TN dT = a_k - a_next, dS = a_k - a_random;
                const TN L_T = sqrt_of_square(dT), L_S = sqrt_of_square(dS);
                TN v[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) v[j] = -((TN)rg[j] + gam * lt[j]);
                TN acc = v[0] * Gn[0]; acc = nfma<TN>(v[1], Gn[1], acc); acc = nfma<TN>(v[2], Gn[2], acc);
                const TN loss = (acc + (TN)hv.hp(RL4_NHP_LAMBDA_T) * L_T) + (TN)hv.hp(RL4_NHP_LAMBDA_S) * L_S;
#pragma unroll
                for (int j = 0; j < 10; ++j) W2a[j] = W2a[j] - lr_a * (loss * (TN)Ea[j]);      // objects.py:417-427
#pragma unroll
                for (int ii = 0; ii < 4; ++ii)
#pragma unroll
                    for (int j = 0; j < 10; ++j) W1a[ii * 10 + j] = W1a[ii * 10 + j] - lr_a * (loss * (TN)Ea[10 + j * 4 + ii]);
                if (LOG) {
                    if (fb) {                                                      // actor_loss_grad (objects.py:1387,1154-1155)
                        const int64_t L = lg.n_agents_logged;
                        for (int ii = 0; ii < 4; ++ii)
                            for (int j = 0; j < 10; ++j) fb[(int64_t)(RL4_NLF_A_GRAD + ii * 10 + j) * L] = (double)(loss * (TN)Ea[10 + j * 4 + ii]);
                        for (int j = 0; j < 10; ++j) fb[(int64_t)(RL4_NLF_A_GRAD + 40 + j) * L] = (double)(loss * (TN)Ea[j]);
                    }
                }
            }
            {   // RLS, n = 3, m = 1 (objects.py:1521-1524, 492-543)
                double Xr[4], Y[3];
#pragma unroll
                for (int ii = 0; ii < 3; ++ii) { Xr[ii] = x_lon[ii] - x_prev_lon[ii]; Y[ii] = x_next_lon[ii] - x_lon[ii]; }
                Xr[3] = (double)(a_k - a_prev);
                nl_rls_update(th, cv, Xr, Y, hv.hp(RL4_NHP_RLS_GAMMA), eps, eps_norm);
            }
            {   // _adapt_check (objects.py:1212-1290)
                const bool cond1 = k < hv.hpi(RL4_NHPI_WARMUP_STEPS);
                if (cooldown > 0) cooldown -= 1;
                const bool np2 = f32 && hv.hpi(RL4_NHPI_NUMPY2) != 0;
                if (!cond1) {
                    const double dec = hv.hp(RL4_NHP_LR_DECAY);
                    const double eal = hv.hp(RL4_NHP_ETA_A_L), ecl = hv.hp(RL4_NHP_ETA_C_L), ll = hv.hp(RL4_NHP_LAMBDA_L);
                    if (!np2) {
                        eta_a = nl_isclose(eta_a, eal) ? eal : nl_decay(eta_a, eal, dec, f32);
                        eta_c = nl_isclose(eta_c, ecl) ? ecl : nl_decay(eta_c, ecl, dec, f32);
                        lambdaa = nl_isclose(lambdaa, ll) ? ll : nl_decay(lambdaa, ll, dec, f32);
                    } else {
                        if (nl_isclose(eta_a, eal)) { eta_a = eal; pyfloat_mask |= 1; } else { eta_a = nl_decay_np2(eta_a, eal, dec, pyfloat_mask & 1); pyfloat_mask &= ~1; }
                        if (nl_isclose(eta_c, ecl)) { eta_c = ecl; pyfloat_mask |= 2; } else { eta_c = nl_decay_np2(eta_c, ecl, dec, pyfloat_mask & 2); pyfloat_mask &= ~2; }
                        if (nl_isclose(lambdaa, ll)) { lambdaa = ll; pyfloat_mask |= 4; } else { lambdaa = nl_decay_np2(lambdaa, ll, dec, pyfloat_mask & 4); pyfloat_mask &= ~4; }
                    }
                }
                double lambda_gamma = lambdaa * gamma_d;
                bool differ = ((double)lr_a != eta_a && (double)lr_c != eta_c);
                if (np2) {                                                         // weak python floats: float32 product / comparisons
                    if (!(pyfloat_mask & 4)) lambda_gamma = (double)__fmul_rn((float)lambdaa, (float)gamma_d);
                    differ = ((float)lr_a != (float)eta_a && (float)lr_c != (float)eta_c);
                }
                if (differ && cooldown <= 0) {
                    lr_a = (TN)eta_a; lr_c = (TN)eta_c;
                    gl = lambda_gamma;
                    cooldown = hv.hpi(RL4_NHPI_COOLDOWN_STEPS);
                }
            }
        }
        // ---- shift (objects.py:1530-1538)
#pragma unroll
        for (int j = 0; j < 4; ++j) { s_prev[j] = s[j]; s[j] = s_next[j]; }
        a_prev = a_k; a = a_next;
#pragma unroll
        for (int j = 0; j < 3; ++j) { x_prev_lon[j] = x_lon[j]; x_lon[j] = x_next_lon[j]; }
        cgp2 = rg2;
#pragma unroll
        for (int j = 0; j < 9; ++j) Mp[j] = M[j];
        if (nans) diverged_step = k;
        if (LOG) {
            if (logged && (k - k0) % lg.every == 0) {
                const int64_t L = lg.n_agents_logged;
                const int nf = lg.level >= 3 ? RL4_NLM_COUNT : (lg.level == 2 ? RL4_NLF_COUNT : RL4_NLL_COUNT);
                double* b = lg.buf + ((int64_t)((k - k0) / lg.every) * nf) * L + i;
                if (nans) {                                                        // objects.py:1168-1175: this row and all later ones
                    for (int f = 0; f < nf; ++f) b[(int64_t)f * L] = __longlong_as_double(0x7ff8000000000000LL);
                } else if (lg.level >= 3) {                                        // functions.py:1040-1052
                    b[(int64_t)RL4_NLM_E * L] = e_th; b[(int64_t)RL4_NLM_THETA * L] = xo[7]; b[(int64_t)RL4_NLM_ALPHA * L] = xo[4];
                    b[(int64_t)RL4_NLM_Q * L] = xo[1]; b[(int64_t)RL4_NLM_V * L] = xo[3]; b[(int64_t)RL4_NLM_H * L] = xo[9];
                    b[(int64_t)RL4_NLM_A_CMD * L] = surf[0]; b[(int64_t)RL4_NLM_A_EFF * L] = ueff[0];
                    double na = 0.0, nc = 0.0;
                    for (int j = 0; j < 40; ++j) { na = __fma_rn((double)W1a[j], (double)W1a[j], na); nc = __fma_rn((double)W1c[j], (double)W1c[j], nc); }
                    b[(int64_t)RL4_NLM_WA_NORM * L] = nsqrt(na); b[(int64_t)RL4_NLM_WC_NORM * L] = nsqrt(nc);
                    b[(int64_t)RL4_NLM_RLS_EPS * L] = eps_norm;
                } else if (lg.level == 2) {
                    b[(int64_t)RL4_NLF_ETA_A * L] = eta_a;
                    for (int j = 0; j < 12; ++j) b[(int64_t)(RL4_NLF_XFULL + j) * L] = xo[j];
                    b[(int64_t)RL4_NLF_RSE * L] = rse_k0; b[(int64_t)(RL4_NLF_RSE + 1) * L] = rse_k1;
                    for (int j = 0; j < 3; ++j) b[(int64_t)(RL4_NLF_X + j) * L] = x_next_lon[j];
                    b[(int64_t)RL4_NLF_A_CMD * L] = surf[0]; b[(int64_t)RL4_NLF_A_EFF * L] = ueff[0];
                    b[(int64_t)RL4_NLF_S * L] = xo[4]; b[(int64_t)RL4_NLF_YREF * L] = yref_k; b[(int64_t)RL4_NLF_E * L] = e_th;
                    for (int j = 0; j < 40; ++j) { b[(int64_t)(RL4_NLF_A_W1 + j) * L] = (double)W1a[j]; b[(int64_t)(RL4_NLF_C_W1 + j) * L] = (double)W1c[j]; }
                    for (int j = 0; j < 10; ++j) b[(int64_t)(RL4_NLF_A_W2 + j) * L] = (double)W2a[j];
                    for (int j = 0; j < 30; ++j) b[(int64_t)(RL4_NLF_C_W2 + j) * L] = (double)W2c[j];
                    if (k == 0) for (int j = 0; j < 120; ++j) b[(int64_t)(RL4_NLF_A_GRAD + j) * L] = 0.0;
                    for (int j = 0; j < 12; ++j) b[(int64_t)(RL4_NLF_RLS_PARAMS + j) * L] = th[j];
                    for (int j = 0; j < 16; ++j) b[(int64_t)(RL4_NLF_RLS_COV + j) * L] = cv[j];
                    for (int j = 0; j < 3; ++j) b[(int64_t)(RL4_NLF_RLS_EPS + j) * L] = eps[j];
                    b[(int64_t)RL4_NLF_RLS_EPS_NORM * L] = eps_norm;
                    b[(int64_t)RL4_NLF_A * L] = (double)a_next; b[(int64_t)RL4_NLF_REWARD * L] = reward;
                } else {
                    for (int j = 0; j < 12; ++j) b[(int64_t)(RL4_NLL_XFULL + j) * L] = xo[j];
                    b[(int64_t)RL4_NLL_A * L] = (double)a_next; b[(int64_t)RL4_NLL_E_THETA * L] = e_th; b[(int64_t)RL4_NLL_REWARD * L] = reward;
                    for (int j = 0; j < 3; ++j) b[(int64_t)(RL4_NLL_SURF + j) * L] = surf[j];
                }
            }
        }
    }
    if (LOG) {
        if (logged) {
            const int nf = lg.level >= 3 ? RL4_NLM_COUNT : (lg.level == 2 ? RL4_NLF_COUNT : RL4_NLL_COUNT);
            for (; k < k0 + n_steps; ++k) {
                if ((k - k0) % lg.every) continue;
                double* b = lg.buf + ((int64_t)((k - k0) / lg.every) * nf) * lg.n_agents_logged + i;
                for (int f = 0; f < nf; ++f) b[(int64_t)f * lg.n_agents_logged] = __longlong_as_double(0x7ff8000000000000LL);
            }
        }
    }

    // ---- store ----
    if (!active) return;
#ifdef RL4_NL_WITH_DASMAT
    if (PLANT == 1) dasmat_thread_store(dzp, dio, i);
#endif
    for (int j = 0; j < 12; ++j) { EF(RL4_NLE_XFULL + j) = x[j]; EF(RL4_NLE_THETA + j) = th[j]; }
    for (int j = 0; j < 3; ++j) { EF(RL4_NLE_XACT + j) = x_act[j]; EF(RL4_NLE_XLON + j) = x_lon[j]; EF(RL4_NLE_XPREVLON + j) = x_prev_lon[j]; EF(RL4_NLE_EPS + j) = eps[j]; }
    for (int j = 0; j < 16; ++j) EF(RL4_NLE_COV + j) = cv[j];
    for (int j = 0; j < 50; ++j) EF(RL4_NLE_EA + j) = Ea[j];
    EF(RL4_NLE_CGRAD_PREV) = cgp2; EF(RL4_NLE_EPS_NORM) = eps_norm; EF(RL4_NLE_RSE) = rse0; EF(RL4_NLE_RSE + 1) = rse1;
    EF(RL4_NLE_RSE_FLIGHT) = rse_f0; EF(RL4_NLE_RSE_FLIGHT + 1) = rse_f1;
    EF(RL4_NLE_NZ_PEAK) = nz_peak; EF(RL4_NLE_ETA_A) = eta_a; EF(RL4_NLE_ETA_C) = eta_c; EF(RL4_NLE_LAMBDAA) = lambdaa; EF(RL4_NLE_GL) = gl;
    for (int j = 0; j < 4; ++j) { NF(RL4_NLN_S + j) = s[j]; NF(RL4_NLN_SPREV + j) = s_prev[j]; }
    NF(RL4_NLN_A) = a; NF(RL4_NLN_APREV) = a_prev; NF(RL4_NLN_LR_A) = lr_a; NF(RL4_NLN_LR_C) = lr_c;
    for (int j = 0; j < 40; ++j) { NF(RL4_NLN_W1A + j) = W1a[j]; NF(RL4_NLN_W1C + j) = W1c[j]; NF(RL4_NLN_W1T + j) = W1t[j]; }
    for (int j = 0; j < 10; ++j) NF(RL4_NLN_W2A + j) = W2a[j];
    for (int j = 0; j < 30; ++j) { NF(RL4_NLN_W2C + j) = W2c[j]; NF(RL4_NLN_W2T + j) = W2t[j]; }
    for (int j = 0; j < 9; ++j) NF(RL4_NLN_MPREV + j) = Mp[j];
    I[(int64_t)RL4_NLI_COOLDOWN * S] = cooldown; I[(int64_t)RL4_NLI_DIVERGED_STEP * S] = diverged_step; I[(int64_t)RL4_NLI_STEPP * S] = stepp;
    I[(int64_t)RL4_NLI_PYFLOAT_MASK * S] = pyfloat_mask;
#undef EF
#undef NF
}

}  // namespace rl4
