// Two-role pipeline form of the fused nonlinear kernel (sm_100a).  Included by nl_kernels.cu.
//
// Why.  One aircraft + agent carries ~1.7 KB of state (mixed policy) plus ~0.6 KB of Runge-Kutta stage vectors: with one
// agent per thread ptxas spills ~1.3 KB per thread, 75 % of the reloads miss L1 and are served by L2, and the kernel issues
// on a quarter of its cycles (profiles/prof_nl_mixed_r01f).  Quirk N1 of the reference (objects.py:1305-1310, 1530-1533:
// the action applied at step k+1 is pi(s_{k-1}) with the weights of step k-1's update) makes plant step k+1 independent
// of step k's critic / actor / model updates, so the step is cut into two ROLES that run concurrently, one step apart:
//
//   role P (plant)    warps 0..3 of the CTA, one agent per thread:  Ce500NonLinear.step(k)  (actuators, faults, 6-DOF
//                     plant, rewards, statistics)  and  RLS.update(k-1)             envs/nonlinear/env.py:182-256, objects.py:492-543
//   role N (networks) warps 4..11, TWO adjacent lanes per agent, each owning 5 of the 10 hidden units of the three
//                     networks: critic / target forward, TD error, critic VJP + SGD, Polyak, actor loss + update and
//                     _adapt_check of step k-1, then the two actor trace passes of step k (-> a_{k+1})     objects.py:1292-1399, 1212-1290
//
// The roles exchange ~20 values per agent and step through a double-buffered mailbox in shared memory (P -> N: s_next,
// reward gradient, NaN flag, F and G of the RLS model as of before its update; N -> P: the next action) and meet at ONE
// __syncthreads() per step, which also keeps the 12 warps of the SM walking their two instruction streams together.
// Per-thread state halves (P: plant + RLS, N: half of each network), nothing spills, 12 warps per SM instead of 8.
//
// Arithmetic is unchanged: every FMA chain keeps the term order of the one-thread kernel.  Chains that run over all 10
// hidden units (output layers, da/ds) are evaluated in two phases inside the lane pair -- lane 0 accumulates units 0..4,
// hands the partial sum to lane 1 with a shuffle, lane 1 continues with units 5..9 -- both lanes execute both phases
// (no divergence), each phase's result is taken from the lane for which it is meaningful.  The state planes in HBM are
// the same as the one-thread kernel's, so the two forms are interchangeable launch by launch; logging launches and
// the 'replacing' trace mode (a norm over all 50 trace elements in a fixed order) stay on the one-thread kernel.
#pragma once

namespace rl4 {

#ifndef RL4_PIPE_REGS_P
#define RL4_PIPE_REGS_P 0        // 0: one register budget for both roles; else setmaxnreg values for the P / N warpgroups
#define RL4_PIPE_REGS_N 0
#endif
#ifndef RL4_PIPE_AGENTS
#define RL4_PIPE_AGENTS 64       // agents per CTA (a multiple of 32): 192 threads, two CTAs per SM (measured: 32 = 64 > 128 by 2 %)
#endif
constexpr int kPipeAgents = RL4_PIPE_AGENTS;                       // agents per CTA: 4 P-warps (1 thread each) + 8 N-warps (2 lanes each)
constexpr int kPipeThreads = 3 * kPipeAgents;
constexpr int kPipeNLanes = 2 * kPipeAgents;

// shared-memory carve-up (elements are laid out [element][thread-of-role]: conflict-free, static element index)
template <typename TN> struct PipeSmem {
    // role N, per lane: own rows of the actor trace (5 + 20 doubles), own half of the target critic (20 + 15 TN)
    static constexpr int kEaLane = 25, kTgtLane = 35;
    // role P, per thread: RLS parameters (12) and covariance (16)
    static constexpr int kRlsThread = 28;
    // mailbox, per agent and buffer: P -> N  s_next[4] TN, F/G of the RLS model cast to TN [12], rg2 (double), nans (int);  N -> P  a_next TN
    static constexpr size_t bytes =
        sizeof(double) * ((size_t)kEaLane * kPipeNLanes + (size_t)kRlsThread * kPipeAgents + 2 * kPipeAgents) +
        sizeof(TN) * ((size_t)kTgtLane * kPipeNLanes + 2 * 16 * kPipeAgents + 2 * kPipeAgents) + sizeof(int) * 2 * kPipeAgents;
};

template <typename T> struct StridedN {                 // per-lane array in shared memory, [element][lane]
    T* p;
    __device__ __forceinline__ T& operator[](int j) const { return p[j * kPipeNLanes]; }
};
template <typename T> struct StridedP {                 // per-thread array of role P, [element][thread]
    T* p;
    __device__ __forceinline__ T& operator[](int j) const { return p[j * kPipeAgents]; }
};

// exchange with the partner lane of the pair.  The mask names the two lanes of the pair only: both always take the same
// path (same agent), while other pairs of the warp may be frozen (diverged agents) and skip the exchange.
__device__ __forceinline__ unsigned pair_mask() { return 3u << (threadIdx.x & 30); }
__device__ __forceinline__ float pair_xchg(float v) { return __shfl_xor_sync(pair_mask(), v, 1); }
__device__ __forceinline__ double pair_xchg(double v) { return __shfl_xor_sync(pair_mask(), v, 1); }

// An in-order chain over the 10 hidden units, term j = x[j] * w[j]: acc = x0 w0; acc = fma(x_j, w_j, acc), j = 1..9, with units
// 0..4 on lane 0 and 5..9 on lane 1 of the pair.  Both lanes run both phases; the value returned is the complete chain.
template <typename TN, typename FX, typename FW>
__device__ __forceinline__ TN pair_chain10(int half, FX x, FW w)
{
    TN a = x(0) * w(0);                                  // phase A: meaningful on lane 0
#pragma unroll
    for (int j = 1; j < 5; ++j) a = nfma<TN>(x(j), w(j), a);
    TN b = pair_xchg(a);                                 // lane 1 receives lane 0's partial sum
#pragma unroll
    for (int j = 0; j < 5; ++j) b = nfma<TN>(x(j), w(j), b);      // phase B: meaningful on lane 1
    const TN other = pair_xchg(b);
    return half ? b : other;
}

// hidden layer of the lane's 5 units: pre_j = s0 W1[0][j]; fma over the other inputs (Network.base_call, objects.py:111-139)
template <typename TN, typename WA>
__device__ __forceinline__ void pipe_hidden5(const TN (&s)[4], const WA W1, TN (&h)[5])
{
    Rn<TN> pre[5], out[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        TN acc = s[0] * W1[j];
#pragma unroll
        for (int i = 1; i < 4; ++i) acc = nfma<TN>(s[i], W1[i * 5 + j], acc);
        pre[j] = Rn<TN>(acc);
    }
    tanh_t13_n<5>(pre, out);
#pragma unroll
    for (int j = 0; j < 5; ++j) h[j] = out[j].v;
}

// Actor_big.call (objects.py:374-407) on the lane pair: forward + the lane's rows of the trace ('none' / 'accumulating')
template <typename TN, typename EA>
__device__ __forceinline__ TN pipe_actor(int half, const TN (&s)[4], const TN (&W1)[20], const TN (&W2)[5], const EA Ea, bool acc,
                                         double gl, TN (&h)[5], TN& ai1)
{
    pipe_hidden5<TN>(s, (const TN*)W1, h);
    const TN o = pair_chain10<TN>(half, [&](int j) { return h[j]; }, [&](int j) { return W2[j]; });
    const TN a = tanh_t13(Rn<TN>(o)).v;
    ai1 = TN(1) - a * a;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const double g = (double)(ai1 * h[j]);                                // objects.py:385
        Ea[j] = acc ? (Ea[j] * gl + g) : g;
        const TN v = (ai1 * W2[j]) * (TN(1) - h[j] * h[j]);                   // objects.py:386
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double gi = (double)(v * s[i]);
            Ea[5 + j * 4 + i] = acc ? (Ea[5 + j * 4 + i] * gl + gi) : gi;
        }
    }
    return a;
}

// the full 6-DOF step of the plant (asymmetric or out-of-range states: faults on aileron / rudder, a flight that is
// blowing up): one out-of-line copy working on a private array, so the symmetric-flight fast path keeps its registers
static __device__ __noinline__ void pipe_full_step(const rl4_cit_params* P, double* x, const double* u, double dt, int integ)
{
    if (integ == RL4_CIT_INTEGRATOR_RK4) rl4_cit_step_rk4(P, x, u, dt); else rl4_cit_step_ode5(P, x, u, dt);
}

// rl4_cit_step_auto (include/rl4_citation_surrogate.h) with the rare full-model path out of line
template <int INTEG>
__device__ __forceinline__ void pipe_plant_step(const rl4_cit_params& P, double (&x)[12], const double (&u)[11], double dt)
{
    if (rl4_cit_is_symmetric(x, u)) {
        const double s0 = x[RL4_CIT_Q], s1 = x[RL4_CIT_V], s2 = x[RL4_CIT_ALPHA], s3 = x[RL4_CIT_THETA], s4 = x[RL4_CIT_H], s5 = x[RL4_CIT_XE];
        if (INTEG == RL4_CIT_INTEGRATOR_RK4) rl4_cit_step_rk4_lon(&P, x, u, dt); else rl4_cit_step_ode5_lon(&P, x, u, dt);
        if (rl4_cit_lon_in_range(x)) return;
        x[RL4_CIT_Q] = s0; x[RL4_CIT_V] = s1; x[RL4_CIT_ALPHA] = s2; x[RL4_CIT_THETA] = s3; x[RL4_CIT_H] = s4; x[RL4_CIT_XE] = s5;
    }
    double xt[12], ut[11];
#pragma unroll
    for (int j = 0; j < 12; ++j) xt[j] = x[j];
#pragma unroll
    for (int j = 0; j < 11; ++j) ut[j] = u[j];
    pipe_full_step(&P, xt, ut, dt, INTEG);
#pragma unroll
    for (int j = 0; j < 12; ++j) x[j] = xt[j];
}

template <typename TN, int INTEG, bool PER_AGENT>
__global__ void __launch_bounds__(kPipeThreads, 128 / kPipeAgents)
nl_pipe_kernel(const __grid_constant__ rl4_nl_params p, const double* __restrict__ theta_ref, const float* __restrict__ noise,
               int64_t noise_stride, int k0, int n_steps, const rl4_nl_state st, int64_t n_agents)
{
    extern __shared__ __align__(16) unsigned char pipe_smem[];
    // ---- carve-up: doubles first
    double* sd = reinterpret_cast<double*>(pipe_smem);
    double* const sm_ea = sd;                       sd += PipeSmem<TN>::kEaLane * kPipeNLanes;
    double* const sm_rls = sd;                      sd += PipeSmem<TN>::kRlsThread * kPipeAgents;
    double* const mb_rg2 = sd;                      sd += 2 * kPipeAgents;
    TN* sn = reinterpret_cast<TN*>(sd);
    TN* const sm_tgt = sn;                          sn += PipeSmem<TN>::kTgtLane * kPipeNLanes;
    TN* const mb_pn = sn;                           sn += 2 * 16 * kPipeAgents;        // [buf][16 fields][agent]: s_next 0..3, F/G 4..15
    TN* const mb_a = sn;                            sn += 2 * kPipeAgents;
    int* const mb_nans = reinterpret_cast<int*>(sn);

    const int tid = threadIdx.x;
    const bool roleP = tid < kPipeAgents;
    const int64_t S = st.stride;
    const int k_end = k0 + n_steps;
    const bool f32 = sizeof(TN) == 4;

#if RL4_PIPE_REGS_P
    // warpgroup-level register re-allocation: the plant role keeps 12 states + six Runge-Kutta stage vectors in flight, the
    // network role holds 60 weights and little else.  128 x P + 256 x N registers must not exceed the CTA's allocation.
    if (roleP) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(RL4_PIPE_REGS_P));
    else       asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(RL4_PIPE_REGS_N));
#endif
    if (roleP) {
        // =========================== role P: plant + env wrapper + RLS + statistics ===========================
        const int la = tid;                                                    // agent slot inside the CTA
        const int64_t i_raw = (int64_t)blockIdx.x * kPipeAgents + la;
        const bool active = i_raw < n_agents;
        const int64_t i = active ? i_raw : n_agents - 1;                       // tail threads shadow the last agent (loads only)
        const NlHp<PER_AGENT> hv{p, i};
        double* __restrict__ E = st.env + i;
        const TN* __restrict__ Nn = (const TN*)st.net + i;
        int32_t* __restrict__ I = st.ints + i;
#define EF(f) E[(int64_t)(f) * S]
        const StridedP<double> th{sm_rls + la};
        const StridedP<double> cv{sm_rls + 12 * kPipeAgents + la};
        double x[12], x_act[3], x_lon[3], x_prev_lon[3], eps[3];
        for (int j = 0; j < 12; ++j) { x[j] = EF(RL4_NLE_XFULL + j); th[j] = EF(RL4_NLE_THETA + j); }
        for (int j = 0; j < 3; ++j) { x_act[j] = EF(RL4_NLE_XACT + j); x_lon[j] = EF(RL4_NLE_XLON + j); x_prev_lon[j] = EF(RL4_NLE_XPREVLON + j); eps[j] = EF(RL4_NLE_EPS + j); }
        for (int j = 0; j < 16; ++j) cv[j] = EF(RL4_NLE_COV + j);
        double eps_norm = EF(RL4_NLE_EPS_NORM), rse0 = EF(RL4_NLE_RSE), rse1 = EF(RL4_NLE_RSE + 1);
        double rse_f0 = EF(RL4_NLE_RSE_FLIGHT), rse_f1 = EF(RL4_NLE_RSE_FLIGHT + 1), nz_peak = EF(RL4_NLE_NZ_PEAK);
        const int flight_step = hv.hpi(RL4_NHPI_FLIGHT_STEP);
        TN a_cur = Nn[(int64_t)RL4_NLN_A * S], a_old = Nn[(int64_t)RL4_NLN_APREV * S];   // a_k and a_{k-1} of the next plant step
        int div = I[(int64_t)RL4_NLI_DIVERGED_STEP * S], stepp = I[(int64_t)RL4_NLI_STEPP * S];
        if (!active) div = 0;
        bool stepped = false;                      // the previous iteration executed a plant step (its RLS update is due)
        TN a_rls = TN(0), a_rls_prev = TN(0);      // a_{k-1}, a_{k-2} as seen by RLS(k-1)
        double xn_prev[3] = {0.0, 0.0, 0.0};       // x_next_lon of the previous plant step

        for (int k = k0; k <= k_end; ++k) {
            // ---- RLS.update of step k-1 and the shift of the model's regressors (objects.py:1521-1524, 1532-1534)
            if (stepped) {
                if (k - 1 > 0) {
                    double Xr[4], Y[3];
#pragma unroll
                    for (int ii = 0; ii < 3; ++ii) { Xr[ii] = x_lon[ii] - x_prev_lon[ii]; Y[ii] = xn_prev[ii] - x_lon[ii]; }
                    Xr[3] = (double)(a_rls - a_rls_prev);
                    nl_rls_update(th, cv, Xr, Y, hv.hp(RL4_NHP_RLS_GAMMA), eps, eps_norm);
                }
#pragma unroll
                for (int j = 0; j < 3; ++j) { x_prev_lon[j] = x_lon[j]; x_lon[j] = xn_prev[j]; }
                stepped = false;
            }
            // ---- Ce500NonLinear.step of step k
            if (k < k_end && div < 0) {
                if (k > k0) { a_old = a_cur; a_cur = mb_a[((k - 1) & 1) * kPipeAgents + la]; }      // a_k = actor forward of step k-1
                const double act[3] = {(double)a_cur, 0.0, 0.0};
                double surf[3], ueff[3], xo[12], e_phi, e_th, e_psi, reward, rg2;
                const double yref_k = __ldg(theta_ref + k);
                {   // nl_env_step with the rare full-model plant path out of line
                    const int fault_step = hv.hpi(RL4_NHPI_FAULT_STEP);
                    const bool faulted = (fault_step >= 0 && stepp >= fault_step);                 // env.py:132,151
                    const int damp = hv.hpi(RL4_NHPI_FAULT_DAMP), sat = hv.hpi(RL4_NHPI_FAULT_SAT);
                    const double omega = (faulted && damp == RL4_NL_SLOW_ALL && stepp > fault_step) ? p.omega_slow : p.omega0;
                    double eff[11];
#pragma unroll
                    for (int j = 0; j < 11; ++j) eff[j] = 0.0;
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const double hi = p.limit_deg[j], lo = -p.limit_deg[j];
                        double v = act[j] * (hi - lo) / 2.0;                                       // _scale_action env.py:111-124
                        v = v + (hi + lo) / 2.0;
                        const double cmd = v * (3.14159265358979323846 / 180.0);
                        double d = cmd - x_act[j];                                                 // _propagate_surfaces_states env.py:161-180
                        d = d * omega;
                        d = d < -p.rate_limit ? -p.rate_limit : (d > p.rate_limit ? p.rate_limit : d);
                        x_act[j] = x_act[j] + p.dt * d;
                        surf[j] = x_act[j];
                    }
                    if (faulted && sat != RL4_NL_SAT_NONE) {                                       // _saturate_surfaces env.py:150-159
#pragma unroll
                        for (int j = 0; j < 3; ++j)
                            if (sat - 1 == j) { const double L = p.sat_limit[j]; surf[j] = surf[j] < -L ? -L : (surf[j] > L ? L : surf[j]); }
                    }
#pragma unroll
                    for (int j = 0; j < 3; ++j) eff[j] = surf[j];
                    if (faulted) {                                                                 // _engage_fault env.py:129-148
                        const double f = hv.hp(RL4_NHP_DAMP_FACTOR);
                        if (damp == RL4_NL_DAMP_ELEVATOR || damp == RL4_NL_DAMP_ALL) eff[0] *= f;
                        if (damp == RL4_NL_DAMP_AILERON || damp == RL4_NL_DAMP_ALL) eff[1] *= f;
                        if (damp == RL4_NL_DAMP_RUDDER || damp == RL4_NL_DAMP_ALL) eff[2] *= f;
                        if (damp == RL4_NL_SHIFT_CG) eff[10] = hv.hp(RL4_NHP_CG_SHIFT);
                    }
                    double u[11];
#pragma unroll
                    for (int j = 0; j < 11; ++j) u[j] = p.trim_input[j] + eff[j];                  // env.py:207-208
                    ueff[0] = u[0]; ueff[1] = u[1]; ueff[2] = u[2];
                    // env.py:210  x_full = model.step(input): the plant returns the state BEFORE the step (xo) and then
                    // integrates the carried state (x) -- output-then-update, as the reference's binary does
#pragma unroll
                    for (int j = 0; j < 12; ++j) xo[j] = x[j];
                    pipe_plant_step<INTEG>(p.plant, x, u, p.dt);
                    const double Q = hv.hp(RL4_NHP_Q_SYM);
                    e_phi = xo[6] - 0.0; e_th = xo[7] - yref_k; e_psi = xo[8] - 0.0;               // env.py:215 (state - ref)
                    reward = (-0.5 * Q) * (e_th * e_th);                                           // env.py:218
                    rg2 = (-Q) * e_th;                                                             // env.py:219-220 (q slot)
                    (void)reward; (void)ueff;
                }
                stepp += 1;
                xn_prev[0] = xo[4]; xn_prev[1] = xo[7]; xn_prev[2] = xo[1];                        // env.py:231
                bool nans = false;
#pragma unroll
                for (int j = 0; j < 12; ++j) nans |= (xo[j] != xo[j]);
                const double rse_k0 = sqrt_of_square(e_th), rse_k1 = nsqrt(e_phi * e_phi + e_psi * e_psi);   // env.py:251
                rse0 += rse_k0;                                                                    // objects.py:1503-1504
                rse1 += rse_k1;
                if (k >= flight_step) { rse_f0 += rse_k0; rse_f1 += rse_k1; }                      // functions.py:917,1039
                { const double nz = fabs(xo[3] * xo[1] / 9.80665); if (nz > nz_peak) nz_peak = nz; } // functions.py:774,1055
                // ---- mailbox for role N's work on step k (read in the next iteration)
                TN* mb = mb_pn + (size_t)(k & 1) * 16 * kPipeAgents + la;
                mb[0 * kPipeAgents] = (TN)xo[4]; mb[1 * kPipeAgents] = (TN)xo[7]; mb[2 * kPipeAgents] = (TN)xo[1]; mb[3 * kPipeAgents] = (TN)e_th;   // env.py:236-238
#pragma unroll
                for (int j = 0; j < 12; ++j) mb[(4 + j) * kPipeAgents] = (TN)th[j];                // F, G BEFORE this step's RLS update (objects.py:1327; Q18)
                mb_rg2[(k & 1) * kPipeAgents + la] = rg2;
                mb_nans[(k & 1) * kPipeAgents + la] = nans ? 1 : 0;
                a_rls = a_cur; a_rls_prev = a_old;
                stepped = true;
                if (nans) div = k;                                                                 // objects.py:1557
            }
            __syncthreads();
        }
        if (active) {
            for (int j = 0; j < 12; ++j) { EF(RL4_NLE_XFULL + j) = x[j]; EF(RL4_NLE_THETA + j) = th[j]; }
            for (int j = 0; j < 3; ++j) { EF(RL4_NLE_XACT + j) = x_act[j]; EF(RL4_NLE_XLON + j) = x_lon[j]; EF(RL4_NLE_XPREVLON + j) = x_prev_lon[j]; EF(RL4_NLE_EPS + j) = eps[j]; }
            for (int j = 0; j < 16; ++j) EF(RL4_NLE_COV + j) = cv[j];
            EF(RL4_NLE_EPS_NORM) = eps_norm; EF(RL4_NLE_RSE) = rse0; EF(RL4_NLE_RSE + 1) = rse1;
            EF(RL4_NLE_RSE_FLIGHT) = rse_f0; EF(RL4_NLE_RSE_FLIGHT + 1) = rse_f1; EF(RL4_NLE_NZ_PEAK) = nz_peak;
            I[(int64_t)RL4_NLI_DIVERGED_STEP * S] = div; I[(int64_t)RL4_NLI_STEPP * S] = stepp;
        }
#undef EF
        return;
    }

    // =========================== role N: networks, two lanes per agent ===========================
    const int nl = tid - kPipeAgents;                                          // lane index among the role's 256 lanes
    const int la = nl >> 1, half = nl & 1;
    const int64_t i_raw = (int64_t)blockIdx.x * kPipeAgents + la;
    const bool active = i_raw < n_agents;
    const int64_t i = active ? i_raw : n_agents - 1;
    const NlHp<PER_AGENT> hv{p, i};
    double* __restrict__ E = st.env + i;
    TN* __restrict__ Nn = (TN*)st.net + i;
    int32_t* __restrict__ I = st.ints + i;
#define EF(f) E[(int64_t)(f) * S]
#define NF(f) Nn[(int64_t)(f) * S]
    const int u0 = half * 5;                                                   // first hidden unit owned by this lane
    const StridedN<double> Ea{sm_ea + nl};                                     // [0..5): E[u0 + j];  [5 + j*4 + i]: E[10 + (u0 + j)*4 + i]
    const StridedN<TN> W1t{sm_tgt + nl};                                       // [i*5 + j]
    const StridedN<TN> W2t{sm_tgt + 20 * kPipeNLanes + nl};                    // [j*3 + q]
    TN s[4], s_prev[4], W1a[20], W2a[5], W1c[20], W2c[15], Mp[9];
    for (int j = 0; j < 5; ++j) {
        Ea[j] = EF(RL4_NLE_EA + u0 + j);
        W2a[j] = NF(RL4_NLN_W2A + u0 + j);
        for (int ii = 0; ii < 4; ++ii) {
            Ea[5 + j * 4 + ii] = EF(RL4_NLE_EA + 10 + (u0 + j) * 4 + ii);
            W1a[ii * 5 + j] = NF(RL4_NLN_W1A + ii * 10 + u0 + j);
            W1c[ii * 5 + j] = NF(RL4_NLN_W1C + ii * 10 + u0 + j);
            W1t[ii * 5 + j] = NF(RL4_NLN_W1T + ii * 10 + u0 + j);
        }
        for (int q = 0; q < 3; ++q) { W2c[j * 3 + q] = NF(RL4_NLN_W2C + (u0 + j) * 3 + q); W2t[j * 3 + q] = NF(RL4_NLN_W2T + (u0 + j) * 3 + q); }
    }
    for (int j = 0; j < 4; ++j) { s[j] = NF(RL4_NLN_S + j); s_prev[j] = NF(RL4_NLN_SPREV + j); }
    for (int j = 0; j < 9; ++j) Mp[j] = NF(RL4_NLN_MPREV + j);
    TN a = NF(RL4_NLN_A), a_prev = NF(RL4_NLN_APREV), lr_a = NF(RL4_NLN_LR_A), lr_c = NF(RL4_NLN_LR_C);
    double cgp2 = EF(RL4_NLE_CGRAD_PREV), eta_a = EF(RL4_NLE_ETA_A), eta_c = EF(RL4_NLE_ETA_C), lambdaa = EF(RL4_NLE_LAMBDAA), gl = EF(RL4_NLE_GL);
    int cooldown = I[(int64_t)RL4_NLI_COOLDOWN * S], pyfloat_mask = I[(int64_t)RL4_NLI_PYFLOAT_MASK * S];
    int div = I[(int64_t)RL4_NLI_DIVERGED_STEP * S];
    if (!active) div = 0;
    // carried from the actor passes of step k to the rest of step k (next iteration)
    TN a_next = TN(0), a_random = TN(0), dads[4] = {TN(0), TN(0), TN(0), TN(0)};
    bool fwd_done = false;

    for (int k = k0; k <= k_end; ++k) {
        // ---- the rest of step kk = k - 1 (objects.py:1292-1399): needs the plant's step kk from the mailbox
        if (fwd_done) {
            const int kk = k - 1;
            const TN* mb = mb_pn + (size_t)(kk & 1) * 16 * kPipeAgents + la;
            const TN s_next[4] = {mb[0], mb[1 * kPipeAgents], mb[2 * kPipeAgents], mb[3 * kPipeAgents]};
            const double rg2 = mb_rg2[(kk & 1) * kPipeAgents + la];
            const bool nans = mb_nans[(kk & 1) * kPipeAgents + la] != 0;
            const TN a_k = a;
            // critic(s_prev), target_critic(s_next)
            TN hc[5], ht[5], lam[3], lt[3];
            pipe_hidden5<TN>(s_prev, (const TN*)W1c, hc);
            pipe_hidden5<TN>(s_next, W1t, ht);
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                lam[q] = pair_chain10<TN>(half, [&](int j) { return hc[j]; }, [&](int j) { return W2c[j * 3 + q]; });
                lt[q] = pair_chain10<TN>(half, [&](int j) { return ht[j]; }, [&](int j) { return (TN)W2t[j * 3 + q]; });
            }
            TN M[9], Gn[3];                                                            // objects.py:1327-1329
#pragma unroll
            for (int ii = 0; ii < 3; ++ii) Gn[ii] = mb[(4 + 9 + ii) * kPipeAgents];
#pragma unroll
            for (int ii = 0; ii < 3; ++ii)
#pragma unroll
                for (int j = 0; j < 3; ++j) M[ii * 3 + j] = mb[(4 + j * 3 + ii) * kPipeAgents] + Gn[ii] * dads[j];

            if (kk > 0) {
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
                {   // critic VJP (tape.gradient with output_gradients = td, objects.py:1365) + SGD (:1368), the lane's units
                    TN dpre[5];
#pragma unroll
                    for (int j = 0; j < 5; ++j) {
                        TN dh = td[0] * W2c[j * 3];
                        dh = nfma<TN>(td[1], W2c[j * 3 + 1], dh);
                        dh = nfma<TN>(td[2], W2c[j * 3 + 2], dh);
                        dpre[j] = dh * (TN(1) - hc[j] * hc[j]);
                    }
#pragma unroll
                    for (int j = 0; j < 5; ++j)
#pragma unroll
                        for (int q = 0; q < 3; ++q) W2c[j * 3 + q] = W2c[j * 3 + q] - lr_c * (hc[j] * td[q]);
#pragma unroll
                    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
                        for (int j = 0; j < 5; ++j) W1c[ii * 5 + j] = W1c[ii * 5 + j] - lr_c * (s_prev[ii] * dpre[j]);
                }
                {   // target soft update (objects.py:1371)
                    const double tau_d = hv.hp(RL4_NHP_TAU);
                    const TN omt = (TN)(1.0 - tau_d), tt = (TN)tau_d;
#pragma unroll
                    for (int j = 0; j < 20; ++j) W1t[j] = omt * W1t[j] + tt * W1c[j];
#pragma unroll
                    for (int j = 0; j < 15; ++j) W2t[j] = omt * W2t[j] + tt * W2c[j];
                }
                {   // actor (objects.py:1379-1388): smoothness terms, loss, update of the lane's units (both trace passes ran before)
                    const TN dT = a_k - a_next, dS = a_k - a_random;
                    const TN L_T = sqrt_of_square(dT), L_S = sqrt_of_square(dS);
                    TN v[3];
#pragma unroll
                    for (int j = 0; j < 3; ++j) v[j] = -((TN)rg[j] + gam * lt[j]);
                    TN acc = v[0] * Gn[0]; acc = nfma<TN>(v[1], Gn[1], acc); acc = nfma<TN>(v[2], Gn[2], acc);
                    const TN loss = (acc + (TN)hv.hp(RL4_NHP_LAMBDA_T) * L_T) + (TN)hv.hp(RL4_NHP_LAMBDA_S) * L_S;
#pragma unroll
                    for (int j = 0; j < 5; ++j) W2a[j] = W2a[j] - lr_a * (loss * (TN)Ea[j]);                   // objects.py:417-427
#pragma unroll
                    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
                        for (int j = 0; j < 5; ++j) W1a[ii * 5 + j] = W1a[ii * 5 + j] - lr_a * (loss * (TN)Ea[5 + j * 4 + ii]);
                }
                {   // _adapt_check (objects.py:1212-1290)
                    const bool cond1 = kk < hv.hpi(RL4_NHPI_WARMUP_STEPS);
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
            cgp2 = rg2;
#pragma unroll
            for (int j = 0; j < 9; ++j) Mp[j] = M[j];
            if (nans) div = kk;
            fwd_done = false;
        }
        // ---- the actor passes of step k: actor(s_prev) with trace pass 1 (objects.py:1310) and, for k > 0, actor(s_random)
        // with trace pass 2 (objects.py:1375-1378), through ONE copy of the code; da/ds by reverse mode (objects.py:1323)
        if (k < k_end && div < 0) {
            const bool acc = hv.hpi(RL4_NHPI_ELIG_A) == RL4_ELIG_ACCUMULATING;
            const TN nz = (TN)__ldg(noise + (int64_t)(k - k0) * noise_stride + i);
            const int n_pass = (k > 0) ? 2 : 1;
            TN ha[5], ai1 = TN(0);
#pragma unroll 1
            for (int pass = 0; pass < n_pass; ++pass) {
                TN sin[4], hh[5], aa1;
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) sin[ii] = pass ? (nz * (TN)p.noise_std[ii] + s_prev[ii]) : s_prev[ii];
                const TN aout = pipe_actor<TN>(half, sin, W1a, W2a, Ea, acc, gl, hh, aa1);
                if (pass == 0) {
                    a_next = aout; ai1 = aa1;
#pragma unroll
                    for (int j = 0; j < 5; ++j) ha[j] = hh[j];
                } else {
                    a_random = aout;
                }
            }
            {
                const TN g_o = TN(1) * ai1;
                TN gp[5];
#pragma unroll
                for (int j = 0; j < 5; ++j) gp[j] = (g_o * W2a[j]) * (TN(1) - ha[j] * ha[j]);
#pragma unroll
                for (int ii = 0; ii < 4; ++ii)
                    dads[ii] = pair_chain10<TN>(half, [&](int j) { return gp[j]; }, [&](int j) { return W1a[ii * 5 + j]; });
            }
            if (half == 0) mb_a[(k & 1) * kPipeAgents + la] = a_next;                  // the action of plant step k + 1
            fwd_done = true;
        }
        __syncthreads();
    }

    if (!active) return;
    for (int j = 0; j < 5; ++j) {
        EF(RL4_NLE_EA + u0 + j) = Ea[j];
        NF(RL4_NLN_W2A + u0 + j) = W2a[j];
        for (int ii = 0; ii < 4; ++ii) {
            EF(RL4_NLE_EA + 10 + (u0 + j) * 4 + ii) = Ea[5 + j * 4 + ii];
            NF(RL4_NLN_W1A + ii * 10 + u0 + j) = W1a[ii * 5 + j];
            NF(RL4_NLN_W1C + ii * 10 + u0 + j) = W1c[ii * 5 + j];
            NF(RL4_NLN_W1T + ii * 10 + u0 + j) = W1t[ii * 5 + j];
        }
        for (int q = 0; q < 3; ++q) { NF(RL4_NLN_W2C + (u0 + j) * 3 + q) = W2c[j * 3 + q]; NF(RL4_NLN_W2T + (u0 + j) * 3 + q) = W2t[j * 3 + q]; }
    }
    if (half == 0) {                                                                   // quantities both lanes hold: one of them stores
        for (int j = 0; j < 4; ++j) { NF(RL4_NLN_S + j) = s[j]; NF(RL4_NLN_SPREV + j) = s_prev[j]; }
        for (int j = 0; j < 9; ++j) NF(RL4_NLN_MPREV + j) = Mp[j];
        NF(RL4_NLN_A) = a; NF(RL4_NLN_APREV) = a_prev; NF(RL4_NLN_LR_A) = lr_a; NF(RL4_NLN_LR_C) = lr_c;
        EF(RL4_NLE_CGRAD_PREV) = cgp2; EF(RL4_NLE_ETA_A) = eta_a; EF(RL4_NLE_ETA_C) = eta_c; EF(RL4_NLE_LAMBDAA) = lambdaa; EF(RL4_NLE_GL) = gl;
        I[(int64_t)RL4_NLI_COOLDOWN * S] = cooldown; I[(int64_t)RL4_NLI_PYFLOAT_MASK * S] = pyfloat_mask;
    }
#undef EF
#undef NF
}

}  // namespace rl4
