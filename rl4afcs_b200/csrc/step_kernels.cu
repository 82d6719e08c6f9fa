// Step-API and episode-statistics kernels that are not part of the fused loops (sm_100a):
//
//   soft_update_kernel           Critic / Actor / Critic_big.soft_update   objects.py:207-215, 273-281, 353-361
//   actor_weight_update_kernel   Actor / Actor_big.get_weight_update       objects.py:261-271, 417-427
//   sp_agent_stats_kernel        MC_run_seed's per-run numbers             functions.py:53,57; utils.py:350-369 (+ nMAE)
//   nl_agent_stats_kernel        MC_test_hparam's per-run numbers          functions.py:916-917,1036-1039,1050-1055
//   stats_partial / stats_final  MC_run's reduction over runs              functions.py:161-182
//
// All HBM-bound element-wise work over SoA planes (agent index fastest, one coalesced line per warp and field).
#include "rl4_math.cuh"
#include "rl4_runtime.h"
#include "../../include/rl4afcs_b200.h"

namespace rl4 {

// target <- (1 - tau) * target + tau * source: three separately rounded operations per weight (TF: two multiplies, one add)
template <typename TN>
__global__ void __launch_bounds__(256)
soft_update_kernel(TN* __restrict__ target, const TN* __restrict__ source, double tau, int n_rows, int64_t stride, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Rn<TN> omt = Rn<TN>(TN(__dsub_rn(1.0, tau))), tt = Rn<TN>(TN(tau));
    for (int r = 0; r < n_rows; ++r) {
        const int64_t o = (int64_t)r * stride + i;
        target[o] = (omt * Rn<TN>(target[o]) + tt * Rn<TN>(source[o])).v;
    }
}

// out[r] = loss * (TN) E[r]   (the trace is stored in TE and cast to the tensor dtype, Q15)
template <typename TN, typename TE>
__global__ void __launch_bounds__(256)
actor_weight_update_kernel(const TN* __restrict__ loss, const TE* __restrict__ E, TN* __restrict__ out, int n_rows,
                           int64_t stride, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Rn<TN> l = Rn<TN>(loss[i]);
    for (int r = 0; r < n_rows; ++r) {
        const int64_t o = (int64_t)r * stride + i;
        out[o] = (l * cvt<TN>(Rn<TE>(E[o]))).v;
    }
}

template <typename TE>
__global__ void __launch_bounds__(256)
sp_agent_stats_kernel(const __grid_constant__ rl4_sp_params p, const TE* __restrict__ env, const int32_t* __restrict__ ints,
                      int64_t stride, int64_t n, int n_steps, double ref_base_min, double ref_base_max,
                      double* __restrict__ out, int64_t out_stride)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double kappa = p.hp_agent[RL4_HP_KAPPA] ? __ldg(p.hp_agent[RL4_HP_KAPPA] + i) : p.hp[RL4_HP_KAPPA];
    const double amp = p.hp_agent[RL4_HP_REF_AMP] ? __ldg(p.hp_agent[RL4_HP_REF_AMP] + i) : p.hp[RL4_HP_REF_AMP];
    const double sum_c = (double)env[(int64_t)RL4_SPE_SUM_C * stride + i];
    const double sum_abs_e = (double)env[(int64_t)RL4_SPE_SUM_ABS_E * stride + i];
    const int conv = ints[(int64_t)RL4_SPI_CONV_STEP * stride + i];
    const int div_step = ints[(int64_t)RL4_SPI_DIVERGED_STEP * stride + i];
    const int flags = ints[(int64_t)RL4_SPI_FLAGS * stride + i];
    const double conv_time = __dmul_rn((double)conv, p.dt);                                   // utils.py:366-368
    const double mean_abs_e = __ddiv_rn(sum_abs_e, (double)(n_steps > 0 ? n_steps : 1));
    // reference range max(ref) - min(ref) with ref_k = amp * base_k (rounding is monotone: the extremes are the products
    // of the extremes); nMAE = mean|e| / range -- an addition of this repo (BASELINE.json names it, the reference has none)
    const double range = fabs(__dsub_rn(__dmul_rn(amp, ref_base_max), __dmul_rn(amp, ref_base_min)));
    out[(int64_t)RL4_SPS_SUM_C * out_stride + i] = __ddiv_rn(sum_c, kappa);                   // functions.py:53
    out[(int64_t)RL4_SPS_CONV_TIME * out_stride + i] = conv_time;
    out[(int64_t)RL4_SPS_DIVERGED * out_stride + i] = (div_step >= 0 || (flags & RL4_SPF_X_NAN)) ? 1.0 : 0.0;   // functions.py:162
    out[(int64_t)RL4_SPS_UNSTEADY * out_stride + i] = (conv_time > 30.0) ? 1.0 : 0.0;        // functions.py:165
    out[(int64_t)RL4_SPS_MEAN_ABS_E * out_stride + i] = mean_abs_e;
    out[(int64_t)RL4_SPS_NMAE * out_stride + i] = __ddiv_rn(mean_abs_e, range);
}

__global__ void __launch_bounds__(256)
nl_agent_stats_kernel(const double* __restrict__ env, const int32_t* __restrict__ ints, int64_t stride, int64_t n,
                      double* __restrict__ out, int64_t out_stride)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double rse = env[(int64_t)RL4_NLE_RSE * stride + i], rse_f = env[(int64_t)RL4_NLE_RSE_FLIGHT * stride + i];
    out[(int64_t)RL4_NLS_RSE_WARMUP * out_stride + i] = __dsub_rn(rse, rse_f);               // functions.py:1036
    out[(int64_t)RL4_NLS_RSE_FLIGHT * out_stride + i] = rse_f;                                // functions.py:1037
    out[(int64_t)RL4_NLS_RSE_LAT * out_stride + i] = env[(int64_t)(RL4_NLE_RSE + 1) * stride + i];
    out[(int64_t)RL4_NLS_NZ_PEAK * out_stride + i] = env[(int64_t)RL4_NLE_NZ_PEAK * stride + i];
    out[(int64_t)RL4_NLS_DIVERGED * out_stride + i] = ints[(int64_t)RL4_NLI_DIVERGED_STEP * stride + i] >= 0 ? 1.0 : 0.0;
}

// ---- deterministic two-stage reduction: per-block partial sums in a fixed order, then one block adds the partials in
// index order (no atomics: the result does not depend on scheduling)
constexpr int kStatMaxFields = 8;
constexpr int kStatMaxBlocks = 1024;
constexpr int kStatRow = 2 * kStatMaxFields + 2;     // per field: sum over kept agents, sum over all; then n_kept, n_excluded

__global__ void __launch_bounds__(256)
stats_partial_kernel(const double* __restrict__ planes, int64_t stride, int n_fields, const double* __restrict__ exclude, int64_t n,
                     double* __restrict__ work)
{
    double acc[kStatRow];
#pragma unroll
    for (int j = 0; j < kStatRow; ++j) acc[j] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const bool ex = exclude && exclude[i] != 0.0;
#pragma unroll
        for (int f = 0; f < kStatMaxFields; ++f) {
            if (f < n_fields) {
                const double v = planes[(int64_t)f * stride + i];
                if (!ex) acc[2 * f] += v;
                acc[2 * f + 1] += v;
            }
        }
        acc[2 * kStatMaxFields] += ex ? 0.0 : 1.0;
        acc[2 * kStatMaxFields + 1] += ex ? 1.0 : 0.0;
    }
    __shared__ double sm[8][kStatRow];
#pragma unroll
    for (int j = 0; j < kStatRow; ++j) {
        double v = acc[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5][j] = v;
    }
    __syncthreads();
    if (threadIdx.x < kStatRow) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += sm[w][threadIdx.x];
        work[(int64_t)blockIdx.x * kStatRow + threadIdx.x] = v;
    }
}

__global__ void __launch_bounds__(32)
stats_final_kernel(const double* __restrict__ work, int n_blocks, int n_fields, double* __restrict__ out)
{
    const int j = threadIdx.x;
    if (j >= kStatRow) return;
    double v = 0.0;
    for (int b = 0; b < n_blocks; ++b) v += work[(int64_t)b * kStatRow + j];
    if (j < 2 * n_fields) out[j] = v;
    else if (j >= 2 * kStatMaxFields) out[2 * n_fields + (j - 2 * kStatMaxFields)] = v;
}

}  // namespace rl4

using namespace rl4;

extern "C" {

int rl4_soft_update(int policy, void* target, const void* source, double tau, int32_t n_rows, int64_t stride, int64_t n, void* stream)
{
    RL4_REQUIRE(target && source, "NULL argument");
    RL4_REQUIRE(n >= 0 && stride >= n && n_rows >= 0, "bad size");
    if (n == 0 || n_rows == 0) return 0;
    const unsigned grid = (unsigned)((n + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
    if (policy == RL4_FP64) soft_update_kernel<double><<<grid, 256, 0, s>>>((double*)target, (const double*)source, tau, n_rows, stride, n);
    else if (policy == RL4_FP32 || policy == RL4_MIXED) soft_update_kernel<float><<<grid, 256, 0, s>>>((float*)target, (const float*)source, tau, n_rows, stride, n);
    else { set_error("rl4_soft_update: unknown policy %d", policy); return -1; }
    return check_launch("soft_update_kernel");
}

int rl4_actor_weight_update(int policy, const void* loss, const void* E, void* out, int32_t n_rows, int64_t stride, int64_t n, void* stream)
{
    RL4_REQUIRE(loss && E && out, "NULL argument");
    RL4_REQUIRE(n >= 0 && stride >= n && n_rows >= 0, "bad size");
    if (n == 0 || n_rows == 0) return 0;
    const unsigned grid = (unsigned)((n + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
    switch (policy) {
    case RL4_FP64:  actor_weight_update_kernel<double, double><<<grid, 256, 0, s>>>((const double*)loss, (const double*)E, (double*)out, n_rows, stride, n); break;
    case RL4_FP32:  actor_weight_update_kernel<float, float><<<grid, 256, 0, s>>>((const float*)loss, (const float*)E, (float*)out, n_rows, stride, n); break;
    case RL4_MIXED: actor_weight_update_kernel<float, double><<<grid, 256, 0, s>>>((const float*)loss, (const double*)E, (float*)out, n_rows, stride, n); break;
    default: set_error("rl4_actor_weight_update: unknown policy %d", policy); return -1;
    }
    return check_launch("actor_weight_update_kernel");
}

int rl4_sp_agent_stats(int policy, const rl4_sp_params* p, rl4_sp_state st, int64_t n, int32_t n_steps, double ref_base_min,
                       double ref_base_max, double* out, int64_t out_stride, void* stream)
{
    RL4_REQUIRE(p && st.env && st.ints && out, "NULL argument");
    RL4_REQUIRE(n >= 0 && st.stride >= n && out_stride >= n && n_steps >= 0, "bad size");
    if (n == 0) return 0;
    const unsigned grid = (unsigned)((n + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
    if (policy == RL4_FP32)
        sp_agent_stats_kernel<float><<<grid, 256, 0, s>>>(*p, (const float*)st.env, st.ints, st.stride, n, n_steps, ref_base_min, ref_base_max, out, out_stride);
    else if (policy == RL4_FP64 || policy == RL4_MIXED)
        sp_agent_stats_kernel<double><<<grid, 256, 0, s>>>(*p, (const double*)st.env, st.ints, st.stride, n, n_steps, ref_base_min, ref_base_max, out, out_stride);
    else { set_error("rl4_sp_agent_stats: unknown policy %d", policy); return -1; }
    return check_launch("sp_agent_stats_kernel");
}

int rl4_nl_agent_stats(rl4_nl_state st, int64_t n, double* out, int64_t out_stride, void* stream)
{
    RL4_REQUIRE(st.env && st.ints && out, "NULL argument");
    RL4_REQUIRE(n >= 0 && st.stride >= n && out_stride >= n, "bad size");
    if (n == 0) return 0;
    nl_agent_stats_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(st.env, st.ints, st.stride, n, out, out_stride);
    return check_launch("nl_agent_stats_kernel");
}

int64_t rl4_stats_reduce_work_doubles(void) { return (int64_t)kStatMaxBlocks * kStatRow; }

int rl4_stats_reduce(const double* planes, int64_t stride, int32_t n_fields, const double* exclude, int64_t n, double* out,
                     double* work, int64_t work_doubles, void* stream)
{
    RL4_REQUIRE(planes && out && work, "NULL argument");
    RL4_REQUIRE(n >= 0 && stride >= n && n_fields >= 1 && n_fields <= kStatMaxFields, "bad size (1 <= n_fields <= 8)");
    RL4_REQUIRE(work_doubles >= rl4_stats_reduce_work_doubles(), "work buffer smaller than rl4_stats_reduce_work_doubles()");
    cudaStream_t s = (cudaStream_t)stream;
    int64_t blocks = (n + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > kStatMaxBlocks) blocks = kStatMaxBlocks;
    stats_partial_kernel<<<(unsigned)blocks, 256, 0, s>>>(planes, stride, n_fields, exclude, n, work);
    int rc = check_launch("stats_partial_kernel");
    if (rc) return rc;
    stats_final_kernel<<<1, 32, 0, s>>>(work, (int)blocks, n_fields, out);
    return check_launch("stats_final_kernel");
}

int64_t rl4_sizeof_sp_params(void) { return (int64_t)sizeof(rl4_sp_params); }
int64_t rl4_sizeof_nl_params(void) { return (int64_t)sizeof(rl4_nl_params); }

}  // extern "C"
