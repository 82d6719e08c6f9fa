// Library plumbing: error text, device check, launch counter, FMA peak micro-benchmark.
#include "rl4_runtime.h"
#include "rl4_math.cuh"
#include "../../include/rl4afcs_b200.h"
#include <cstdarg>
#include <cstdio>

namespace rl4 {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launch_count{0};

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what)
{
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    (void)cudaGetLastError();        // a non-sticky error (e.g. out of memory) must not be reported again by the next launch check
    return (int)e;
}

// 16 independent dependent-FMA chains per thread: enough ILP to saturate the pipe at
// modest occupancy.  Used only as the measured denominator of the pipe roofline.
template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* out, int iters, T a, T b)
{
    T acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = (T)(threadIdx.x + j) * (T)1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = ::fma(acc[j], a, b);
    }
    T s = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += acc[j];
    if (s == (T)123456789) out[0] = s;   // never true; keeps the chains alive
}

// The same with three DISTINCT register operands per FMA (two ping-pong sets of 16 values): what the pipe sustains when
// no operand is shared between instructions, as in the agent kernels (64-bit operands: six register reads per DFMA).
template <typename T>
__global__ void __launch_bounds__(256) fma3_peak_kernel(T* out, int iters)
{
    T a[16], b[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = (T)(threadIdx.x + j) * (T)1e-4;
    for (int it = 0; it < iters; it += 2) {
#pragma unroll
        for (int j = 0; j < 16; ++j) b[j] = ::fma(a[j], a[(j + 5) & 15], a[(j + 11) & 15]);
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = ::fma(b[j], b[(j + 3) & 15], b[(j + 9) & 15]);
    }
    T s = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += a[j];
    if (s == (T)123456789) out[0] = s;
}

template <typename T, bool DISTINCT = false>
static int run_peak(double* out_flops, cudaStream_t stream)
{
    int dev = 0, sms = 0;
    RL4_CUDA(cudaGetDevice(&dev));
    RL4_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    T* d = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = 0;
    double best = 0.0;
    const int blocks = sms * 8, threads = 256, iters = 4096;
    // single cleanup path: the buffer and both events are released on every exit
    auto fail = [&](cudaError_t e, const char* what) { if (e != cudaSuccess && rc == 0) rc = cuda_fail(e, what); return rc != 0; };
    if (fail(cudaMalloc(&d, sizeof(T)), "cudaMalloc") || fail(cudaEventCreate(&e0), "cudaEventCreate") ||
        fail(cudaEventCreate(&e1), "cudaEventCreate"))
        goto done;
    for (int rep = 0; rep < 6; ++rep) {
        if (fail(cudaEventRecord(e0, stream), "cudaEventRecord")) goto done;
        if (DISTINCT) fma3_peak_kernel<T><<<blocks, threads, 0, stream>>>(d, iters);
        else fma_peak_kernel<T><<<blocks, threads, 0, stream>>>(d, iters, (T)0.999, (T)1e-3);
        rc = check_launch("fma_peak_kernel");
        if (rc) goto done;
        if (fail(cudaEventRecord(e1, stream), "cudaEventRecord") || fail(cudaEventSynchronize(e1), "cudaEventSynchronize")) goto done;
        float ms = 0.f;
        if (fail(cudaEventElapsedTime(&ms, e0, e1), "cudaEventElapsedTime")) goto done;
        const double flops = 2.0 * 16.0 * iters * (double)blocks * threads / (ms * 1e-3);
        if (rep > 0 && flops > best) best = flops;
    }
    *out_flops = best;
done:
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(d);
    return rc;
}

// element-wise probes of the arithmetic primitives (tests/test_gpu_math.py)
__global__ void math_probe_kernel(int op, const void* a, const void* b, void* out, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    switch (op) {
    case 0: ((double*)out)[i] = tanh_t13(Rn<double>(((const double*)a)[i])).v; break;
    case 1: ((float*)out)[i] = tanh_t13(Rn<float>(((const float*)a)[i])).v; break;
    case 2: { const double d = ((const double*)b)[i]; ((double*)out)[i] = div_rn_shared(((const double*)a)[i], d, rcp_refined(d)); break; }
    case 3: ((double*)out)[i] = __ddiv_rn(((const double*)a)[i], ((const double*)b)[i]); break;
    case 4: case 5: {   // the surrogate plant's sine / cosine (include/rl4_citation_surrogate.h)
        double sn, cs;
        rl4_sincos(((const double*)a)[i], &sn, &cs);
        ((double*)out)[i] = (op == 4) ? sn : cs;
        break;
    }
    case 6: case 7: {   // ISA density / thrust lapse series of the default plant at altitude a[i]; b = rl4_cit_params on the device
        const rl4_cit_air air = rl4_cit_airdata((const rl4_cit_params*)b, ((const double*)a)[i]);
        ((double*)out)[i] = (op == 6) ? air.rho : air.thrust_lapse;
        break;
    }
    default: break;
    }
}

// exhaustive check of the re-spelled float quotient em/(em+2) against __fdiv_rn for every float
// bit pattern in [lo, hi): counts mismatches (fast path where its guard accepts, fallback otherwise)
__global__ void t13_div_f32_exhaustive_kernel(unsigned lo, unsigned hi, unsigned long long* mismatches)
{
    unsigned long long bad = 0;
    const unsigned long long n = (unsigned long long)hi - lo;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float em = __uint_as_float(lo + (unsigned)i);
        const float den = __fadd_rn(em, 2.0f);
        bool ok;
        float q = t13_div_fast(em, den, ok);
        if (!ok) q = __fdiv_rn(em, den);
        const float want = __fdiv_rn(em, den);
        if (__float_as_uint(q) != __float_as_uint(want)) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// exhaustive check of the seed of the double t13 quotient: rcp_rn_f32_normal(d) against __frcp_rn(d) (IEEE 1/d) for every
// float bit pattern in [lo, hi)
__global__ void rcp_f32_exhaustive_kernel(unsigned lo, unsigned hi, unsigned long long* mismatches)
{
    unsigned long long bad = 0;
    const unsigned long long n = (unsigned long long)hi - lo;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float d = __uint_as_float(lo + (unsigned)i);
        if (__float_as_uint(rcp_rn_f32_normal(d)) != __float_as_uint(__frcp_rn(d))) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

}  // namespace rl4

extern "C" {

int rl4_abi_version(void) { return RL4_ABI_VERSION; }

const char* rl4_last_error(void) { return rl4::g_err; }

int rl4_device_check(int device)
{
    cudaDeviceProp prop;
    RL4_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        rl4::set_error("rl4_device_check: device %d is sm_%d%d; this library is built for sm_100a only and has no fallback",
                       device, prop.major, prop.minor);
        return -2;
    }
    return 0;
}

int rl4_peak_fma(int is_double, double* out_flops_per_s, void* stream)
{
    RL4_REQUIRE(out_flops_per_s != nullptr, "out_flops_per_s is NULL");
    switch (is_double) {
    case 0: return rl4::run_peak<float>(out_flops_per_s, (cudaStream_t)stream);
    case 1: return rl4::run_peak<double>(out_flops_per_s, (cudaStream_t)stream);
    case 2: return rl4::run_peak<float, true>(out_flops_per_s, (cudaStream_t)stream);
    case 3: return rl4::run_peak<double, true>(out_flops_per_s, (cudaStream_t)stream);
    default: rl4::set_error("rl4_peak_fma: mode %d (0 float, 1 double, 2 / 3 the same with three distinct register operands)", is_double); return -1;
    }
}

int rl4_test_math(int op, const void* a, const void* b, void* out, int64_t n, void* stream)
{
    RL4_REQUIRE(a && out && n >= 0 && op >= 0 && op <= 7, "bad argument");
    RL4_REQUIRE(!((op == 2 || op == 3 || op == 6 || op == 7) && !b), "second operand is NULL");
    if (n == 0) return 0;
    rl4::math_probe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(op, a, b, out, n);
    return rl4::check_launch("math_probe_kernel");
}

int rl4_test_t13_div_f32(uint32_t lo_bits, uint32_t hi_bits, unsigned long long* device_mismatch_counter, void* stream)
{
    RL4_REQUIRE(device_mismatch_counter && hi_bits >= lo_bits, "bad argument");
    rl4::t13_div_f32_exhaustive_kernel<<<148 * 16, 256, 0, (cudaStream_t)stream>>>(lo_bits, hi_bits, device_mismatch_counter);
    return rl4::check_launch("t13_div_f32_exhaustive_kernel");
}

int rl4_test_rcp_f32(uint32_t lo_bits, uint32_t hi_bits, unsigned long long* device_mismatch_counter, void* stream)
{
    RL4_REQUIRE(device_mismatch_counter && hi_bits >= lo_bits, "bad argument");
    rl4::rcp_f32_exhaustive_kernel<<<148 * 16, 256, 0, (cudaStream_t)stream>>>(lo_bits, hi_bits, device_mismatch_counter);
    return rl4::check_launch("rcp_f32_exhaustive_kernel");
}

int64_t rl4_launch_count(void) { return rl4::g_launch_count.load(); }

}  // extern "C"
