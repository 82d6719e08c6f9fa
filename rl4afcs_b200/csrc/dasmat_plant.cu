// The reference's OWN nonlinear aircraft on the GPU (row a25: `_citation.step`, envs/nonlinear/citation.py:62-69, called at
// envs/nonlinear/env.py:210,288-291).
//
// The reference ships the DASMAT Citation model only as x86-64 machine code inside a Windows .pyd.  tools/lift_plant.py
// translates that machine code, instruction by instruction, into C over an explicit machine state (include/rl4_lift_runtime.h); the
// build (rl4afcs_b200/build.py, only where /root/reference exists) writes the translation to csrc/_gen/ and this unit
// compiles it for sm_100a: one aircraft per thread, every thread executing the model's own instruction stream on its own
// copy of the model's writable memory.  What the binary computes, this computes -- the only arithmetic that is not the
// binary's own is inside the eight C-runtime functions it imports (sin cos tan exp log10 pow floor sqrt: CUDA's here,
// the Windows UCRT's in the reference, glibc's in the in-process run the golden fixtures come from; 1-2 ulp apart).
//
// Memory model of one aircraft (emulated virtual addresses, LIFT_BASE = image base):
//   image   [0x00000, 0x40000)  the DLL's sections after initialize() has run -- ONE copy in HBM, read-only during steps
//   A       [0x2eb00, 0x2ec00)  \  the only parts of .data that step() writes (model time, block signals, continuous and
//   D       [0x3a000, 0x3c200)  /  discrete states, solver work arrays): PER AIRCRAFT, 1120 eight-byte words
//   stack   [0x40000, 0x42000)  per thread, scratch
// A and D persist between launches in an SoA plane `state[word * stride + aircraft]` (coalesced copy in / out); inside a
// launch they and the stack live in thread-local memory, which the hardware interleaves per lane: one window
// [0x3a000, 0x42000) (D, the never-touched tail of the image, the stack -- a single range check decodes a dynamic address;
// untouched lines cost no cache) plus A.  Memory operands are addressed with the low 32 bits of the emulated address.
// A store outside A / D / stack, an indirect call to an unknown target or an untranslated instruction sets an error bit
// that the host reads back: the translation never silently computes something else.
#include "rl4_runtime.h"
#include "../../include/rl4afcs_b200.h"
#include <math_constants.h>
#include <vector>

#if __has_include("_gen/dasmat_code_step.inc")
#define RL4_HAVE_DASMAT 1
#else
#define RL4_HAVE_DASMAT 0
#endif

namespace rl4 {

constexpr uint32_t kImg = 0x40000, kStack = 0x2000, kFlat = kImg + kStack;
constexpr uint32_t kALo = 0x2eb00, kASz = 0x100, kDLo = 0x3a000, kDSz = 0x2200;
constexpr uint32_t kWLo = kDLo, kWSz = kFlat - kDLo;            // thread-private window: D, the unused tail of the image, the stack
constexpr int kStateWords = (int)((kASz + kDSz) / 8);           // 1120
constexpr uint32_t kLocalBytes = kWSz + kASz;                    // window + A; only the touched lines of it ever occupy cache
constexpr uint32_t kRvaX = 0x3c120, kRvaEngine = 0x3c198;       // the 16 continuous states (established by running the binary)
constexpr uint32_t kBase32 = 0x80000000u;                       // low 32 bits of the image base: memory operands arrive as low halves

#ifndef RL4_DASMAT_THREADS
#define RL4_DASMAT_THREADS 1024
#endif
#ifndef RL4_DASMAT_MIN_BLOCKS
#define RL4_DASMAT_MIN_BLOCKS 1
#endif

#if RL4_HAVE_DASMAT

#define LIFT_HD __device__ __forceinline__
#define LIFT_CPU_EXTRA uint8_t* G; uint8_t* m; int err; int sync;
#define F_ADD(a, b) __dadd_rn((a), (b))
#define F_SUB(a, b) __dsub_rn((a), (b))
#define F_MUL(a, b) __dmul_rn((a), (b))
#define F_DIV(a, b) __ddiv_rn((a), (b))
#define F_SQRT(a) __dsqrt_rn(a)
#define U2D(u) __longlong_as_double((long long)(u))
#define D2U(d) ((uint64_t)__double_as_longlong(d))
#include "../../include/rl4_lift_runtime.h"

enum { kErrTrap = 1, kErrWildAccess = 2, kErrStoreToImage = 4 };

#define LIFT_LF_OFF(o) min((o), 0x7f8u)
#define LIFT_FN static __device__ __noinline__
#define LIFT_FN_INLINE static __device__ __forceinline__
#define LIFT_TRAP(msg, v) do { c->err |= kErrTrap; LIFT_TRAP_RETURN; } while (0)
// The axis-transformation S-function takes cos AND sin of the same five angles (ten library calls per block, ~1 200 per plant
// step, a quarter of all instructions): both go through sincos(), so that the two inlined expansions for one angle are the same
// computation and the compiler keeps one of them.
#ifndef RL4_DASMAT_SINCOS
#define RL4_DASMAT_SINCOS 1
#endif
#if RL4_DASMAT_SINCOS
__device__ __forceinline__ double lift_sin(double x) { double sn, cs; sincos(x, &sn, &cs); return sn; }
__device__ __forceinline__ double lift_cos(double x) { double sn, cs; sincos(x, &sn, &cs); return cs; }
#else
#define lift_cos cos
#define lift_sin sin
#endif
#define lift_tan tan
#define lift_exp exp
#define lift_floor floor
#define lift_log10 log10
#define lift_sqrt __dsqrt_rn
#define lift_pow pow
__device__ __forceinline__ uint64_t lift_CVTR32(double v) { return (v > -2147483649.0 && v < 2147483648.0) ? (uint64_t)(uint32_t)__double2int_rn(v) : 0x80000000ULL; }
__device__ __forceinline__ uint64_t lift_CVTR64(double v) { return (uint64_t)__double2ll_rn(v); }
__device__ __forceinline__ uint64_t lift_malloc(cpu_t* c, uint64_t) { c->err |= kErrTrap; return 0; }   // the model allocates nothing

// byte-wise helpers over whichever LD8 / ST8 the enclosing namespace defines
#define LIFT_DEFINE_BULK                                                                                           \
    __device__ __noinline__ void lift_memcpy(cpu_t* c, uint64_t d64, uint64_t s64, uint64_t n)                      \
    {                                                                                                               \
        LIFT_MEM_CTX const uint32_t d = (uint32_t)d64, s = (uint32_t)s64;                                           \
        if (d <= s) for (uint32_t k = 0; k < n; ++k) ST8(d + k, LD8(s + k));                                        \
        else for (uint32_t k = (uint32_t)n; k-- > 0;) ST8(d + k, LD8(s + k));                                       \
    }                                                                                                               \
    __device__ __noinline__ void lift_memset(cpu_t* c, uint64_t d64, int v, uint64_t n)                             \
    {                                                                                                               \
        LIFT_MEM_CTX const uint32_t d = (uint32_t)d64;                                                              \
        for (uint32_t k = 0; k < n; ++k) ST8(d + k, v);                                                             \
    }                                                                                                               \
    __device__ __noinline__ void lift_REPSTOS(cpu_t* c, uint64_t* rcx, uint64_t* rdi, uint64_t rax, unsigned w)    \
    {                                                                                                               \
        LIFT_MEM_CTX                                                                                                \
        for (; *rcx; --*rcx, *rdi += w)                                                                             \
            for (unsigned k = 0; k < w; ++k) ST8((uint32_t)*rdi + k, rax >> (8 * k));                               \
    }                                                                                                               \
    __device__ __noinline__ void lift_REPMOVS(cpu_t* c, uint64_t* rcx, uint64_t* rdi, uint64_t* rsi, unsigned w)   \
    {                                                                                                               \
        LIFT_MEM_CTX                                                                                                \
        for (; *rcx; --*rcx, *rdi += w, *rsi += w)                                                                  \
            for (unsigned k = 0; k < w; ++k) ST8((uint32_t)*rdi + k, LD8((uint32_t)*rsi + k));                      \
    }

// ---- initialize(): flat model, everything (image + stack) in one global buffer ------------------------------------------
namespace init_mode {
#define LIFT_MEM_CTX uint8_t* const G_ = c->G; (void)G_;
template <typename T> __device__ __forceinline__ T* at(uint8_t* G_, cpu_t* c, uint32_t a32)
{
    uint32_t off = a32 - kBase32;
    if (off >= kFlat) { c->err |= kErrWildAccess; off = kImg; }
    return reinterpret_cast<T*>(G_ + off);
}
#define LD8(a)  ((uint64_t)*at<uint8_t>(G_, c, (a)))
#define LD16(a) ((uint64_t)*at<uint16_t>(G_, c, (a)))
#define LD32(a) ((uint64_t)*at<uint32_t>(G_, c, (a)))
#define LD64(a) (*at<uint64_t>(G_, c, (a)))
#define LDD(a)  (*at<double>(G_, c, (a)))
#define ST8(a, v)  (*at<uint8_t>(G_, c, (a)) = (uint8_t)(v))
#define ST16(a, v) (*at<uint16_t>(G_, c, (a)) = (uint16_t)(v))
#define ST32(a, v) (*at<uint32_t>(G_, c, (a)) = (uint32_t)(v))
#define ST64(a, v) (*at<uint64_t>(G_, c, (a)) = (uint64_t)(v))
#define LDS8 LD8
#define LDS16 LD16
#define LDS32 LD32
#define LDS64 LD64
#define LDSD LDD
#define STS8 ST8
#define STS16 ST16
#define STS32 ST32
#define STS64 ST64
#define LDW8 LD8
#define LDW16 LD16
#define LDW32 LD32
#define LDW64 LD64
#define LDWD LDD
#define STW8 ST8
#define STW16 ST16
#define STW32 ST32
#define STW64 ST64
#define LDI8 LD8
#define LDI16 LD16
#define LDI32 LD32
#define LDI64 LD64
#define LDID LDD
LIFT_DEFINE_BULK
#include "_gen/dasmat_code_init.inc"
#undef LIFT_MEM_CTX
#undef LD8
#undef LD16
#undef LD32
#undef LD64
#undef LDD
#undef ST8
#undef ST16
#undef ST32
#undef ST64
#undef LDS8
#undef LDS16
#undef LDS32
#undef LDS64
#undef LDSD
#undef STS8
#undef STS16
#undef STS32
#undef STS64
#undef LDW8
#undef LDW16
#undef LDW32
#undef LDW64
#undef LDWD
#undef STW8
#undef STW16
#undef STW32
#undef STW64
#undef LDI8
#undef LDI16
#undef LDI32
#undef LDI64
#undef LDID
#undef LIFT_LOCALS
#undef LIFT_ENTER
#undef LIFT_PRECALL
#undef LIFT_POSTCALL
#undef LIFT_EXIT
}  // namespace init_mode

// ---- step(): thread-private window (D, stack) + A in local memory, the image shared and read-only ------------------------
namespace step_mode {
// m_ is the thread's local array: telling the compiler so turns the accesses into LDL / STL with immediate offsets
#define LIFT_MEM_CTX uint8_t* const m_ = c->m; const uint8_t* const G_ = c->G; __builtin_assume(__isLocal(m_)); (void)m_; (void)G_;
// keeps the warps of a CTA on the same stretch of the code (far larger than the instruction cache); c->sync is uniform over
// the CTA by construction (set by the kernels from a CTA-wide vote)
#define LIFT_SYNC if (c->sync) __syncthreads()
// the same inside step(), in front of calls that every execution passes exactly once (tools/lift_plant.py: spine_calls)
#ifndef RL4_DASMAT_SYNC_SFUN
#define RL4_DASMAT_SYNC_SFUN 0   // measured: 1.33e7 (entry only) -> 1.22e7 (+ S-function calls) -> 1.17e7 (+ helper calls) plant steps/s
#endif
#ifndef RL4_DASMAT_SYNC_HELPER
#define RL4_DASMAT_SYNC_HELPER 0
#endif
#if RL4_DASMAT_SYNC_SFUN
#define LIFT_SYNC_SFUN if (c->sync) __syncthreads()
#endif
#if RL4_DASMAT_SYNC_HELPER
#define LIFT_SYNC_HELPER if (c->sync) __syncthreads()
#endif
// Run-time decode of an address the translator could not place (a pointer through the Simulink SimStruct): BRANCH-FREE --
// the pointer is selected, then ONE generic load / store follows.  (A branchy decode made ptxas duplicate the code behind
// every access per region: the 59-instruction table-interpolation helper at RVA 0xe8d0 became 5 268 SASS instructions.)
#ifndef RL4_DASMAT_BRANCHFREE
#define RL4_DASMAT_BRANCHFREE 1
#endif
template <typename T> __device__ __forceinline__ T lift_load(uint8_t* m_, const uint8_t* G_, cpu_t* c, uint32_t a32)
{
    const uint32_t off = a32 - kBase32;
#if RL4_DASMAT_BRANCHFREE
    (void)c;
    const uint32_t ow = off - kWLo, oa = off - kALo;
    const uint8_t* p = G_ + (off < kImg ? off : 0u);                  // a wild address reads the image header: garbage in, flagged on stores
    p = oa < kASz ? m_ + kWSz + oa : p;
    p = ow < kWSz ? m_ + ow : p;
    return *reinterpret_cast<const T*>(p);
#else
    if (off - kWLo < kWSz) return *reinterpret_cast<const T*>(m_ + (off - kWLo));
    if (off - kALo < kASz) return *reinterpret_cast<const T*>(m_ + kWSz + (off - kALo));
    if (off >= kImg) { c->err |= kErrWildAccess; return T(0); }
    return __ldg(reinterpret_cast<const T*>(G_ + off));              // shared image: read-only while aircraft are stepping
#endif
}
template <typename T> __device__ __forceinline__ void lift_store(uint8_t* m_, cpu_t* c, uint32_t a32, T v)
{
    const uint32_t off = a32 - kBase32;
#if RL4_DASMAT_BRANCHFREE
    const uint32_t ow = off - kWLo, oa = off - kALo;
    // gaps of the window that step() never writes: the translator folded loads from them (lift_plant.py FROZEN_GAPS)
    const bool gap = off - 0x3a000u < 0x78u || off - 0x3a4b8u < 0x88u || off - 0x3a574u < 0x74u || off - 0x3a5f0u < 0x590u;
    const bool bad = (!(ow < kWSz) && !(oa < kASz)) || gap;
    uint8_t* p = m_ + (oa < kASz ? kWSz + oa : (ow < kWSz ? ow : kDSz));     // a store outside A / D / stack lands in the window's unused gap
    *reinterpret_cast<T*>(p) = v;
    if (bad) c->err |= kErrStoreToImage;
#else
    if (off - kWLo < kWSz) { *reinterpret_cast<T*>(m_ + (off - kWLo)) = v; return; }
    if (off - kALo < kASz) { *reinterpret_cast<T*>(m_ + kWSz + (off - kALo)) = v; return; }
    c->err |= kErrStoreToImage;
#endif
}
#define LD8(a)  ((uint64_t)lift_load<uint8_t>(m_, G_, c, (a)))
#define LD16(a) ((uint64_t)lift_load<uint16_t>(m_, G_, c, (a)))
#define LD32(a) ((uint64_t)lift_load<uint32_t>(m_, G_, c, (a)))
#define LD64(a) lift_load<uint64_t>(m_, G_, c, (a))
#define LDD(a)  lift_load<double>(m_, G_, c, (a))
#define ST8(a, v)  lift_store<uint8_t>(m_, c, (a), (uint8_t)(v))
#define ST16(a, v) lift_store<uint16_t>(m_, c, (a), (uint16_t)(v))
#define ST32(a, v) lift_store<uint32_t>(m_, c, (a), (uint32_t)(v))
#define ST64(a, v) lift_store<uint64_t>(m_, c, (a), (uint64_t)(v))
// operands the translator knows to be on the stack: inside the window by construction, no decoding
// (offsets are clamped: a NaN that reaches an index computation of a diverging aircraft must not become an out-of-bounds access;
// for the constant addresses -- nearly all of them -- the compiler folds the clamp away)
#define LIFT_STK(T, a) reinterpret_cast<T*>(m_ + min((uint32_t)((a) - (kBase32 + kWLo)), kWSz - 8u))
#define LDS8(a)  ((uint64_t)*LIFT_STK(uint8_t, a))
#define LDS16(a) ((uint64_t)*LIFT_STK(uint16_t, a))
#define LDS32(a) ((uint64_t)*LIFT_STK(uint32_t, a))
#define LDS64(a) (*LIFT_STK(uint64_t, a))
#define LDSD(a)  (*LIFT_STK(double, a))
#define STS8(a, v)  (*LIFT_STK(uint8_t, a) = (uint8_t)(v))
#define STS16(a, v) (*LIFT_STK(uint16_t, a) = (uint16_t)(v))
#define STS32(a, v) (*LIFT_STK(uint32_t, a) = (uint32_t)(v))
#define STS64(a, v) (*LIFT_STK(uint64_t, a) = (uint64_t)(v))
// operands whose object the translator knows (include/rl4_lift_runtime.h, tools/lift_plant.py "region hints"; the CPU build
// of the same translation checks every hint at run time): W = inside the writable window, I = read-only image
#define LDW8(a)  ((uint64_t)*LIFT_STK(uint8_t, a))
#define LDW16(a) ((uint64_t)*LIFT_STK(uint16_t, a))
#define LDW32(a) ((uint64_t)*LIFT_STK(uint32_t, a))
#define LDW64(a) (*LIFT_STK(uint64_t, a))
#define LDWD(a)  (*LIFT_STK(double, a))
#define STW8(a, v)  (*LIFT_STK(uint8_t, a) = (uint8_t)(v))
#define STW16(a, v) (*LIFT_STK(uint16_t, a) = (uint16_t)(v))
#define STW32(a, v) (*LIFT_STK(uint32_t, a) = (uint32_t)(v))
#define STW64(a, v) (*LIFT_STK(uint64_t, a) = (uint64_t)(v))
#define LIFT_IMG(T, a) __ldg(reinterpret_cast<const T*>(G_ + min((uint32_t)((a) - kBase32), kImg - 8u)))
#define LDI8(a)  ((uint64_t)LIFT_IMG(uint8_t, a))
#define LDI16(a) ((uint64_t)LIFT_IMG(uint16_t, a))
#define LDI32(a) ((uint64_t)LIFT_IMG(uint32_t, a))
#define LDI64(a) LIFT_IMG(uint64_t, a)
#define LDID(a)  LIFT_IMG(double, a)
LIFT_DEFINE_BULK
#include "_gen/dasmat_code_step.inc"
}  // namespace step_mode

// ---- one aircraft inside the fused env + agent kernel (nl_core.cuh, PLANT = 1) ------------------------------------------------
struct DasmatThread {
    cpu_t c;
    __align__(16) uint8_t m[kLocalBytes];
};
}  // namespace rl4
#define RL4_NL_WITH_DASMAT 1
namespace rl4 {
struct DasmatIo;
__device__ __forceinline__ void dasmat_thread_load(DasmatThread* t, const DasmatIo& io, int64_t i);
__device__ __forceinline__ void dasmat_thread_store(DasmatThread* t, const DasmatIo& io, int64_t i);
__device__ __forceinline__ void dasmat_thread_set_sync(DasmatThread* t, bool on) { t->c.sync = on ? 1 : 0; }
__device__ __forceinline__ void enter(cpu_t& c)
{
    // Windows x64 frame at the top of the stack: return address slot at rsp, 32 bytes of home space above it
    c.r[4] = LIFT_BASE + kFlat - 0x100 - 8;
}
constexpr uint64_t kArgIn = LIFT_BASE + kFlat - 0x100 + 0x20, kArgOut = kArgIn + 96;   // the caller's buffers, above the frame
// citation.step(u) -> xo (envs/nonlinear/citation.py:62-69)
__device__ __forceinline__ void dasmat_thread_step(DasmatThread* t, const double (&u)[11], double (&xo)[12])
{
    double* in_p = reinterpret_cast<double*>(t->m + ((uint32_t)kArgIn - kBase32 - kWLo));
    const double* out_p = reinterpret_cast<const double*>(t->m + ((uint32_t)kArgOut - kBase32 - kWLo));
#pragma unroll
    for (int j = 0; j < 11; ++j) in_p[j] = u[j];
    enter(t->c);
    t->c.r[1] = kArgOut; t->c.r[2] = kArgIn;
    LIFT_INVOKE(step_mode::f_180003720, &t->c);
#pragma unroll
    for (int j = 0; j < 12; ++j) xo[j] = out_p[j];
}
}  // namespace rl4
#include "nl_core.cuh"
namespace rl4 {
__device__ __forceinline__ void dasmat_thread_load(DasmatThread* t, const DasmatIo& io, int64_t i)
{
    uint64_t* wd = reinterpret_cast<uint64_t*>(t->m);
    uint64_t* wa = reinterpret_cast<uint64_t*>(t->m + kWSz);
    for (int w = 0; w < (int)(kASz / 8); ++w) wa[w] = io.state[(int64_t)w * io.stride + i];
    for (int w = 0; w < (int)(kDSz / 8); ++w) wd[w] = io.state[(int64_t)(kASz / 8 + w) * io.stride + i];
    memset(&t->c, 0, sizeof t->c);
    t->c.G = io.image; t->c.m = t->m; t->c.sync = 1;       // every thread of the CTA takes every plant step (nl_core.cuh)
}
__device__ __forceinline__ void dasmat_thread_store(DasmatThread* t, const DasmatIo& io, int64_t i)
{
    const uint64_t* wd = reinterpret_cast<const uint64_t*>(t->m);
    const uint64_t* wa = reinterpret_cast<const uint64_t*>(t->m + kWSz);
    for (int w = 0; w < (int)(kASz / 8); ++w) io.state[(int64_t)w * io.stride + i] = wa[w];
    for (int w = 0; w < (int)(kDSz / 8); ++w) io.state[(int64_t)(kASz / 8 + w) * io.stride + i] = wd[w];
    if (t->c.err) atomicOr(io.err, t->c.err);
}

// fused episode launches with the translated plant: float32 networks (the reference's own mix), general hyper-parameters
template <bool LOG>
static int dasmat_launch_fused(const rl4_nl_params* p, const double* theta_ref, const float* noise, int64_t noise_stride, int k0,
                               int n_steps, rl4_nl_state st, int64_t n, rl4_sp_log lg, DasmatIo dio, cudaStream_t s)
{
    constexpr int BLK = kNlBlockDasmat;
    auto kern = nl_run_kernel<float, RL4_CIT_INTEGRATOR_ODE5, LOG, true, 1>;      // no shared memory: see nl_core.cuh
    kern<<<(unsigned)((n + BLK - 1) / BLK), BLK, 0, s>>>(*p, theta_ref, noise, noise_stride, k0, n_steps, st, n, lg, dio);
    return check_launch("nl_run_kernel<dasmat>");
}

// Ce500NonLinear.step with the translated plant, step-API form (envs/nonlinear/env.py:182-256); same outputs as nl_env_step_kernel
__global__ void __launch_bounds__(RL4_DASMAT_THREADS, RL4_DASMAT_MIN_BLOCKS)
dasmat_env_step_kernel(const __grid_constant__ rl4_nl_params p, const double* __restrict__ theta_ref, int stepp, double* __restrict__ x_full,
                       double* __restrict__ x_act_p, const double* __restrict__ action, double* __restrict__ out_mdp,
                       double* __restrict__ out_reward, double* __restrict__ out_e, double* __restrict__ out_surf,
                       double* __restrict__ out_eff, double* __restrict__ out_x_obs, int64_t S, int64_t n_agents, const DasmatIo dio)
{
    const int64_t i_raw = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i_raw < n_agents;
    const int64_t i = active ? i_raw : n_agents - 1;
    const NlHp<true> hv{p, i};
    DasmatThread dz;
    dasmat_thread_load(&dz, dio, i);
    dz.c.sync = 1;                                  // every thread of the CTA takes the step (tail threads shadow the last aircraft)
    double x[12], xo[12], xa[3], act[3], surf[3], ueff[3], e_phi, e_th, e_psi, reward, rg2;
    for (int j = 0; j < 12; ++j) x[j] = x_full[j * S + i];
    for (int j = 0; j < 3; ++j) { xa[j] = x_act_p[j * S + i]; act[j] = action[j * S + i]; }
    nl_env_step<true, RL4_CIT_INTEGRATOR_ODE5, 1>(p, hv, stepp, __ldg(theta_ref + stepp), act, x, xa, surf, e_phi, e_th, e_psi, reward, rg2, ueff, xo, &dz);
    if (!active) return;
    dasmat_thread_store(&dz, dio, i);
    for (int j = 0; j < 12; ++j) x_full[j * S + i] = x[j];
    for (int j = 0; j < 3; ++j) x_act_p[j * S + i] = xa[j];
    if (out_x_obs) for (int j = 0; j < 12; ++j) out_x_obs[j * S + i] = xo[j];
    out_mdp[i] = xo[4]; out_mdp[S + i] = xo[7]; out_mdp[2 * S + i] = xo[1]; out_mdp[3 * S + i] = e_th;
    out_reward[i] = reward; out_e[i] = e_th;
    if (out_surf) for (int j = 0; j < 3; ++j) out_surf[j * S + i] = surf[j];
    if (out_eff) for (int j = 0; j < 3; ++j) out_eff[j * S + i] = ueff[j];
}

// the pristine image (sections at their RVAs), embedded at build time
static const uint8_t kPristine[] = {
#include "_gen/dasmat_image.inc"
};

// initialize() in flat mode: one thread, everything (image + stack) in the global buffer G
__global__ void dasmat_initialize_kernel(uint8_t* G, int* err)
{
    if (blockIdx.x || threadIdx.x) return;
    cpu_t c;
    memset(&c, 0, sizeof c);
    c.G = G; c.m = nullptr;
    enter(c);
    LIFT_INVOKE(init_mode::f_1800096f0, &c);
    *err = c.err;
}

// per-aircraft state words <- the image's A and D regions
__global__ void dasmat_reset_kernel(const uint64_t* __restrict__ G, uint64_t* __restrict__ state, int64_t stride, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int w = 0; w < (int)(kASz / 8); ++w) state[(int64_t)w * stride + i] = G[kALo / 8 + w];
    for (int w = 0; w < (int)(kDSz / 8); ++w) state[(int64_t)(kASz / 8 + w) * stride + i] = G[kDLo / 8 + w];
}

// n_steps calls of step(u) per aircraft.  u: [11][u_stride] (held for all steps of the launch); out: what the LAST call
// returned, [12][out_stride]; out_all (optional): every call's return, [n_steps][12][out_stride].
// Every thread of the CTA runs the loop (the translated code synchronises the CTA at each step() entry): aircraft beyond
// n_agents are clamped copies of the last one and store nothing.
__global__ void __launch_bounds__(RL4_DASMAT_THREADS, RL4_DASMAT_MIN_BLOCKS)
dasmat_step_kernel(uint8_t* G, uint64_t* __restrict__ state, int64_t stride, int64_t n, const double* __restrict__ u, int64_t u_stride,
                   int n_steps, double* __restrict__ out, int64_t out_stride, double* __restrict__ out_all, int* __restrict__ err)
{
    const int64_t i_raw = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i_raw < n;
    const int64_t i = active ? i_raw : n - 1;
    __align__(16) uint8_t m[kLocalBytes];
    uint64_t* wd = reinterpret_cast<uint64_t*>(m);                       // D at the bottom of the window
    uint64_t* wa = reinterpret_cast<uint64_t*>(m + kWSz);                // A behind the window
    for (int w = 0; w < (int)(kASz / 8); ++w) wa[w] = state[(int64_t)w * stride + i];
    for (int w = 0; w < (int)(kDSz / 8); ++w) wd[w] = state[(int64_t)(kASz / 8 + w) * stride + i];
    cpu_t c;
    memset(&c, 0, sizeof c);
    c.G = G; c.m = m; c.sync = 1;
    // the caller's buffers sit above the frame, inside the stack region (as a C caller's locals would)
    const uint64_t a_in = LIFT_BASE + kFlat - 0x100 + 0x20, a_out = a_in + 96;
    double* in_p = reinterpret_cast<double*>(m + ((uint32_t)a_in - kBase32 - kWLo));
    double* out_p = reinterpret_cast<double*>(m + ((uint32_t)a_out - kBase32 - kWLo));
    for (int k = 0; k < n_steps; ++k) {
        for (int j = 0; j < 11; ++j) in_p[j] = u[(int64_t)j * u_stride + i];
        enter(c);
        c.r[1] = a_out; c.r[2] = a_in;
        LIFT_INVOKE(step_mode::f_180003720, &c);
        if (out_all && active)
            for (int j = 0; j < 12; ++j) out_all[((int64_t)k * 12 + j) * out_stride + i] = out_p[j];
    }
    if (!active) return;
    if (out)
        for (int j = 0; j < 12; ++j) out[(int64_t)j * out_stride + i] = out_p[j];
    for (int w = 0; w < (int)(kASz / 8); ++w) state[(int64_t)w * stride + i] = wa[w];
    for (int w = 0; w < (int)(kDSz / 8); ++w) state[(int64_t)(kASz / 8 + w) * stride + i] = wd[w];
    if (c.err) atomicOr(err, c.err);
}

// aircraft `src_index` of one state plane copied to every aircraft of another (the trimmed state after reset)
__global__ void dasmat_broadcast_kernel(const uint64_t* __restrict__ src, int64_t src_stride, int64_t src_index, uint64_t* __restrict__ dst,
                                        int64_t dst_stride, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int w = 0; w < kStateWords; ++w) dst[(int64_t)w * dst_stride + i] = src[(int64_t)w * src_stride + src_index];
}

#endif  // RL4_HAVE_DASMAT

static int no_plant()
{
    set_error("this library was built without the reference's plant binary (rl4afcs_b200/csrc/_gen/ is produced by "
              "rl4afcs_b200/tools/lift_plant.py where /root/reference exists): the 'dasmat' plant is not available");
    return -2;
}

}  // namespace rl4

using namespace rl4;

extern "C" {

int rl4_dasmat_available(void) { return RL4_HAVE_DASMAT; }
int64_t rl4_dasmat_image_bytes(void) { return (int64_t)kFlat; }
int32_t rl4_dasmat_state_words(void) { return kStateWords; }
/* word index (in the per-aircraft state) of the 12 airframe states / the 4 engine states */
int32_t rl4_dasmat_word_x(void) { return (int32_t)((kASz + (kRvaX - kDLo)) / 8); }
int32_t rl4_dasmat_word_engine(void) { return (int32_t)((kASz + (kRvaEngine - kDLo)) / 8); }

int rl4_dasmat_initialize(void* image, void* stream)
{
#if RL4_HAVE_DASMAT
    RL4_REQUIRE(image != nullptr, "null image buffer");
    cudaStream_t s = (cudaStream_t)stream;
    static_assert(sizeof(kPristine) <= kImg, "image larger than the emulated address space");
    RL4_CUDA(cudaMemsetAsync(image, 0, kFlat, s));
    RL4_CUDA(cudaMemcpyAsync(image, kPristine, sizeof(kPristine), cudaMemcpyHostToDevice, s));
    int* d_err = nullptr;
    RL4_CUDA(cudaMalloc(&d_err, sizeof(int)));
    RL4_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), s));
    dasmat_initialize_kernel<<<1, 1, 0, s>>>((uint8_t*)image, d_err);
    int rc = check_launch("dasmat_initialize_kernel");
    int h_err = 0;
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = cuda_fail(e, "dasmat initialize");
    }
    cudaFree(d_err);
    if (rc) return rc;
    if (h_err) { set_error("rl4_dasmat_initialize: the translated initialize() raised error bits 0x%x", h_err); return -3; }
    return 0;
#else
    (void)image; (void)stream;
    return no_plant();
#endif
}

int rl4_dasmat_reset(const void* image, uint64_t* state, int64_t stride, int64_t n_agents, void* stream)
{
#if RL4_HAVE_DASMAT
    RL4_REQUIRE(image && state && stride >= n_agents && n_agents > 0, "bad arguments");
    dasmat_reset_kernel<<<(unsigned)((n_agents + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint64_t*)image, state, stride, n_agents);
    return check_launch("dasmat_reset_kernel");
#else
    (void)image; (void)state; (void)stride; (void)n_agents; (void)stream;
    return no_plant();
#endif
}

int rl4_dasmat_step(void* image, uint64_t* state, int64_t stride, int64_t n_agents, const double* u, int64_t u_stride,
                    int32_t n_steps, double* out, int64_t out_stride, double* out_all, int32_t* device_err, void* stream)
{
#if RL4_HAVE_DASMAT
    RL4_REQUIRE(image && state && u && device_err && stride >= n_agents && u_stride >= n_agents && n_agents > 0 && n_steps > 0, "bad arguments");
    RL4_REQUIRE((!out && !out_all) || out_stride >= n_agents, "out_stride < n_agents");
    dasmat_step_kernel<<<(unsigned)((n_agents + RL4_DASMAT_THREADS - 1) / RL4_DASMAT_THREADS), RL4_DASMAT_THREADS, 0, (cudaStream_t)stream>>>(
        (uint8_t*)image, state, stride, n_agents, u, u_stride, n_steps, out, out_stride, out_all, device_err);
    return check_launch("dasmat_step_kernel");
#else
    (void)image; (void)state; (void)stride; (void)n_agents; (void)u; (void)u_stride; (void)n_steps; (void)out; (void)out_stride;
    (void)out_all; (void)device_err; (void)stream;
    return no_plant();
#endif
}

int rl4_dasmat_broadcast(const uint64_t* src, int64_t src_stride, int64_t src_index, uint64_t* dst, int64_t dst_stride, int64_t n_agents,
                         void* stream)
{
#if RL4_HAVE_DASMAT
    RL4_REQUIRE(src && dst && src_index >= 0 && src_stride > src_index && dst_stride >= n_agents && n_agents > 0, "bad arguments");
    dasmat_broadcast_kernel<<<(unsigned)((n_agents + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, src_stride, src_index, dst, dst_stride, n_agents);
    return check_launch("dasmat_broadcast_kernel");
#else
    (void)src; (void)src_stride; (void)src_index; (void)dst; (void)dst_stride; (void)n_agents; (void)stream;
    return no_plant();
#endif
}

int rl4_nl_env_step_dasmat(const rl4_nl_params* p, const double* theta_ref, int32_t stepp, double* x_full, double* x_act,
                           const double* action, double* out_mdp, double* out_reward, double* out_e_theta, double* out_surf,
                           double* out_eff, double* out_x_obs, int64_t stride, int64_t n, void* image, uint64_t* plant_state,
                           int64_t plant_stride, int32_t* device_err, void* stream)
{
#if RL4_HAVE_DASMAT
    RL4_REQUIRE(p && theta_ref && x_full && x_act && action && out_mdp && out_reward && out_e_theta && image && plant_state && device_err, "NULL argument");
    RL4_REQUIRE(n >= 0 && stride >= n && plant_stride >= n && stepp >= 0, "bad size");
    if (n == 0) return 0;
    const DasmatIo dio{(uint8_t*)image, plant_state, plant_stride, device_err};
    dasmat_env_step_kernel<<<(unsigned)((n + RL4_DASMAT_THREADS - 1) / RL4_DASMAT_THREADS), RL4_DASMAT_THREADS, 0, (cudaStream_t)stream>>>(
        *p, theta_ref, stepp, x_full, x_act, action, out_mdp, out_reward, out_e_theta, out_surf, out_eff, out_x_obs, stride, n, dio);
    return check_launch("dasmat_env_step_kernel");
#else
    (void)p; (void)theta_ref; (void)stepp; (void)x_full; (void)x_act; (void)action; (void)out_mdp; (void)out_reward; (void)out_e_theta;
    (void)out_surf; (void)out_eff; (void)out_x_obs; (void)stride; (void)n; (void)image; (void)plant_state; (void)plant_stride;
    (void)device_err; (void)stream;
    return no_plant();
#endif
}

int rl4_nl_run_dasmat(int policy, const rl4_nl_params* p, const double* theta_ref, const float* noise, int64_t noise_stride,
                      int32_t k0, int32_t n_steps, rl4_nl_state st, int64_t n, rl4_sp_log lg, void* image, uint64_t* plant_state,
                      int64_t plant_stride, int32_t* device_err, void* stream)
{
#if RL4_HAVE_DASMAT
    RL4_REQUIRE(p && theta_ref && noise && st.env && st.net && st.ints && image && plant_state && device_err, "NULL argument");
    RL4_REQUIRE(n >= 0 && st.stride >= n && noise_stride >= n && plant_stride >= n && k0 >= 0 && n_steps >= 0, "bad size");
    RL4_REQUIRE(policy == RL4_MIXED, "the dasmat plant runs with float32 networks (policy mixed, the reference's own arithmetic)");
    if (lg.level != RL4_LOG_NONE) RL4_REQUIRE(lg.level >= 1 && lg.level <= 3 && lg.buf && lg.every >= 1 && lg.n_agents_logged >= 0 && lg.n_agents_logged <= n, "bad log descriptor");
    if (n == 0 || n_steps == 0) return 0;
    const DasmatIo dio{(uint8_t*)image, plant_state, plant_stride, device_err};
    cudaStream_t s = (cudaStream_t)stream;
    return lg.level != RL4_LOG_NONE ? dasmat_launch_fused<true>(p, theta_ref, noise, noise_stride, k0, n_steps, st, n, lg, dio, s)
                                    : dasmat_launch_fused<false>(p, theta_ref, noise, noise_stride, k0, n_steps, st, n, lg, dio, s);
#else
    (void)policy; (void)p; (void)theta_ref; (void)noise; (void)noise_stride; (void)k0; (void)n_steps; (void)st; (void)n; (void)lg;
    (void)image; (void)plant_state; (void)plant_stride; (void)device_err; (void)stream;
    return no_plant();
#endif
}

}  // extern "C"
