// Host-buffer episodes: what a reference user calls -- IDHPsp(...).train() / IDHPnonlin(...).train() for a batch -- with
// HOST input and output buffers (idhp_sp.py:176-181, idhp_nonlin.py:156-157).  The batch is cut into chunks of agents
// that flow through kStreams streams, so that the H2D copy of chunk c+1, the fused kernel of chunk c and the D2H copy of
// chunk c-1 overlap (agents are independent; a chunk is a column range of the SoA planes).  With pinned host buffers the
// copies are fully asynchronous.  Only the field groups selected by io->out_mask travel back.
#include "rl4_runtime.h"
#include "../../include/rl4afcs_b200.h"
#include <math_constants.h>

namespace rl4 {

// ---- N(0,1) stream of the nonlinear agent: Philox4x32-10 (Salmon et al. 2011), counter = (agent, step), key = seed
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1)
{
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

__global__ void __launch_bounds__(256)
nl_noise_kernel(uint64_t seed, int64_t agent0, int k0, int n_steps, int64_t n, float* __restrict__ out, int64_t stride)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (i >= n || r >= n_steps) return;
    const uint64_t g = (uint64_t)(agent0 + i);
    uint32_t c[4] = {(uint32_t)g, (uint32_t)(g >> 32), (uint32_t)(k0 + r), 0u};
    uint32_t ka = (uint32_t)seed, kb = (uint32_t)(seed >> 32);
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        philox_round(c, ka, kb);
        ka += 0x9E3779B9u; kb += 0xBB67AE85u;
    }
    const float u1 = ((float)(c[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);     // (0, 1): 24 random bits, never 0 or 1
    const float u2 = ((float)(c[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    out[(int64_t)r * stride + i] = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);   // Box-Muller
}

// contiguous row ranges of a plane selected by an output mask
struct RowRange { int lo, hi; };
static int ranges_from_rows(const bool* rows, int n_rows, RowRange* out)
{
    int n = 0;
    for (int r = 0; r < n_rows;) {
        if (!rows[r]) { ++r; continue; }
        int e = r;
        while (e < n_rows && rows[e]) ++e;
        out[n++] = {r, e};
        r = e;
    }
    return n;
}
static void mark(bool* rows, int lo, int hi) { for (int r = lo; r < hi; ++r) rows[r] = true; }

}  // namespace rl4

using namespace rl4;

struct rl4_ctx {
    int device;
    int policy;
    int64_t max_agents;
    int32_t max_steps;
    static constexpr int kStreams = 4;
    static constexpr int kNoiseSteps = 250;      // steps per nonlinear launch inside rl4_nl_episode_host
    cudaStream_t streams[kStreams];
    cudaEvent_t ref_ready;
    double* d_in;        // SP: [22][max_agents] x0(2) w1a(4) w2a(4) w1c(4) w2c(8);  NL: [120][max_agents] (allocated on first use)
    int64_t d_in_rows;
    double* d_ref;       // [max_steps]
    void* d_env;
    void* d_net;
    int32_t* d_ints;
    // nonlinear path (allocated by the first rl4_nl_episode_host call)
    double* nl_env;
    void* nl_net;
    int32_t* nl_ints;
    float* nl_noise;     // [kNoiseSteps][max_agents]
    size_t te, tn;
};

extern "C" {

int rl4_ctx_destroy(rl4_ctx* c);

int rl4_ctx_create(int device, int policy, int64_t max_agents, int32_t max_steps, rl4_ctx** out)
{
    RL4_REQUIRE(out != nullptr, "out is NULL");
    RL4_REQUIRE(max_agents > 0 && max_steps > 0, "bad capacity");
    RL4_REQUIRE(policy == RL4_FP64 || policy == RL4_FP32 || policy == RL4_MIXED, "unknown policy");
    int rc = rl4_device_check(device);
    if (rc) return rc;
    RL4_CUDA(cudaSetDevice(device));
    rl4_ctx* c = new rl4_ctx();                        // value-initialised: every handle / pointer starts as null
    c->device = device; c->policy = policy; c->max_agents = max_agents; c->max_steps = max_steps;
    c->te = (policy == RL4_FP32) ? 4 : 8;
    c->tn = (policy == RL4_FP64) ? 8 : 4;
    // a failing allocation must not leak the handles created before it: collect the first error, then destroy
    cudaError_t err = cudaSuccess;
    const char* what = "";
    auto step = [&](cudaError_t e, const char* w) { if (err == cudaSuccess && e != cudaSuccess) { err = e; what = w; } };
    for (int i = 0; i < rl4_ctx::kStreams; ++i) step(cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking), "cudaStreamCreateWithFlags");
    step(cudaEventCreateWithFlags(&c->ref_ready, cudaEventDisableTiming), "cudaEventCreateWithFlags");
    c->d_in_rows = 22;
    if (err == cudaSuccess) step(cudaMalloc(&c->d_in, sizeof(double) * c->d_in_rows * max_agents), "cudaMalloc(d_in)");
    if (err == cudaSuccess) step(cudaMalloc(&c->d_ref, sizeof(double) * max_steps), "cudaMalloc(d_ref)");
    if (err == cudaSuccess) step(cudaMalloc(&c->d_env, c->te * RL4_SPE_COUNT * max_agents), "cudaMalloc(d_env)");
    if (err == cudaSuccess) step(cudaMalloc(&c->d_net, c->tn * RL4_SPN_COUNT * max_agents), "cudaMalloc(d_net)");
    if (err == cudaSuccess) step(cudaMalloc(&c->d_ints, sizeof(int32_t) * RL4_SPI_COUNT * max_agents), "cudaMalloc(d_ints)");
    if (err != cudaSuccess) {
        rl4_ctx_destroy(c);
        return rl4::cuda_fail(err, what);
    }
    *out = c;
    return 0;
}

int rl4_ctx_destroy(rl4_ctx* c)
{
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaFree(c->d_in); cudaFree(c->d_ref); cudaFree(c->d_env); cudaFree(c->d_net); cudaFree(c->d_ints);
    cudaFree(c->nl_env); cudaFree(c->nl_net); cudaFree(c->nl_ints); cudaFree(c->nl_noise);
    for (int i = 0; i < rl4_ctx::kStreams; ++i) if (c->streams[i]) cudaStreamDestroy(c->streams[i]);
    if (c->ref_ready) cudaEventDestroy(c->ref_ready);
    delete c;
    return 0;
}

// every exit path of the episodes goes through here: copies already queued keep touching the caller's host buffers until
// the streams are idle, so a failing call must drain them before it reports the failure
static int drain(rl4_ctx* c, int rc)
{
    for (int i = 0; i < rl4_ctx::kStreams; ++i) {
        const cudaError_t e = cudaStreamSynchronize(c->streams[i]);
        if (e != cudaSuccess && rc == 0) rc = rl4::cuda_fail(e, "cudaStreamSynchronize");
    }
    return rc;
}

#define EP_CUDA(call)                                                              \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) { rc = rl4::cuda_fail(e__, #call); goto done; }    \
    } while (0)

static int chunking(int64_t n, int64_t unit, int64_t max_chunks, int64_t* per)
{
    int64_t n_chunks = n / unit;
    if (n_chunks < 1) n_chunks = 1;
    if (n_chunks > max_chunks) n_chunks = max_chunks;
    *per = ((n + n_chunks - 1) / n_chunks + 255) / 256 * 256;
    return (int)((n + *per - 1) / *per);
}

int rl4_sp_episode_host(rl4_ctx* c, const rl4_sp_params* p, const rl4_sp_host_io* io, int64_t n, int32_t n_steps,
                        int32_t use_traces)
{
    RL4_REQUIRE(c && p && io, "NULL argument");
    RL4_REQUIRE(n > 0 && n <= c->max_agents && n_steps > 0 && n_steps <= c->max_steps, "size exceeds the context capacity");
    RL4_REQUIRE(io->x0 && io->w1a && io->w2a && io->w1c && io->w2c && io->ref_base, "NULL input buffer");
    RL4_CUDA(cudaSetDevice(c->device));
    const int mask = io->out_mask ? io->out_mask : RL4_OUT_ALL;
    bool env_rows[RL4_SPE_COUNT] = {}, net_rows[RL4_SPN_COUNT] = {};
    if (mask & RL4_OUT_STATE) { mark(env_rows, RL4_SPE_X, RL4_SPE_THETA); mark(env_rows, RL4_SPE_CGRAD_PREV, RL4_SPE_EPS);
                                mark(net_rows, RL4_SPN_A, RL4_SPN_W1A); mark(net_rows, RL4_SPN_MPREV, RL4_SPN_COUNT); }
    if (mask & RL4_OUT_RLS) { mark(env_rows, RL4_SPE_THETA, RL4_SPE_CGRAD_PREV); mark(env_rows, RL4_SPE_EPS, RL4_SPE_SUM_C); }
    if (mask & RL4_OUT_STATS) mark(env_rows, RL4_SPE_SUM_C, RL4_SPE_EA);
    if (mask & RL4_OUT_TRACES) mark(env_rows, RL4_SPE_EA, RL4_SPE_COUNT);
    if (mask & RL4_OUT_WEIGHTS) mark(net_rows, RL4_SPN_W1A, RL4_SPN_W1T);
    if (mask & RL4_OUT_TARGET) mark(net_rows, RL4_SPN_W1T, RL4_SPN_MPREV);
    RowRange er[RL4_SPE_COUNT], nr[RL4_SPN_COUNT];
    const int n_er = io->out_env ? ranges_from_rows(env_rows, RL4_SPE_COUNT, er) : 0;
    const int n_nr = io->out_net ? ranges_from_rows(net_rows, RL4_SPN_COUNT, nr) : 0;
    const bool want_ints = io->out_ints && (mask & (RL4_OUT_STATS | RL4_OUT_STATE));

    const int64_t S = c->max_agents;
    int rc = 0;
    int64_t per = 0;
    chunking(n, 65536, 8, &per);
    const struct { const double* host; int rows; int64_t dev_row; } ins[5] = {
        {io->x0, 2, 0}, {io->w1a, 4, 2}, {io->w2a, 4, 6}, {io->w1c, 4, 10}, {io->w2c, 8, 14}};
    int ci = 0;
    EP_CUDA(cudaMemcpyAsync(c->d_ref, io->ref_base, sizeof(double) * n_steps, cudaMemcpyHostToDevice, c->streams[0]));
    EP_CUDA(cudaEventRecord(c->ref_ready, c->streams[0]));
    for (int64_t off = 0; off < n; off += per, ++ci) {
        const int64_t m = (n - off < per) ? (n - off) : per;
        cudaStream_t s = c->streams[ci % rl4_ctx::kStreams];
        if (ci % rl4_ctx::kStreams != 0 || ci >= rl4_ctx::kStreams) EP_CUDA(cudaStreamWaitEvent(s, c->ref_ready, 0));
        for (const auto& in : ins)   // host planes are [rows][n]; device planes are [rows][S]
            EP_CUDA(cudaMemcpy2DAsync(c->d_in + in.dev_row * S + off, S * 8, in.host + off, n * 8, m * 8, in.rows,
                                      cudaMemcpyHostToDevice, s));
        rl4_sp_params pc = *p;       // per-agent override arrays follow the chunk
        for (int j = 0; j < RL4_HP_COUNT; ++j) if (pc.hp_agent[j]) pc.hp_agent[j] += off;
        for (int j = 0; j < RL4_HPI_COUNT; ++j) if (pc.hpi_agent[j]) pc.hpi_agent[j] += off;
        rl4_sp_state st{(char*)c->d_env + off * c->te, (char*)c->d_net + off * c->tn, c->d_ints + off, S};
        rc = rl4_sp_init(c->policy, &pc, c->d_in + off, c->d_in + 2 * S + off, c->d_in + 6 * S + off, c->d_in + 10 * S + off,
                         c->d_in + 14 * S + off, S, st, m, s);
        if (rc) goto done;
        rl4_sp_log lg{nullptr, RL4_LOG_NONE, 1, 0};
        rc = rl4_sp_run(c->policy, &pc, c->d_ref, 0, n_steps, st, m, use_traces, lg, s);
        if (rc) goto done;
        for (int j = 0; j < n_er; ++j)
            EP_CUDA(cudaMemcpy2DAsync((char*)io->out_env + ((int64_t)er[j].lo * n + off) * c->te, n * c->te,
                                      (char*)st.env + (int64_t)er[j].lo * S * c->te, S * c->te, m * c->te, er[j].hi - er[j].lo,
                                      cudaMemcpyDeviceToHost, s));
        for (int j = 0; j < n_nr; ++j)
            EP_CUDA(cudaMemcpy2DAsync((char*)io->out_net + ((int64_t)nr[j].lo * n + off) * c->tn, n * c->tn,
                                      (char*)st.net + (int64_t)nr[j].lo * S * c->tn, S * c->tn, m * c->tn, nr[j].hi - nr[j].lo,
                                      cudaMemcpyDeviceToHost, s));
        if (want_ints)
            EP_CUDA(cudaMemcpy2DAsync(io->out_ints + off, n * 4, st.ints, S * 4, m * 4, RL4_SPI_COUNT, cudaMemcpyDeviceToHost, s));
    }
done:
    return drain(c, rc);
}

int rl4_nl_noise_fill(uint64_t seed, int64_t agent0, int32_t k0, int32_t n_steps, int64_t n, float* out, int64_t stride, void* stream)
{
    RL4_REQUIRE(out != nullptr, "out is NULL");
    RL4_REQUIRE(n >= 0 && stride >= n && k0 >= 0 && agent0 >= 0 && n_steps >= 0 && n_steps <= 65535, "bad size (n_steps <= 65535 per call)");
    if (n == 0 || n_steps == 0) return 0;
    const dim3 grid((unsigned)((n + 255) / 256), (unsigned)n_steps);
    nl_noise_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(seed, agent0, k0, n_steps, n, out, stride);
    return check_launch("nl_noise_kernel");
}

int rl4_nl_episode_host(rl4_ctx* c, const rl4_nl_params* p, const rl4_nl_host_io* io, int64_t n, int32_t n_steps)
{
    RL4_REQUIRE(c && p && io, "NULL argument");
    RL4_REQUIRE(c->policy == RL4_MIXED || c->policy == RL4_FP64, "the nonlinear path supports the mixed and fp64 policies");
    RL4_REQUIRE(n > 0 && n <= c->max_agents && n_steps > 0 && n_steps <= c->max_steps, "size exceeds the context capacity");
    RL4_REQUIRE(io->w1a && io->w2a && io->w1c && io->w2c && io->theta_ref, "NULL input buffer");
    RL4_CUDA(cudaSetDevice(c->device));
    const int64_t S = c->max_agents;
    if (!c->nl_env) {                                   // first nonlinear call on this context
        if (c->d_in_rows < 120) {
            double* bigger = nullptr;
            RL4_CUDA(cudaMalloc(&bigger, sizeof(double) * 120 * S));
            cudaFree(c->d_in);
            c->d_in = bigger; c->d_in_rows = 120;
        }
        RL4_CUDA(cudaMalloc(&c->nl_env, sizeof(double) * RL4_NLE_COUNT * S));
        RL4_CUDA(cudaMalloc(&c->nl_net, c->tn * RL4_NLN_COUNT * S));
        RL4_CUDA(cudaMalloc(&c->nl_ints, sizeof(int32_t) * RL4_NLI_COUNT * S));
        RL4_CUDA(cudaMalloc(&c->nl_noise, sizeof(float) * rl4_ctx::kNoiseSteps * S));
    }
    const int mask = io->out_mask ? io->out_mask : RL4_OUT_ALL;
    bool env_rows[RL4_NLE_COUNT] = {}, net_rows[RL4_NLN_COUNT] = {};
    if (mask & RL4_OUT_STATE) { mark(env_rows, RL4_NLE_XFULL, RL4_NLE_THETA); mark(env_rows, RL4_NLE_CGRAD_PREV, RL4_NLE_EPS);
                                mark(env_rows, RL4_NLE_ETA_A, RL4_NLE_EA);
                                mark(net_rows, RL4_NLN_S, RL4_NLN_W1A); mark(net_rows, RL4_NLN_MPREV, RL4_NLN_COUNT); }
    if (mask & RL4_OUT_RLS) { mark(env_rows, RL4_NLE_THETA, RL4_NLE_CGRAD_PREV); mark(env_rows, RL4_NLE_EPS, RL4_NLE_RSE); }
    if (mask & RL4_OUT_STATS) { mark(env_rows, RL4_NLE_RSE, RL4_NLE_EA); mark(env_rows, RL4_NLE_RSE_FLIGHT, RL4_NLE_COUNT); }
    if (mask & RL4_OUT_TRACES) mark(env_rows, RL4_NLE_EA, RL4_NLE_RSE_FLIGHT);
    if (mask & RL4_OUT_WEIGHTS) mark(net_rows, RL4_NLN_W1A, RL4_NLN_W1T);
    if (mask & RL4_OUT_TARGET) mark(net_rows, RL4_NLN_W1T, RL4_NLN_MPREV);
    RowRange er[RL4_NLE_COUNT], nr[RL4_NLN_COUNT];
    const int n_er = io->out_env ? ranges_from_rows(env_rows, RL4_NLE_COUNT, er) : 0;
    const int n_nr = io->out_net ? ranges_from_rows(net_rows, RL4_NLN_COUNT, nr) : 0;
    const bool want_ints = io->out_ints && (mask & (RL4_OUT_STATS | RL4_OUT_STATE));

    int rc = 0;
    int64_t per = 0;
    chunking(n, 32768, rl4_ctx::kStreams, &per);
    const struct { const double* host; int rows; int64_t dev_row; } ins[4] = {
        {io->w1a, 40, 0}, {io->w2a, 10, 40}, {io->w1c, 40, 50}, {io->w2c, 30, 90}};
    int ci = 0;
    EP_CUDA(cudaMemcpyAsync(c->d_ref, io->theta_ref, sizeof(double) * n_steps, cudaMemcpyHostToDevice, c->streams[0]));
    EP_CUDA(cudaEventRecord(c->ref_ready, c->streams[0]));
    for (int64_t off = 0; off < n; off += per, ++ci) {
        const int64_t m = (n - off < per) ? (n - off) : per;
        cudaStream_t s = c->streams[ci % rl4_ctx::kStreams];
        if (ci % rl4_ctx::kStreams != 0) EP_CUDA(cudaStreamWaitEvent(s, c->ref_ready, 0));
        for (const auto& in : ins)
            EP_CUDA(cudaMemcpy2DAsync(c->d_in + in.dev_row * S + off, S * 8, in.host + off, n * 8, m * 8, in.rows,
                                      cudaMemcpyHostToDevice, s));
        rl4_nl_params pc = *p;
        for (int j = 0; j < RL4_NHP_COUNT; ++j) if (pc.hp_agent[j]) pc.hp_agent[j] += off;
        for (int j = 0; j < RL4_NHPI_COUNT; ++j) if (pc.hpi_agent[j]) pc.hpi_agent[j] += off;
        rl4_nl_state st{c->nl_env + off, (char*)c->nl_net + off * c->tn, c->nl_ints + off, S};
        rc = rl4_nl_init(c->policy, &pc, c->d_in + off, c->d_in + 40 * S + off, c->d_in + 50 * S + off, c->d_in + 90 * S + off,
                         S, st, m, s);
        if (rc) goto done;
        rl4_sp_log lg{nullptr, RL4_LOG_NONE, 1, 0};
        for (int k = 0; k < n_steps; k += rl4_ctx::kNoiseSteps) {
            const int ks = (n_steps - k < rl4_ctx::kNoiseSteps) ? (n_steps - k) : rl4_ctx::kNoiseSteps;
            float* nz = c->nl_noise + off;              // this chunk's columns of the [kNoiseSteps][S] buffer
            if (io->noise)
                EP_CUDA(cudaMemcpy2DAsync(nz, S * 4, io->noise + (int64_t)k * n + off, n * 4, m * 4, ks, cudaMemcpyHostToDevice, s));
            else {
                rc = rl4_nl_noise_fill(io->noise_seed, io->noise_agent0 + off, k, ks, m, nz, S, s);
                if (rc) goto done;
            }
            rc = rl4_nl_run(c->policy, &pc, c->d_ref, nz, S, k, ks, st, m, lg, s);
            if (rc) goto done;
        }
        for (int j = 0; j < n_er; ++j)
            EP_CUDA(cudaMemcpy2DAsync(io->out_env + (int64_t)er[j].lo * n + off, n * 8, st.env + (int64_t)er[j].lo * S, S * 8, m * 8,
                                      er[j].hi - er[j].lo, cudaMemcpyDeviceToHost, s));
        for (int j = 0; j < n_nr; ++j)
            EP_CUDA(cudaMemcpy2DAsync((char*)io->out_net + ((int64_t)nr[j].lo * n + off) * c->tn, n * c->tn,
                                      (char*)st.net + (int64_t)nr[j].lo * S * c->tn, S * c->tn, m * c->tn, nr[j].hi - nr[j].lo,
                                      cudaMemcpyDeviceToHost, s));
        if (want_ints)
            EP_CUDA(cudaMemcpy2DAsync(io->out_ints + off, n * 4, st.ints, S * 4, m * 4, RL4_NLI_COUNT, cudaMemcpyDeviceToHost, s));
    }
done:
    return drain(c, rc);
}

}  // extern "C"
