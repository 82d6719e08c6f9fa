// Host-side helpers shared by the translation units of librl4afcs_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>

namespace rl4 {

// thread-local message returned by rl4_last_error()
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);   // records the message, returns (int)e
extern std::atomic<int64_t> g_launch_count;

inline int check_launch(const char* what)
{
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, what);
    return 0;
}

#define RL4_REQUIRE(cond, msg)                          \
    do {                                                \
        if (!(cond)) { rl4::set_error("%s: %s", __func__, msg); return -1; } \
    } while (0)

#define RL4_CUDA(call)                                  \
    do {                                                \
        cudaError_t e__ = (call);                       \
        if (e__ != cudaSuccess) return rl4::cuda_fail(e__, #call); \
    } while (0)

}  // namespace rl4
