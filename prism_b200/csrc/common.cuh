// common.cuh -- shared helpers for libprism_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include "prism_b200.h"

extern std::atomic<long long> g_pb_launches;

#define PB_CHECK_ARG(cond) do { if (!(cond)) return PB_E_ARG; } while (0)

// Launch + count + surface launch errors as positive cudaError_t codes.
#define PB_LAUNCH(kernel, grid, block, smem, stream, ...)                                   \
    do {                                                                                    \
        kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);           \
        g_pb_launches.fetch_add(1, std::memory_order_relaxed);                              \
        cudaError_t pb_e_ = cudaGetLastError();                                             \
        if (pb_e_ != cudaSuccess) return (int)pb_e_;                                        \
    } while (0)

static inline int pb_ilog2(long long v) { int l = 0; while ((1LL << l) < v) ++l; return l; }
static inline bool pb_is_pow2(long long v) { return v > 0 && (v & (v - 1)) == 0; }

// SM count of the current device (148 on B200); cached per device.
int pb_sm_count();

// One-time per-device setup (cudaFuncSetAttribute is a per-device property): a bitmask over device ordinals.
struct PbPerDeviceOnce {
    std::atomic<unsigned long long> mask{0};
    static int device() { int d = 0; return cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < 64 ? d : 0; }
    bool done() const { return (mask.load(std::memory_order_acquire) >> device()) & 1ull; }
    void mark() { mask.fetch_or(1ull << device(), std::memory_order_release); }
};

namespace pb {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// streaming 128-bit accesses that do not pollute L1
__device__ __forceinline__ uint4 ldg_stream(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(uint4 *p, const uint4 &v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

}  // namespace pb
