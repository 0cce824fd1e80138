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

// Same, as a PROGRAMMATIC DEPENDENT LAUNCH: the kernel may be scheduled while its predecessor on the stream (or the
// graph branch it is captured on) is still running; it must call pb::pdl_wait() before touching anything the
// predecessor wrote.  Chains of short dependent kernels (the priority store's sample -> mark -> lines -> rebuild) save
// the launch latency at every boundary.  PB_NO_PDL=1 launches them the ordinary way.
// Two groups: the priority-store kernels (PB_LAUNCH_PDL; on by default) and the kernels of the agent's update chain
// (PB_LAUNCH_PDL_CHAIN; OFF by default: measured -- a dependent kernel that is resident early holds SM slots while it
// waits, and in the step graph those slots belong to the GEMMs of the parallel branches: 86.8 vs 83.8 us per step).
// PB_PDL=all | per | none selects; PB_NO_PDL=1 is "none".
bool pb_pdl_enabled();
bool pb_pdl_chain_enabled();
#define PB_LAUNCH_PDL_IF(enabled, kernel, grid, block, smem, stream, ...)                   \
    do {                                                                                    \
        cudaLaunchConfig_t pb_cfg_ = {};                                                    \
        pb_cfg_.gridDim = dim3(grid); pb_cfg_.blockDim = dim3(block);                       \
        pb_cfg_.dynamicSmemBytes = (smem); pb_cfg_.stream = (cudaStream_t)(stream);         \
        cudaLaunchAttribute pb_at_[1];                                                      \
        pb_at_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                  \
        pb_at_[0].val.programmaticStreamSerializationAllowed = 1;                           \
        pb_cfg_.attrs = pb_at_; pb_cfg_.numAttrs = (enabled) ? 1 : 0;                       \
        cudaError_t pb_e_ = cudaLaunchKernelEx(&pb_cfg_, kernel, __VA_ARGS__);              \
        g_pb_launches.fetch_add(1, std::memory_order_relaxed);                              \
        if (pb_e_ != cudaSuccess) return (int)pb_e_;                                        \
    } while (0)
#define PB_LAUNCH_PDL(kernel, grid, block, smem, stream, ...)                               \
    PB_LAUNCH_PDL_IF(pb_pdl_enabled(), kernel, grid, block, smem, stream, __VA_ARGS__)
#define PB_LAUNCH_PDL_CHAIN(kernel, grid, block, smem, stream, ...)                         \
    PB_LAUNCH_PDL_IF(pb_pdl_chain_enabled(), kernel, grid, block, smem, stream, __VA_ARGS__)

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

// programmatic dependent launch (PB_LAUNCH_PDL): wait until the preceding grid has completed and its writes are
// visible / allow the next grid to be scheduled.  Both are no-ops in a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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
