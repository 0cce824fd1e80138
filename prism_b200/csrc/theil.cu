// theil.cu -- the Q-ensemble's head-diversity regulariser as three launches (sm_100a).
//
// Reference (prism/agents/models/q_ensemble.py:86-92): the Theil index of the K heads' parameter L2 norms,
//     l_k = || theta_k ||_2,   r_k = l_k / mean_j l_j,   T = mean_k r_k log r_k,
// built there from parameters_to_vector + ~10 ATen ops per head and differentiated by autograd -- about a hundred
// tiny launches per update for K = 10.  Here the heads' parameters are stacked tensors (K, ...), so
//   1. theil_sumsq_kernel : per (tensor, head, chunk) partial sums of squares (multi-tensor table, 128-bit loads),
//   2. theil_finalize_kernel: l, r, T and the per-head gradient factor  c_k = (log r_k - T) / (K * mean(l) * l_k)
//      (d T / d theta_k = c_k * theta_k),
//   3. theil_bwd_kernel    : grad_k = upstream * c_k * theta_k for every stacked tensor in one launch.
#include "common.cuh"

namespace {

using namespace pb;

constexpr int TH_MAX_CHUNKS = 128;   // blocks per (tensor, head): chosen by the caller from the largest per-head size

// table[t] = {pointer to the stacked tensor (K, per_head), per_head element count, offset of its gradient in `out`}
__global__ void __launch_bounds__(256) theil_sumsq_kernel(const long long *__restrict__ table, float *__restrict__ partial)
{
    __shared__ float red[8];
    const int t = blockIdx.z, k = blockIdx.y, c = blockIdx.x, K = gridDim.y, CH = gridDim.x;
    const long long per = table[3 * t + 1];
    const float *p = reinterpret_cast<const float *>(table[3 * t + 0]) + (long long)k * per;
    float acc = 0.f;
    if ((per & 3) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        const float4 *p4 = reinterpret_cast<const float4 *>(p);
        const long long n4 = per >> 2, stride = (long long)CH * blockDim.x;
        long long i = (long long)c * blockDim.x + threadIdx.x;
        for (; i + 3 * stride < n4; i += 4 * stride) {
            float4 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = p4[i + q * stride];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc += (v[q].x * v[q].x + v[q].y * v[q].y) + (v[q].z * v[q].z + v[q].w * v[q].w);
        }
        for (; i < n4; i += stride) {
            const float4 v = p4[i];
            acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        }
    } else {
        for (long long i = (long long)c * blockDim.x + threadIdx.x; i < per; i += (long long)CH * blockDim.x) {
            const float v = p[i];
            acc = fmaf(v, v, acc);
        }
    }
    acc = warp_sum(acc);
    if (lane_id() == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[w];
        partial[((size_t)t * K + k) * CH + c] = s;
    }
}

__global__ void __launch_bounds__(64) theil_finalize_kernel(int n_tensors, int K, int CH, const float *__restrict__ partial,
                                                            float *__restrict__ theil_out, float *__restrict__ coef)
{
    __shared__ float l[64];
    const int k = threadIdx.x;
    float sq = 0.f;
    if (k < K)
        for (int t = 0; t < n_tensors; ++t)
            for (int c = 0; c < CH; ++c) sq += partial[((size_t)t * K + k) * CH + c];
    l[k] = k < K ? sqrtf(sq) : 0.f;
    __syncthreads();
    float m = 0.f;
    for (int j = 0; j < K; ++j) m += l[j];
    m /= (float)K;
    float T = 0.f;
    for (int j = 0; j < K; ++j) { const float r = l[j] / m; T += r * logf(r); }
    T /= (float)K;
    if (k == 0) *theil_out = T;
    if (k < K) coef[k] = (logf(l[k] / m) - T) / ((float)K * m * l[k]);
}

__global__ void __launch_bounds__(256) theil_bwd_kernel(const long long *__restrict__ table, const float *__restrict__ coef,
                                                        const float *__restrict__ upstream, float *__restrict__ out)
{
    const int t = blockIdx.z, k = blockIdx.y, c = blockIdx.x, CH = gridDim.x;
    const long long per = table[3 * t + 1];
    const float *p = reinterpret_cast<const float *>(table[3 * t + 0]) + (long long)k * per;
    float *o = out + table[3 * t + 2] + (long long)k * per;
    const float s = coef[k] * (*upstream);
    if ((per & 3) == 0 && ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(o)) & 15) == 0) {
        const float4 *p4 = reinterpret_cast<const float4 *>(p);
        float4 *o4 = reinterpret_cast<float4 *>(o);
        const long long n4 = per >> 2, stride = (long long)CH * blockDim.x;
        long long i = (long long)c * blockDim.x + threadIdx.x;
        for (; i + 3 * stride < n4; i += 4 * stride) {             // 4 independent 128-bit loads in flight
            float4 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = p4[i + q * stride];
#pragma unroll
            for (int q = 0; q < 4; ++q) o4[i + q * stride] = make_float4(s * v[q].x, s * v[q].y, s * v[q].z, s * v[q].w);
        }
        for (; i < n4; i += stride) {
            const float4 v = p4[i];
            o4[i] = make_float4(s * v.x, s * v.y, s * v.z, s * v.w);
        }
    } else {
        for (long long i = (long long)c * blockDim.x + threadIdx.x; i < per; i += (long long)CH * blockDim.x) o[i] = s * p[i];
    }
}

}  // namespace

extern "C" {

// blocks per (tensor, head) for tensors of up to max_per_head elements per head
int pb_theil_chunks(long long max_per_head)
{
    long long c = (max_per_head + 4095) / 4096;      // 4 float4 per thread per CTA of 256
    if (c < 8) c = 8;
    if (c > TH_MAX_CHUNKS) c = TH_MAX_CHUNKS;
    return (int)c;
}

// partial: n_tensors * K * chunks floats of scratch; coef: K floats (kept for the backward call); K <= 64
int pb_theil_fwd(int n_tensors, int K, int chunks, const long long *table, float *partial, float *theil_out, float *coef,
                 void *stream)
{
    if (n_tensors <= 0 || K <= 0 || K > 64 || chunks <= 0 || chunks > TH_MAX_CHUNKS || !table || !partial || !theil_out || !coef)
        return PB_E_ARG;
    dim3 grid((unsigned)chunks, (unsigned)K, (unsigned)n_tensors);
    PB_LAUNCH(theil_sumsq_kernel, grid, 256, 0, stream, table, partial);
    PB_LAUNCH(theil_finalize_kernel, 1, 64, 0, stream, n_tensors, K, chunks, partial, theil_out, coef);
    return PB_OK;
}

// out[table[t].offset + k * per + i] = *upstream * coef[k] * theta_t[k][i]
int pb_theil_bwd(int n_tensors, int K, int chunks, const long long *table, const float *coef, const float *upstream,
                 float *out, void *stream)
{
    if (n_tensors <= 0 || K <= 0 || K > 64 || chunks <= 0 || chunks > TH_MAX_CHUNKS || !table || !coef || !upstream || !out)
        return PB_E_ARG;
    dim3 grid((unsigned)chunks, (unsigned)K, (unsigned)n_tensors);
    PB_LAUNCH(theil_bwd_kernel, grid, 256, 0, stream, table, coef, upstream, out);
    return PB_OK;
}

}  // extern "C"
