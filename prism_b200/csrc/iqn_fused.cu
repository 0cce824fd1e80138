// iqn_fused.cu -- element-wise halves of the IQN trunk that sit between the tensor-core GEMMs (sm_100a).
//
// Reference: prism/agents/models/iqn_model.py:70-71 builds h = phi(tau) (.) tile(x) for quantile-major rows
// r = q*B + b and autograd then runs ~8 ATen kernels over the (n*B, F) tensors in the backward pass (two
// broadcast multiplies, ReLU mask, two reductions).  Here the forward product is fused into the phi GEMM's
// epilogue (pb_tc_gemm `mul`), and the whole backward of  h = relu(pre) (.) x  is one pass:
//     dpre[q,b,:] = dh[q,b,:] * x[b,:] * [phi[q,b,:] > 0]      (feeds the weight-gradient GEMM)
//     dx[b,:]     = sum_q dh[q,b,:] * phi[q,b,:]
//     dbp[b,:]    = sum_q dpre[q,b,:]                           (bias gradient = column sum of dbp over b)
// One thread owns a (b, 4 columns) strip and walks q: every load/store is a coalesced 128-bit access and the
// q-reductions stay in registers (no atomics: deterministic).  HBM-bound: 12 B read + 4 B written per element.
#include "common.cuh"

namespace {

using namespace pb;

__global__ void __launch_bounds__(256) iqn_phi_bwd_kernel(int n, int B, int F4, const float4 *__restrict__ dh,
                                                          const float4 *__restrict__ phi, const float4 *__restrict__ x,
                                                          float4 *__restrict__ dpre, float4 *__restrict__ dx,
                                                          float4 *__restrict__ dbp)
{
    const long long per = (long long)B * F4;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per) return;
    const float4 xv = x[i];
    float4 ax = make_float4(0.f, 0.f, 0.f, 0.f), ab = ax;
#pragma unroll 4
    for (int q = 0; q < n; ++q) {
        const uint4 du = ldg_stream(reinterpret_cast<const uint4 *>(dh + q * per + i));
        const uint4 pu = ldg_stream(reinterpret_cast<const uint4 *>(phi + q * per + i));
        const float4 d = make_float4(__uint_as_float(du.x), __uint_as_float(du.y), __uint_as_float(du.z), __uint_as_float(du.w));
        const float4 p = make_float4(__uint_as_float(pu.x), __uint_as_float(pu.y), __uint_as_float(pu.z), __uint_as_float(pu.w));
        float4 o;
        o.x = p.x > 0.f ? d.x * xv.x : 0.f;
        o.y = p.y > 0.f ? d.y * xv.y : 0.f;
        o.z = p.z > 0.f ? d.z * xv.z : 0.f;
        o.w = p.w > 0.f ? d.w * xv.w : 0.f;
        ax.x += d.x * p.x; ax.y += d.y * p.y; ax.z += d.z * p.z; ax.w += d.w * p.w;
        ab.x += o.x; ab.y += o.y; ab.z += o.z; ab.w += o.w;
        stg_stream(reinterpret_cast<uint4 *>(dpre + q * per + i),
                   make_uint4(__float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), __float_as_uint(o.w)));
    }
    if (dx) dx[i] = ax;
    dbp[i] = ab;
}

// ReLU backward + bias gradient of a dense layer in one pass: dz = dy * [y > 0] (or dy itself), and per-strip column
// sums of dz -- replaces ATen's compare + multiply + sum(dim) (three passes over an (M x N) tensor per layer).
// Block = one strip of `strip` rows; thread t owns the 4-column groups t, t + 256, ...; coalesced 128-bit accesses.
__global__ void __launch_bounds__(256) relu_bwd_colsum_kernel(long long rows, int N4, int strip, const float4 *__restrict__ dy,
                                                              const float4 *__restrict__ y, float4 *__restrict__ dz,
                                                              float4 *__restrict__ partial)
{
    const long long r0 = (long long)blockIdx.x * strip, r1 = min(rows, r0 + strip);
    const size_t head = (size_t)blockIdx.y * rows * N4;                 // heads are stacked matrices
    for (int c = threadIdx.x; c < N4; c += blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (long long r = r0; r < r1; ++r) {
            const size_t o = head + (size_t)r * N4 + c;
            float4 d = dy[o];
            if (y) {
                const float4 m = y[o];
                d.x = m.x > 0.f ? d.x : 0.f; d.y = m.y > 0.f ? d.y : 0.f;
                d.z = m.z > 0.f ? d.z : 0.f; d.w = m.w > 0.f ? d.w : 0.f;
                dz[o] = d;
            }
            acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
        }
        partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * N4 + c] = acc;
    }
}

// out[h][c] = sum over strips (fixed order); 32 columns x 32 strip groups per block
__global__ void __launch_bounds__(1024) colsum_reduce_kernel(int nstrips, int N, const float *__restrict__ partial,
                                                             float *__restrict__ out)
{
    __shared__ float sm[32][33];
    const int c = blockIdx.x * 32 + threadIdx.x, r = threadIdx.y;
    const float *p = partial + (size_t)blockIdx.y * nstrips * N;
    float a = 0.f;
    if (c < N)
        for (int k = r; k < nstrips; k += 32) a += p[(size_t)k * N + c];
    sm[r][threadIdx.x] = a;
    __syncthreads();
    if (r == 0 && c < N) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) t += sm[k][threadIdx.x];
        out[(size_t)blockIdx.y * N + c] = t;
    }
}

}  // namespace

extern "C" {

int pb_relu_bwd_bias_strips(long long rows) { long long s = (rows + 63) / 64; return (int)(s < 1 ? 1 : s); }

// heads stacked (heads, rows, N) matrices.  y NULL: no activation (dz is not written; dz may be NULL).
// partials: heads * pb_relu_bwd_bias_strips(rows) * N floats.  dbias: (heads, N), may be NULL.
int pb_relu_bwd_bias(int heads, long long rows, int N, const float *dy, const float *y, float *dz, float *dbias,
                     float *partials, void *stream)
{
    if (heads <= 0 || rows <= 0 || N <= 0 || (N % 4) != 0 || !dy || !partials || (y && !dz)) return PB_E_ARG;
    if ((((uintptr_t)dy) | ((uintptr_t)y) | ((uintptr_t)dz) | ((uintptr_t)partials)) & 15) return PB_E_ARG;
    const int strips = pb_relu_bwd_bias_strips(rows);
    dim3 grid((unsigned)strips, (unsigned)heads);
    PB_LAUNCH(relu_bwd_colsum_kernel, grid, 256, 0, stream, rows, N / 4, 64, reinterpret_cast<const float4 *>(dy),
              reinterpret_cast<const float4 *>(y), reinterpret_cast<float4 *>(dz), reinterpret_cast<float4 *>(partials));
    if (dbias) {
        dim3 rgrid((unsigned)((N + 31) / 32), (unsigned)heads);
        PB_LAUNCH(colsum_reduce_kernel, rgrid, dim3(32, 32), 0, stream, strips, N, partials, dbias);
    }
    return PB_OK;
}

int pb_iqn_phi_bwd(int n, int B, int F, const float *dh, const float *phi, const float *x, float *dpre, float *dx,
                   float *dbias_partial, void *stream)
{
    if (n <= 0 || B <= 0 || F <= 0 || (F % 4) != 0 || !dh || !phi || !x || !dpre || !dbias_partial) return PB_E_ARG;
    if ((((uintptr_t)dh) | ((uintptr_t)phi) | ((uintptr_t)x) | ((uintptr_t)dpre) | ((uintptr_t)dx) | ((uintptr_t)dbias_partial)) & 15)
        return PB_E_ARG;
    const long long per = (long long)B * (F / 4);
    const long long blocks = (per + 255) / 256;
    PB_LAUNCH(iqn_phi_bwd_kernel, (unsigned)blocks, 256, 0, stream, n, B, F / 4, reinterpret_cast<const float4 *>(dh),
              reinterpret_cast<const float4 *>(phi), reinterpret_cast<const float4 *>(x), reinterpret_cast<float4 *>(dpre),
              reinterpret_cast<float4 *>(dx), reinterpret_cast<float4 *>(dbias_partial));
    return PB_OK;
}

}  // extern "C"
