// ln.cu -- LayerNorm forward / backward for the (T*B)-row IQN activations (sm_100a), HBM-bound streaming kernels.
//
// The reference normalises before every Linear of its heads (nn.LayerNorm inside nn.Sequential,
// prism/agents/models/ffnn_model.py:17-18, iqn_model.py:42-46, q_ensemble.py:26-40).  At configs[4] that is a
// (32768 x 3136) fp32 tensor per pass: ATen's backward needs two kernels that each re-read x and dy
// (layer_norm_grad_input 221 us + GammaBetaBackward 297 us on a B200, profiles/launches_atari_r01_step.txt).  Here
//   forward : one pass, row kept in registers (two-pass mean / variance: no cancellation), saves mean and rstd;
//   backward: ONE pass over (x, dy) producing dx and per-CTA partial sums of dgamma / dbeta held in registers by
//             column-owning threads (persistent CTAs striding over rows), + a small column reduction.
// Rows up to 1024 floats are handled by one warp (shuffle reductions only), up to 4096 by one 256-thread CTA.
// Algorithmic traffic: forward 8 B / element, backward 12 B / element.
#include "common.cuh"

namespace {

using namespace pb;

constexpr int LN_THREADS = 256;

template <int TPR>
__device__ __forceinline__ float row_sum(float v, float *scratch)
{
    v = warp_sum(v);
    if (TPR == 32) return v;
    // whole CTA = one row
    const int w = threadIdx.x >> 5;
    __syncthreads();                                   // scratch reuse across successive reductions
    if (lane_id() == 0) scratch[w] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < LN_THREADS / 32; ++k) t += scratch[k];
    return t;
}

// two sums in one pass (one pair of barriers instead of two)
template <int TPR>
__device__ __forceinline__ float2 row_sum2(float a, float b, float2 *scratch2)
{
    a = warp_sum(a);
    b = warp_sum(b);
    if (TPR == 32) return make_float2(a, b);
    const int w = threadIdx.x >> 5;
    __syncthreads();
    if (lane_id() == 0) scratch2[w] = make_float2(a, b);
    __syncthreads();
    float2 t = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < LN_THREADS / 32; ++k) { t.x += scratch2[k].x; t.y += scratch2[k].y; }
    return t;
}

// TPR threads per row (32: one warp, 8 rows per CTA; 256: one CTA), SLOTS float4 per thread
template <int TPR, int SLOTS>
__global__ void __launch_bounds__(LN_THREADS) ln_fwd_kernel(long long rows, int F4, float eps, const float4 *__restrict__ x,
                                                            long long x_rows, const float4 *__restrict__ gamma,
                                                            const float4 *__restrict__ beta, long long group_rows,
                                                            float4 *__restrict__ y, float *__restrict__ mean_out,
                                                            float *__restrict__ rstd_out)
{
    // output row r normalises input row r % x_rows (K heads sharing one input) with the affine pair of group
    // r / group_rows (stacked (K, F) parameters of the ensemble heads); plain LayerNorm: x_rows = group_rows = rows
    __shared__ float scratch[LN_THREADS / 32];
    constexpr int RPC = LN_THREADS / TPR;
    const int lr = threadIdx.x % TPR, grp = threadIdx.x / TPR;
    const float inv_f = 1.0f / (float)(F4 * 4);
    for (long long row = (long long)blockIdx.x * RPC + grp; row < rows; row += (long long)gridDim.x * RPC) {
        const float4 *xr = x + (row % x_rows) * F4;
        const long long gofs = (row / group_rows) * F4;
        float4 v[SLOTS];
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < SLOTS; ++k) {
            const int c = lr + k * TPR;
            v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < F4) {
                const uint4 u = ldg_stream(reinterpret_cast<const uint4 *>(xr + c));
                v[k] = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
            }
            s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
        }
        const float mean = row_sum<TPR>(s, scratch) * inv_f;
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < SLOTS; ++k) {
            if (lr + k * TPR < F4) {
                const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
                q += (a * a + b * b) + (c * c + d * d);
            }
        }
        const float var = row_sum<TPR>(q, scratch) * inv_f;
        const float rstd = rsqrtf(var + eps);
#pragma unroll
        for (int k = 0; k < SLOTS; ++k) {
            const int c = lr + k * TPR;
            if (c < F4) {
                const float4 g = gamma ? __ldg(gamma + gofs + c) : make_float4(1.f, 1.f, 1.f, 1.f);
                const float4 b = beta ? __ldg(beta + gofs + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 o;
                o.x = (v[k].x - mean) * rstd * g.x + b.x;
                o.y = (v[k].y - mean) * rstd * g.y + b.y;
                o.z = (v[k].z - mean) * rstd * g.z + b.z;
                o.w = (v[k].w - mean) * rstd * g.w + b.w;
                stg_stream(reinterpret_cast<uint4 *>(y + row * F4 + c),
                           make_uint4(__float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), __float_as_uint(o.w)));
            }
        }
        if (lr == 0) {
            if (mean_out) mean_out[row] = mean;
            if (rstd_out) rstd_out[row] = rstd;
        }
    }
}

// dx = rstd * (g - mean_j(g) - xhat * mean_j(g * xhat)),  g = dy * gamma,  xhat = (x - mean) * rstd
// part_g[blockIdx][F], part_b[blockIdx][F]: this CTA's sums over its rows of dy * xhat and dy
template <int TPR, int SLOTS>
__global__ void __launch_bounds__(LN_THREADS, (TPR == 256 ? 3 : 1)) ln_bwd_kernel(long long rows, int F4, const float4 *__restrict__ x,
                                                            long long x_rows,
                                                            const float4 *__restrict__ dy, const float4 *__restrict__ gamma,
                                                            const float *__restrict__ mean_in, const float *__restrict__ rstd_in,
                                                            float4 *__restrict__ dx, float4 *__restrict__ part_g,
                                                            float4 *__restrict__ part_b)
{
    __shared__ float2 scratch2[LN_THREADS / 32];
    extern __shared__ float4 colsum[];                       // TPR == 32: [2][F4] CTA-level combine of the 8 warps
    constexpr int RPC = LN_THREADS / TPR;
    const int lr = threadIdx.x % TPR, grp = threadIdx.x / TPR;
    const float inv_f = 1.0f / (float)(F4 * 4);
    float4 ag[SLOTS], ab[SLOTS];
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) ag[k] = ab[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    // grid.y = affine groups (rows = rows PER GROUP; group g covers output rows [g * rows, (g + 1) * rows)); the
    // partial sums of a CTA belong to one group.  Output row r reads input row r % x_rows.
    const long long row0 = (long long)blockIdx.y * rows;
    const size_t pblock = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    if (gamma) gamma += (size_t)blockIdx.y * F4;
    // every thread of a CTA must take the same number of trips when the row reduction uses __syncthreads
    const long long trips = (rows + (long long)gridDim.x * RPC - 1) / ((long long)gridDim.x * RPC);
    for (long long it = 0; it < trips; ++it) {
        const long long lrow = (it * gridDim.x + blockIdx.x) * RPC + grp;
        const bool live = lrow < rows;
        const long long row = row0 + lrow;
        const long long xrow = row % x_rows;
        const float mean = live ? mean_in[row] : 0.f, rstd = live ? rstd_in[row] : 0.f;
        float4 xh[SLOTS], g[SLOTS];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < SLOTS; ++k) {
            const int c = lr + k * TPR;
            xh[k] = g[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live && c < F4) {
                const uint4 xu = ldg_stream(reinterpret_cast<const uint4 *>(x + xrow * F4 + c));
                const uint4 du = ldg_stream(reinterpret_cast<const uint4 *>(dy + row * F4 + c));
                const float4 d = make_float4(__uint_as_float(du.x), __uint_as_float(du.y), __uint_as_float(du.z), __uint_as_float(du.w));
                xh[k].x = (__uint_as_float(xu.x) - mean) * rstd; xh[k].y = (__uint_as_float(xu.y) - mean) * rstd;
                xh[k].z = (__uint_as_float(xu.z) - mean) * rstd; xh[k].w = (__uint_as_float(xu.w) - mean) * rstd;
                ag[k].x += d.x * xh[k].x; ag[k].y += d.y * xh[k].y; ag[k].z += d.z * xh[k].z; ag[k].w += d.w * xh[k].w;
                ab[k].x += d.x; ab[k].y += d.y; ab[k].z += d.z; ab[k].w += d.w;
                const float4 gm = gamma ? __ldg(gamma + c) : make_float4(1.f, 1.f, 1.f, 1.f);      // L1-resident
                g[k].x = d.x * gm.x; g[k].y = d.y * gm.y; g[k].z = d.z * gm.z; g[k].w = d.w * gm.w;
                s1 += (g[k].x + g[k].y) + (g[k].z + g[k].w);
                s2 += (g[k].x * xh[k].x + g[k].y * xh[k].y) + (g[k].z * xh[k].z + g[k].w * xh[k].w);
            }
        }
        const float2 ss = row_sum2<TPR>(s1, s2, scratch2);
        const float m1 = ss.x * inv_f, m2 = ss.y * inv_f;
#pragma unroll
        for (int k = 0; k < SLOTS; ++k) {
            const int c = lr + k * TPR;
            if (live && c < F4) {
                float4 o;
                o.x = rstd * (g[k].x - m1 - xh[k].x * m2);
                o.y = rstd * (g[k].y - m1 - xh[k].y * m2);
                o.z = rstd * (g[k].z - m1 - xh[k].z * m2);
                o.w = rstd * (g[k].w - m1 - xh[k].w * m2);
                stg_stream(reinterpret_cast<uint4 *>(dx + row * F4 + c),
                           make_uint4(__float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), __float_as_uint(o.w)));
            }
        }
    }
    if (TPR == 32) {
        // the CTA's 8 warps own the same columns: add them in warp order through shared memory (deterministic)
        for (int w = 0; w < RPC; ++w) {
            __syncthreads();
            if (grp == w) {
#pragma unroll
                for (int k = 0; k < SLOTS; ++k) {
                    const int c = lr + k * TPR;
                    if (c < F4) {
                        if (w == 0) { colsum[c] = ag[k]; colsum[F4 + c] = ab[k]; }
                        else {
                            float4 a = colsum[c], b = colsum[F4 + c];
                            a.x += ag[k].x; a.y += ag[k].y; a.z += ag[k].z; a.w += ag[k].w;
                            b.x += ab[k].x; b.y += ab[k].y; b.z += ab[k].z; b.w += ab[k].w;
                            colsum[c] = a; colsum[F4 + c] = b;
                        }
                    }
                }
            }
        }
        __syncthreads();
        for (int c = threadIdx.x; c < F4; c += LN_THREADS) {
            part_g[pblock * F4 + c] = colsum[c];
            part_b[pblock * F4 + c] = colsum[F4 + c];
        }
    } else {
#pragma unroll
        for (int k = 0; k < SLOTS; ++k) {
            const int c = lr + k * TPR;
            if (c < F4) {
                part_g[pblock * F4 + c] = ag[k];
                part_b[pblock * F4 + c] = ab[k];
            }
        }
    }
}

// dgamma[c] = sum_b part_g[b][c], dbeta likewise.  Block = 32 columns x 32 row groups: group r adds partial rows
// r, r+32, ... (coalesced 128-byte loads), then the 32 groups are combined in a fixed order through shared memory.
__global__ void __launch_bounds__(1024) ln_colreduce_kernel(int nblocks, int F, const float *__restrict__ part_g,
                                                            const float *__restrict__ part_b, float *__restrict__ dgamma,
                                                            float *__restrict__ dbeta)
{
    __shared__ float sg[32][33], sb[32][33];
    part_g += (size_t)blockIdx.y * nblocks * F;                 // grid.y = affine groups
    part_b += (size_t)blockIdx.y * nblocks * F;
    if (dgamma) dgamma += (size_t)blockIdx.y * F;
    if (dbeta) dbeta += (size_t)blockIdx.y * F;
    const int c = blockIdx.x * 32 + threadIdx.x, r = threadIdx.y;
    float a = 0.f, b = 0.f;
    if (c < F) {
        for (int k = r; k < nblocks; k += 32) {
            a += part_g[(size_t)k * F + c];
            b += part_b[(size_t)k * F + c];
        }
    }
    sg[r][threadIdx.x] = a;
    sb[r][threadIdx.x] = b;
    __syncthreads();
    if (r == 0 && c < F) {
        float ta = 0.f, tb = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) { ta += sg[k][threadIdx.x]; tb += sb[k][threadIdx.x]; }
        if (dgamma) dgamma[c] = ta;
        if (dbeta) dbeta[c] = tb;
    }
}

int ln_grid(long long rows, int rpc)
{
    long long nb = (rows + rpc - 1) / rpc;
    const long long cap = (long long)pb_sm_count() * 4;
    if (nb > cap) nb = cap;
    return (int)(nb < 1 ? 1 : nb);
}

}  // namespace

extern "C" {

int pb_layer_norm_supported(long long rows, int F) { return (rows > 0 && F >= 4 && (F % 4) == 0 && F <= 4096) ? 1 : 0; }

// Grouped form: `groups` affine pairs (gamma / beta are (groups, F)); output row r of `groups * rows_per_group` rows
// normalises input row r % x_rows and uses the affine pair r / rows_per_group.  Plain LayerNorm: groups = 1,
// x_rows = rows.  K ensemble heads on one shared input: groups = K, x_rows = rows_per_group = B.
int pb_layer_norm_grouped_fwd(int groups, long long rows_per_group, long long x_rows, int F, float eps, const float *x,
                              const float *gamma, const float *beta, float *y, float *mean_out, float *rstd_out,
                              void *stream)
{
    const long long rows = (long long)groups * rows_per_group;
    if (groups < 1 || rows_per_group < 1 || x_rows < 1 || !pb_layer_norm_supported(rows, F) || !x || !y) return PB_E_ARG;
    if ((((uintptr_t)x) | ((uintptr_t)y) | ((uintptr_t)gamma) | ((uintptr_t)beta)) & 15) return PB_E_ARG;
    const int F4 = F / 4;
    const float4 *x4 = reinterpret_cast<const float4 *>(x), *g4 = reinterpret_cast<const float4 *>(gamma),
                 *b4 = reinterpret_cast<const float4 *>(beta);
    float4 *y4 = reinterpret_cast<float4 *>(y);
    if (F <= 512) {
        const int nb = (int)((rows + 7) / 8);
        PB_LAUNCH((ln_fwd_kernel<32, 4>), nb, LN_THREADS, 0, stream, rows, F4, eps, x4, x_rows, g4, b4, rows_per_group, y4, mean_out, rstd_out);
    } else if (F <= 1024) {
        const int nb = (int)((rows + 7) / 8);
        PB_LAUNCH((ln_fwd_kernel<32, 8>), nb, LN_THREADS, 0, stream, rows, F4, eps, x4, x_rows, g4, b4, rows_per_group, y4, mean_out, rstd_out);
    } else {
        long long nb = rows;
        if (nb > (1LL << 30)) return PB_E_ARG;
        PB_LAUNCH((ln_fwd_kernel<256, 4>), (unsigned)nb, LN_THREADS, 0, stream, rows, F4, eps, x4, x_rows, g4, b4, rows_per_group, y4, mean_out, rstd_out);
    }
    return PB_OK;
}

int pb_layer_norm_fwd(long long rows, int F, float eps, const float *x, const float *gamma, const float *beta, float *y,
                      float *mean_out, float *rstd_out, void *stream)
{
    return pb_layer_norm_grouped_fwd(1, rows, rows, F, eps, x, gamma, beta, y, mean_out, rstd_out, stream);
}

// CTAs per group of the grouped backward (partials: 2 * groups * blocks * F floats)
int pb_layer_norm_grouped_bwd_blocks(int groups, long long rows_per_group, int F)
{
    int nb = ln_grid(rows_per_group, F <= 1024 ? LN_THREADS / 32 : 1);
    const int cap = (pb_sm_count() * 4 + groups - 1) / (groups > 0 ? groups : 1);
    if (groups > 1 && nb > cap) nb = cap < 1 ? 1 : cap;
    return nb;
}

// dx is per OUTPUT row (groups * rows_per_group, F): for heads that share their input the caller sums it over the
// groups (pb_sum_heads).  dgamma / dbeta: (groups, F).
int pb_layer_norm_grouped_bwd(int groups, long long rows_per_group, long long x_rows, int F, const float *x,
                              const float *dy, const float *gamma, const float *mean, const float *rstd, float *dx,
                              float *dgamma, float *dbeta, float *partials, void *stream)
{
    const long long rows = (long long)groups * rows_per_group;
    if (groups < 1 || rows_per_group < 1 || x_rows < 1 || !pb_layer_norm_supported(rows, F) || !x || !dy || !mean || !rstd || !dx || !partials)
        return PB_E_ARG;
    if ((((uintptr_t)x) | ((uintptr_t)dy) | ((uintptr_t)dx) | ((uintptr_t)gamma) | ((uintptr_t)partials)) & 15) return PB_E_ARG;
    const int F4 = F / 4;
    const int nb = pb_layer_norm_grouped_bwd_blocks(groups, rows_per_group, F);
    const float4 *x4 = reinterpret_cast<const float4 *>(x), *d4 = reinterpret_cast<const float4 *>(dy),
                 *g4 = reinterpret_cast<const float4 *>(gamma);
    float4 *dx4 = reinterpret_cast<float4 *>(dx), *pg = reinterpret_cast<float4 *>(partials),
           *pbeta = reinterpret_cast<float4 *>(partials + (size_t)groups * nb * F);
    const dim3 grid(nb, groups);
    if (F <= 512) {
        PB_LAUNCH((ln_bwd_kernel<32, 4>), grid, LN_THREADS, 2 * F * sizeof(float), stream, rows_per_group, F4, x4, x_rows, d4, g4, mean, rstd, dx4, pg, pbeta);
    } else if (F <= 1024) {
        PB_LAUNCH((ln_bwd_kernel<32, 8>), grid, LN_THREADS, 2 * F * sizeof(float), stream, rows_per_group, F4, x4, x_rows, d4, g4, mean, rstd, dx4, pg, pbeta);
    } else {
        PB_LAUNCH((ln_bwd_kernel<256, 4>), grid, LN_THREADS, 0, stream, rows_per_group, F4, x4, x_rows, d4, g4, mean, rstd, dx4, pg, pbeta);
    }
    if (dgamma || dbeta)
        PB_LAUNCH(ln_colreduce_kernel, dim3((F + 31) / 32, groups), dim3(32, 32), 0, stream, nb, F, partials,
                  partials + (size_t)groups * nb * F, dgamma, dbeta);
    return PB_OK;
}

int pb_layer_norm_bwd_blocks(long long rows, int F) { return pb_layer_norm_grouped_bwd_blocks(1, rows, F); }

// partials: 2 * pb_layer_norm_bwd_blocks(rows, F) * F floats of scratch
int pb_layer_norm_bwd(long long rows, int F, const float *x, const float *dy, const float *gamma, const float *mean,
                      const float *rstd, float *dx, float *dgamma, float *dbeta, float *partials, void *stream)
{
    return pb_layer_norm_grouped_bwd(1, rows, rows, F, x, dy, gamma, mean, rstd, dx, dgamma, dbeta, partials, stream);
}

}  // extern "C"
