// conv.cu -- MinAtar embedding network as one launch: conv3x3 (stride 1, no padding) on a
// channels-last observation + bias + ReLU + NCHW flatten (sm_100a).
//
// Reference: MinAtarModel (prism/agents/models/minatar_cnn_model.py:13-18, 41-44):
//     x.permute(0,3,1,2) -> Conv2d(C,16,3,1) -> ReLU -> Flatten.
// Through PyTorch/cuDNN that is 5 launches forward (layout copy, NHWC conv engine, bias add, clamp,
// flatten copy) and ~8 backward for a 14 MFLOP problem.  Here: one CTA per observation holds the
// 10x10xC frame and the whole filter bank in shared memory.  The observation needs no gradient
// (it comes from the replay ring), so backward is only dW / db: one CTA per group of samples
// accumulates its share in registers, partial sums are combined by a second tiny launch in a fixed
// order (deterministic).
#include "common.cuh"

namespace {

using namespace pb;

constexpr int MAX_X = 16 * 16 * 16;      // floats of one frame in smem
constexpr int MAX_W = 32 * 16 * 9;       // floats of the filter bank in smem
constexpr int MAX_OUT = 32 * 14 * 14;

struct ConvDims { int B, H, W, C, OC, OH, OW; };

__global__ void __launch_bounds__(256) conv3x3_relu_fwd_kernel(ConvDims d, const float *__restrict__ x,
                                                               const float *__restrict__ w,
                                                               const float *__restrict__ bias, float *__restrict__ out)
{
    extern __shared__ float sm[];
    float *sx = sm, *sw = sm + d.H * d.W * d.C;
    const int b = blockIdx.x;
    const int nx = d.H * d.W * d.C, nw = d.OC * d.C * 9;
    const float *xb = x + (size_t)b * nx;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) sx[i] = xb[i];
    for (int i = threadIdx.x; i < nw; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const int plane = d.OH * d.OW, n_out = d.OC * plane;
    float *ob = out + (size_t)b * n_out;
    for (int o = threadIdx.x; o < n_out; o += blockDim.x) {
        const int oc = o / plane, p = o - oc * plane, y = p / d.OW, xx = p - y * d.OW;
        float acc = bias ? bias[oc] : 0.0f;
        const float *wo = sw + oc * d.C * 9;
        for (int ky = 0; ky < 3; ++ky)
            for (int kx = 0; kx < 3; ++kx) {
                const float *px = sx + ((y + ky) * d.W + (xx + kx)) * d.C;
                for (int c = 0; c < d.C; ++c) acc = fmaf(px[c], wo[c * 9 + ky * 3 + kx], acc);
            }
        ob[o] = fmaxf(acc, 0.0f);
    }
}

// partial[g][e], e in [0, OC*C*9) = dW, e in [OC*C*9, OC*C*9+OC) = db, for the samples of CTA g
__global__ void __launch_bounds__(256) conv3x3_relu_bwd_kernel(ConvDims d, const float *__restrict__ x,
                                                               const float *__restrict__ out,
                                                               const float *__restrict__ dout,
                                                               float *__restrict__ partial)
{
    extern __shared__ float sm[];
    const int nx = d.H * d.W * d.C, plane = d.OH * d.OW, n_out = d.OC * plane;
    float *sx = sm, *sg = sm + nx;
    const int nw = d.OC * d.C * 9, ne = nw + d.OC;
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.0f;
    for (int b = blockIdx.x; b < d.B; b += gridDim.x) {
        __syncthreads();
        const float *xb = x + (size_t)b * nx;
        for (int i = threadIdx.x; i < nx; i += blockDim.x) sx[i] = xb[i];
        const float *ob = out + (size_t)b * n_out, *gb = dout + (size_t)b * n_out;
        for (int i = threadIdx.x; i < n_out; i += blockDim.x) sg[i] = ob[i] > 0.0f ? gb[i] : 0.0f;   // ReLU mask
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int e = threadIdx.x + q * 256;
            if (e >= ne) break;
            float a = 0.0f;
            if (e < nw) {
                const int oc = e / (d.C * 9), r = e - oc * d.C * 9, c = r / 9, k = r - c * 9, ky = k / 3, kx = k - ky * 3;
                const float *g = sg + oc * plane;
                for (int y = 0; y < d.OH; ++y)
                    for (int xx = 0; xx < d.OW; ++xx)
                        a = fmaf(g[y * d.OW + xx], sx[((y + ky) * d.W + (xx + kx)) * d.C + c], a);
            } else {
                const float *g = sg + (e - nw) * plane;
                for (int p = 0; p < plane; ++p) a += g[p];
            }
            acc[q] += a;
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int e = threadIdx.x + q * 256;
        if (e < ne) partial[(size_t)blockIdx.x * ne + e] = acc[q];
    }
}

__global__ void conv_bwd_reduce_kernel(int n_groups, int nw, int n_bias, const float *__restrict__ partial,
                                       float *__restrict__ dw, float *__restrict__ db)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x, ne = nw + n_bias;
    if (e >= ne) return;
    float a = 0.0f;
    for (int g = 0; g < n_groups; ++g) a += partial[(size_t)g * ne + e];
    if (e < nw) dw[e] = a; else if (db) db[e - nw] = a;
}

int check_dims(const ConvDims &d)
{
    if (d.B <= 0 || d.H < 3 || d.W < 3 || d.C <= 0 || d.OC <= 0) return PB_E_ARG;
    if (d.H * d.W * d.C > MAX_X || d.OC * d.C * 9 > MAX_W || d.OC * d.OH * d.OW > MAX_OUT) return PB_E_UNSUPPORTED;
    if (d.OC * d.C * 9 + d.OC > 8 * 256) return PB_E_UNSUPPORTED;
    return PB_OK;
}

}  // namespace

extern "C" {

int pb_conv3x3_relu_fwd(int B, int H, int W, int C, int OC, const float *x, const float *w, const float *bias,
                        float *out, void *stream)
{
    ConvDims d = {B, H, W, C, OC, H - 2, W - 2};
    int rc = check_dims(d);
    if (rc) return rc;
    if (!x || !w || !out) return PB_E_ARG;
    const size_t smem = sizeof(float) * (size_t)(H * W * C + OC * C * 9);
    PB_LAUNCH(conv3x3_relu_fwd_kernel, (unsigned)B, 256, smem, stream, d, x, w, bias, out);
    return PB_OK;
}

int pb_conv3x3_relu_bwd_groups(int B) { return B < 128 ? B : 128; }

int pb_conv3x3_relu_bwd(int B, int H, int W, int C, int OC, const float *x, const float *out, const float *dout,
                        float *partial_scratch, float *dw, float *db, void *stream)
{
    ConvDims d = {B, H, W, C, OC, H - 2, W - 2};
    int rc = check_dims(d);
    if (rc) return rc;
    if (!x || !out || !dout || !partial_scratch || !dw) return PB_E_ARG;
    const int groups = pb_conv3x3_relu_bwd_groups(B);
    const size_t smem = sizeof(float) * (size_t)(H * W * C + OC * d.OH * d.OW);
    if (smem > 48 * 1024) return PB_E_UNSUPPORTED;
    PB_LAUNCH(conv3x3_relu_bwd_kernel, (unsigned)groups, 256, smem, stream, d, x, out, dout, partial_scratch);
    const int ne = OC * C * 9 + OC;
    PB_LAUNCH(conv_bwd_reduce_kernel, (unsigned)((ne + 255) / 256), 256, 0, stream, groups, OC * C * 9, OC,
              partial_scratch, dw, db);
    return PB_OK;
}

}  // extern "C"
