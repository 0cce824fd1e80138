// tc_gemm.cu -- tcgen05 (5th-gen tensor core) GEMM with fp32-level accuracy for the IQN / ensemble
// dense layers (sm_100a only).
//
//     C[b] (M x N) = act( A[b] (M x K) . B[b]^T (N x K) + bias[b] )          fp32 in, fp32 out
//
// The reference computes these layers in fp32 (nn.Linear, prism/agents/models/ffnn_model.py:61-76,
// iqn_model.py:30-46, 89-93, q_ensemble.py:26-48) and parity is judged at 1e-4 relative on losses and
// gradients, which a single TF32 or BF16 pass (~1e-3) cannot hold.  So every operand is split on the fly
// into a TF32 "hi" part and a TF32 "lo" residual (x = hi + lo exactly) and each product is issued as three
// tensor-core MMAs  hi.hi + hi.lo + lo.hi  accumulating in fp32 in TENSOR MEMORY -- error ~1e-6, at a third
// of the TF32 rate, i.e. several times the SIMT fp32 rate the reference's cuBLAS path gets.
//
// Structure (one CTA = one 128 x BN output tile, 128 threads):
//   * all four warps stream the fp32 operand tiles from global memory (128-bit loads), split them and
//     write the hi / lo tiles to shared memory in the canonical K-major no-swizzle UMMA layout
//     (8-row x 16-byte core matrices);
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M = 128, N = BN, K = 8 per instruction)
//     from shared-memory descriptors; the accumulator (128 lanes x BN columns) lives in TMEM;
//   * tcgen05.commit -> mbarrier tells the loaders when a stage may be overwritten (2 stages: the tensor
//     core works on stage s while the warps fill stage s^1);
//   * epilogue: tcgen05.ld 32x32b (each warp its own 32 TMEM lanes) -> bias + ReLU -> global.
#include "common.cuh"

namespace {

using namespace pb;

constexpr int BM = 128, BK = 32, STAGES = 2;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred P1;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
                     "selp.b32 %0, 1, 0, P1;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 B (rows 16 B apart);
// SBO = byte distance between 8-row groups, LBO = byte distance between the two 16-byte K chunks of one MMA
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                       // descriptor version for sm_100
    return d;                                     // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

// instruction descriptor: D fp32, A/B tf32, both K-major, M = 128, N = BN
__host__ __device__ constexpr uint32_t make_idesc(int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                 ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
                 : "memory");
}

__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo)
{
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));     // round-to-nearest TF32 (low 13 mantissa bits zero)
    hi = __uint_as_float(h);
    lo = x - hi;                                            // exact; the tensor core truncates it to TF32
}

struct TcArgs {
    const float *A, *B, *bias;
    float *C;
    long long a_bs, b_bs, c_bs, bias_bs;
    int M, N, K, act;
};

// tile of R rows x BK floats: chunk c (16 B = 4 floats of K) of row r lives at c*(R*16) + r*16
template <int R>
__device__ __forceinline__ void fill_tile(float *hi_tile, float *lo_tile, const float *__restrict__ src, int ld,
                                          int row0, int rows, int k0, int t)
{
#pragma unroll
    for (int i = 0; i < R * (BK / 4) / 128; ++i) {
        const int e = t + i * 128, r = e >> 3, c = e & 7;           // 8 threads cover one row's 128 bytes
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < rows) v = *reinterpret_cast<const float4 *>(src + (size_t)(row0 + r) * ld + k0 + c * 4);
        float4 h, l;
        split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
        const int off = c * (R * 4) + r * 4;                         // in floats
        *reinterpret_cast<float4 *>(hi_tile + off) = h;
        *reinterpret_cast<float4 *>(lo_tile + off) = l;
    }
}

template <int BN>
__global__ void __launch_bounds__(128) tc_gemm_kernel(TcArgs g)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    constexpr int A_TILE = BM * BK, B_TILE = BN * BK;               // floats
    float *tiles = reinterpret_cast<float *>(smem_raw);
    // per stage: A_hi | A_lo | B_hi | B_lo
    constexpr int STAGE_FLOATS = 2 * A_TILE + 2 * B_TILE;
    __shared__ uint64_t bar_free[STAGES];
    __shared__ uint64_t bar_done;
    __shared__ uint32_t tmem_base_slot;

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM, batch = blockIdx.z;
    const float *A = g.A + batch * g.a_bs, *B = g.B + batch * g.b_bs;
    float *C = g.C + batch * g.c_bs;
    const float *bias = g.bias ? g.bias + batch * g.bias_bs : nullptr;

    if (t == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&bar_free[s], 1);
        mbar_init(&bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"((uint32_t)BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_c = tmem_base_slot;
    constexpr uint32_t idesc = make_idesc(BN);

    const int n_kb = g.K / BK;
    for (int kb = 0; kb < n_kb; ++kb) {
        const int s = kb % STAGES, use = kb / STAGES;
        if (use > 0) mbar_wait(&bar_free[s], (uint32_t)((use - 1) & 1));   // the MMAs that read this stage are done
        float *a_hi = tiles + s * STAGE_FLOATS, *a_lo = a_hi + A_TILE, *b_hi = a_lo + A_TILE, *b_lo = b_hi + B_TILE;
        fill_tile<BM>(a_hi, a_lo, A, g.K, m0, g.M, kb * BK, t);
        fill_tile<BN>(b_hi, b_lo, B, g.K, n0, g.N, kb * BK, t);
        fence_async_smem();                                         // generic-proxy writes -> visible to the tensor core
        __syncthreads();
        if (t == 0) {
            tc_fence_after();
            const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
            for (int ks = 0; ks < BK / 8; ++ks) {
                const uint32_t ao = 2 * ks * (BM * 16), bo = 2 * ks * (BN * 16);     // two 16-byte K chunks per MMA
                const uint64_t dah = make_desc(ah + ao, BM * 16, 128), dal = make_desc(al + ao, BM * 16, 128);
                const uint64_t dbh = make_desc(bh + bo, BN * 16, 128), dbl = make_desc(bl + bo, BN * 16, 128);
                umma_tf32(tmem_c, dal, dbh, idesc, (kb > 0 || ks > 0) ? 1u : 0u);    // small terms first
                umma_tf32(tmem_c, dah, dbl, idesc, 1u);
                umma_tf32(tmem_c, dah, dbh, idesc, 1u);
            }
            umma_commit(&bar_free[s]);                              // arrives when the MMAs above have finished
            if (kb == n_kb - 1) umma_commit(&bar_done);
        }
    }
    mbar_wait(&bar_done, 0);
    tc_fence_after();

    // epilogue: warp w owns TMEM lanes [32w, 32w+32) = output rows m0 + 32w + lane
    const int row = m0 + warp * 32 + lane;
#pragma unroll 1
    for (int cc = 0; cc < BN / 32; ++cc) {
        uint32_t r[32];
        const uint32_t taddr = tmem_c + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cc * 32);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32"
                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
                     " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                       "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                       "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < g.M) {
            float *crow = C + (size_t)row * g.N + n0 + cc * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                       __uint_as_float(r[j + 3]));
                if (bias) {
                    const float4 bb = *reinterpret_cast<const float4 *>(bias + n0 + cc * 32 + j);
                    v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
                }
                if (g.act == 1) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                *reinterpret_cast<float4 *>(crow + j) = v;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_c), "r"((uint32_t)BN) : "memory");
}

template <int BN>
int launch_tc(const TcArgs &g, int batch, void *stream)
{
    static bool attr_set = false;
    const size_t smem = sizeof(float) * STAGES * (2 * BM * BK + 2 * BN * BK);
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    dim3 grid((unsigned)(g.N / BN), (unsigned)((g.M + BM - 1) / BM), (unsigned)batch);
    PB_LAUNCH(tc_gemm_kernel<BN>, grid, 128, smem, stream, g);
    return PB_OK;
}

}  // namespace

extern "C" {

// 1 when pb_linear_fwd_tc accepts the shape (K % 32 == 0, N % 64 == 0, 16-byte aligned rows)
int pb_linear_fwd_tc_supported(int M, int N, int J)
{
    return (M > 0 && N > 0 && J > 0 && (J % BK) == 0 && (N % 64) == 0) ? 1 : 0;
}

// Y[k] (M x N) = act(X[k] (M x J) W[k]^T (N x J) + b[k]) on the tensor cores (3xTF32, fp32-level accuracy)
int pb_linear_fwd_tc(int K, int M, int N, int J, const float *X, long long x_head_stride, const float *W,
                     const float *b, int act, float *Y, void *stream)
{
    if (K <= 0 || !X || !W || !Y || (act != 0 && act != 1)) return PB_E_ARG;
    if (!pb_linear_fwd_tc_supported(M, N, J)) return PB_E_UNSUPPORTED;
    if ((((uintptr_t)X) | ((uintptr_t)W) | ((uintptr_t)Y) | ((uintptr_t)b)) & 15) return PB_E_ARG;
    TcArgs g = {};
    g.A = X; g.a_bs = x_head_stride;
    g.B = W; g.b_bs = (long long)N * J;
    g.C = Y; g.c_bs = (long long)M * N;
    g.bias = b; g.bias_bs = N;
    g.M = M; g.N = N; g.K = J; g.act = act;
    if (N % 128 == 0) return launch_tc<128>(g, K, stream);
    return launch_tc<64>(g, K, stream);
}

}  // extern "C"
