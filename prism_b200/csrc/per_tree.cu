// per_tree.cu -- sum-tree / min-tree priority store for prioritized replay (sm_100a).
//
// Replaces the segment tree + PrioritizedSampler that the reference reaches through
// torchrl (prism/factory/exp_buffer_factory.py:22-28; call sites
// prism/experience/timestep_buffer.py:33,37,54; prism/learner.py:100,120).
//
// Design (B200-first, not a port of a pointer-walking CPU tree):
//  * level-ordered fp32 heap array; node = fl32(left+right) -> the reference add order,
//    so sampled indices are bit-exact whatever order updates arrive in.
//  * one primitive everywhere: a warp loads one aligned 128-byte line (32 sibling nodes)
//    and rebuilds the 5 levels above it with xor-shuffles.  Those rebuilt values are
//    bit-identical to the stored internal nodes, so
//      - sampling descends 5 levels per dependent memory round trip (5 trips for 2^24
//        leaves instead of 24),
//      - a batched priority update needs 2 sparse phases + 1 single-CTA top phase (the
//        top <=14 levels live in shared memory) instead of 24 level-synchronous steps,
//      - the bulk build streams leaves once at HBM speed.
//  * duplicates in an update batch: last occurrence wins (sequential reference loop),
//    resolved deterministically (owner scratch + atomicMax, or adjacency when sorted).
//  * no host sync anywhere: len / cursor / max_priority / p_sum / p_min live in a 64-byte
//    device state block, so the whole sample->update loop is CUDA-graph capturable.
#include "common.cuh"
#include <math.h>

namespace {

using namespace pb;

struct TreeView {
    float *sum, *min;
    int *owner;
    pb_per_state *st;
    long long cap, size;
    int L;
    float alpha, eps32;
    double eps64;
    int weps, dp64;
};

enum { MODE_RAW = 0, MODE_PRIORITY = 1, MODE_EXTEND = 2 };
constexpr int TOP_MAX = 14;       // top kernel holds depths [0, T], T <= 14 -> 2^(T+1)*4 B = 128 KB smem
constexpr int STAGE_TILE = 2048;  // bulk build: source nodes per CTA (11 levels)
constexpr long long THREAD_MODE_MIN = 16384;  // samples per call from which one-thread-per-sample wins

__device__ __forceinline__ float op_sum(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float op_min(float a, float b) { return fminf(a, b); }

__device__ __forceinline__ float pow_leaf(float p, const TreeView &t)
{
    // torch.pow(priority + eps, alpha) on an fp32 tensor; alpha == 0.5 is torch's sqrt path
    float x = __fadd_rn(p, t.eps32);
    if (t.alpha == 0.5f) return __fsqrt_rn(x);
    if (t.alpha == 1.0f) return x;
    return powf(x, t.alpha);
}

__device__ __forceinline__ float default_priority(const TreeView &t)
{
    float mp = t.st->max_priority;
    if (t.dp64) {
        double x = (double)mp + t.eps64;
        double r = (t.alpha == 0.5f) ? sqrt(x) : ((t.alpha == 1.0f) ? x : pow(x, (double)t.alpha));
        return (float)r;
    }
    return pow_leaf(mp, t);
}

__device__ __forceinline__ long long entry_index(const TreeView &t, const long long *idx, long long j,
                                                 int mode)
{
    if (mode == MODE_EXTEND) return (t.st->seq + j) % t.size;
    return idx[j];
}

__device__ __forceinline__ float entry_leaf(const TreeView &t, const float *val, long long j, int mode,
                                            float defp)
{
    if (mode == MODE_RAW) return val[j];
    if (mode == MODE_PRIORITY) return pow_leaf(fabsf(val[j]), t);   // learner.py:120 passes |td|; idempotent
    return defp;
}

// ---------------------------------------------------------------------------------
// init
// ---------------------------------------------------------------------------------
__global__ void tree_init_kernel(TreeView t)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    const float inf = __int_as_float(0x7f800000);
    for (long long k = i; k < 2 * t.cap; k += stride) { t.sum[k] = 0.0f; t.min[k] = inf; }
    if (t.owner) for (long long k = i; k < t.cap; k += stride) t.owner[k] = -1;
    if (i == 0) {
        pb_per_state s;
        s.len = 0; s.seq = 0; s.max_priority = 1.0f; s.p_sum = 0.0f; s.p_min = inf; s.status = 0;
        s.batch_max = 0.0f; s.owned_lo = 0; s.owned_n = 0;
        for (int k = 0; k < 5; ++k) s.pad[k] = 0;
        *t.st = s;
    }
}

// ---------------------------------------------------------------------------------
// query(0, len) with torchrl's interval-walk association (one warp)
// ---------------------------------------------------------------------------------
template <bool IS_MIN>
__device__ float tree_query_prefix(const TreeView &t, long long len)
{
    const float *tree = IS_MIN ? t.min : t.sum;
    const float ident = IS_MIN ? __int_as_float(0x7f800000) : 0.0f;
    if (len >= t.size) return tree[1];
    if (len <= 0) return ident;
    int k = lane_id();
    float v = ident;
    if (k < t.L && ((len >> k) & 1)) v = tree[((t.cap + len) >> k) - 1];
    float ret = ident;
    for (int b = 0; b < t.L; ++b) {
        float vb = __shfl_sync(FULL, v, b);
        if ((len >> b) & 1) ret = IS_MIN ? op_min(ret, vb) : op_sum(ret, vb);
    }
    return ret;
}

__global__ void tree_stats_kernel(TreeView t)
{
    long long len = t.st->len;
    float ps = tree_query_prefix<false>(t, len);
    float pm = tree_query_prefix<true>(t, len);
    if (threadIdx.x == 0) { t.st->p_sum = ps; t.st->p_min = pm; }
}

// ---------------------------------------------------------------------------------
// bulk build, stage 1: each CTA reduces 2048 source nodes by 11 levels (both trees via
// blockIdx.y).  Streaming: one 2x float4 load per thread, coalesced stores per level.
// ---------------------------------------------------------------------------------
template <bool FROM_LEAVES>
__global__ void __launch_bounds__(256) tree_reduce11_kernel(TreeView t, int d_src, const float *leaves,
                                                            long long n_leaves)
{
    const bool is_min = blockIdx.y != 0;
    float *tree = is_min ? t.min : t.sum;
    const float ident = is_min ? __int_as_float(0x7f800000) : 0.0f;
    const long long n_src = 1LL << d_src;
    const long long base = (long long)blockIdx.x * STAGE_TILE + (long long)threadIdx.x * 8;
    float a[8];
    if (FROM_LEAVES) {
        if (base + 8 <= n_leaves) {
            float4 x = *reinterpret_cast<const float4 *>(leaves + base);
            float4 y = *reinterpret_cast<const float4 *>(leaves + base + 4);
            a[0] = x.x; a[1] = x.y; a[2] = x.z; a[3] = x.w; a[4] = y.x; a[5] = y.y; a[6] = y.z; a[7] = y.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = (base + i < n_leaves) ? leaves[base + i] : ident;
        }
        float4 *dst = reinterpret_cast<float4 *>(tree + n_src + base);
        dst[0] = make_float4(a[0], a[1], a[2], a[3]);
        dst[1] = make_float4(a[4], a[5], a[6], a[7]);
    } else {
        const float4 *src = reinterpret_cast<const float4 *>(tree + n_src + base);
        float4 x = src[0], y = src[1];
        a[0] = x.x; a[1] = x.y; a[2] = x.z; a[3] = x.w; a[4] = y.x; a[5] = y.y; a[6] = y.z; a[7] = y.w;
    }
    auto op = [&](float l, float r) { return is_min ? op_min(l, r) : op_sum(l, r); };
    float b0 = op(a[0], a[1]), b1 = op(a[2], a[3]), b2 = op(a[4], a[5]), b3 = op(a[6], a[7]);
    *reinterpret_cast<float4 *>(tree + (n_src >> 1) + (base >> 1)) = make_float4(b0, b1, b2, b3);
    float c0 = op(b0, b1), c1 = op(b2, b3);
    *reinterpret_cast<float2 *>(tree + (n_src >> 2) + (base >> 2)) = make_float2(c0, c1);
    float d = op(c0, c1);
    const long long e = base >> 3;  // element index at depth d_src-3
    tree[(n_src >> 3) + e] = d;
    const int lane = lane_id();
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        d = op(d, __shfl_xor_sync(FULL, d, 1 << s));
        if ((lane & ((2 << s) - 1)) == 0) tree[(n_src >> (4 + s)) + (e >> (s + 1))] = d;
    }
    __shared__ float warp_part[8];
    if (lane == 0) warp_part[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x < 32) {
        float x = lane < 8 ? warp_part[lane] : ident;
        const long long e8 = (long long)blockIdx.x * 8 + lane;  // element index at depth d_src-8
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            x = op(x, __shfl_xor_sync(FULL, x, 1 << s));
            if (lane < 8 && (lane & ((2 << s) - 1)) == 0) tree[(n_src >> (9 + s)) + (e8 >> (s + 1))] = x;
        }
    }
}

// small trees (L <= 14): copy leaves (identity padded) into both leaf regions
__global__ void tree_fill_leaves_kernel(TreeView t, const float *leaves, long long n_leaves)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t.cap) return;
    bool in = i < n_leaves;
    float v = in ? leaves[i] : 0.0f;
    t.sum[t.cap + i] = in ? v : 0.0f;
    t.min[t.cap + i] = in ? v : __int_as_float(0x7f800000);
}

// ---------------------------------------------------------------------------------
// top phase: one CTA per tree holds depths [0, T] in shared memory (heap layout),
// reduces them pairwise, writes nodes [1, 2^T) back, and the last CTA to finish
// finalises the state block (len/seq advance, max_priority merge, p_sum/p_min).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) tree_top_kernel(TreeView t, int T, int mode, long long n_new,
                                                        long long set_len)
{
    extern __shared__ float sm[];
    const bool is_min = blockIdx.x != 0;
    float *tree = is_min ? t.min : t.sum;
    const int n_top = 1 << T;
    if (T >= 2) {
        const float4 *src = reinterpret_cast<const float4 *>(tree + n_top);
        float4 *dst = reinterpret_cast<float4 *>(sm + n_top);
        for (int i = threadIdx.x; i < (n_top >> 2); i += blockDim.x) dst[i] = src[i];
    } else {
        for (int i = threadIdx.x; i < n_top; i += blockDim.x) sm[n_top + i] = tree[n_top + i];
    }
    __syncthreads();
    for (int d = T - 1; d >= 0; --d) {
        const int n_d = 1 << d;
        for (int i = threadIdx.x; i < n_d; i += blockDim.x) {
            float l = sm[2 * (n_d + i)], r = sm[2 * (n_d + i) + 1];
            sm[n_d + i] = is_min ? op_min(l, r) : op_sum(l, r);
        }
        __syncthreads();
    }
    for (int i = threadIdx.x + 1; i < n_top; i += blockDim.x) tree[i] = sm[i];
    __threadfence();
    __syncthreads();
    __shared__ int is_last;
    if (threadIdx.x == 0) {
        int ticket = atomicAdd(&t.st->pad[0], 1);
        is_last = (ticket == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last || threadIdx.x >= 32) return;
    __threadfence();
    // finalise (one warp of the last CTA)
    long long len = t.st->len;
    if (mode == MODE_EXTEND) {
        len = len + n_new < t.size ? len + n_new : t.size;
    } else if (set_len >= 0) {
        len = set_len < t.size ? set_len : t.size;
    }
    float ps = tree_query_prefix<false>(t, len);
    float pm = tree_query_prefix<true>(t, len);
    if (threadIdx.x == 0) {
        pb_per_state *s = t.st;
        if (mode == MODE_EXTEND) s->seq += n_new;
        else if (set_len >= 0) s->seq = set_len;
        s->len = len;
        if (mode == MODE_PRIORITY) {
            float bm = s->batch_max;
            if (bm > s->max_priority) s->max_priority = bm;
        }
        s->batch_max = 0.0f;
        s->p_sum = ps; s->p_min = pm;
        s->pad[0] = 0;
    }
}

// ---------------------------------------------------------------------------------
// sparse update, general (unsorted) path: A) mark owner = last occurrence, B) winner
// writes the leaf and clears the scratch.
// ---------------------------------------------------------------------------------
__global__ void upd_mark_kernel(TreeView t, long long n, const long long *idx, const float *val, int mode)
{
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float bm = 0.0f;
    if (j < n) {
        long long i = entry_index(t, idx, j, mode);
        if (i >= 0 && i < t.size) {
            atomicMax(&t.owner[i], (int)j);
            if (mode == MODE_PRIORITY) bm = fabsf(val[j]);
        }
    }
    if (mode == MODE_PRIORITY) {
        bm = warp_max(bm);
        if (lane_id() == 0 && bm > 0.0f) atomicMax(reinterpret_cast<int *>(&t.st->batch_max), __float_as_int(bm));
    }
}

__global__ void upd_leaf_kernel(TreeView t, long long n, const long long *idx, const float *val, int mode,
                                long long *idx_out)
{
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    long long i = entry_index(t, idx, j, mode);
    if (idx_out) idx_out[j] = i;
    if (i < 0 || i >= t.size) return;
    if (t.owner[i] != (int)j) return;
    float defp = (mode == MODE_EXTEND) ? default_priority(t) : 0.0f;
    float v = entry_leaf(t, val, j, mode, defp);
    t.sum[t.cap + i] = v;
    t.min[t.cap + i] = v;
    t.owner[i] = -1;
}

// dense path, sorted input: the last entry of every run of equal indices writes the leaf
__global__ void upd_leaf_sorted_kernel(TreeView t, long long n, const long long *idx, const float *val, int mode,
                                       long long *idx_out)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float bm = 0.0f;
    if (j < n) {
        const long long i = entry_index(t, idx, j, mode);
        if (idx_out) idx_out[j] = i;
        if (i >= 0 && i < t.size) {
            if (mode == MODE_PRIORITY) bm = fabsf(val[j]);
            const long long nx = (j + 1 < n) ? entry_index(t, idx, j + 1, mode) : -1;
            if (nx != i) {
                const float defp = (mode == MODE_EXTEND) ? default_priority(t) : 0.0f;
                const float v = entry_leaf(t, val, j, mode, defp);
                t.sum[t.cap + i] = v;
                t.min[t.cap + i] = v;
            }
        }
    }
    if (mode == MODE_PRIORITY) {
        bm = warp_max(bm);
        if (lane_id() == 0 && bm > 0.0f) atomicMax(reinterpret_cast<int *>(&t.st->batch_max), __float_as_int(bm));
    }
}

// ---------------------------------------------------------------------------------
// sparse phase p: one warp per batch entry; loads the 32 sibling nodes at depth
// d = L-5p that contain the entry's ancestor and rebuilds the 5 levels above with
// shuffles.  FUSED (sorted input, p == 0): the run leader also applies every leaf of
// its run (sequentially -> last wins) before reducing.
// ---------------------------------------------------------------------------------
template <bool FUSED>
__device__ __forceinline__ void sparse_entry(const TreeView &t, long long n, const long long *idx, const float *val,
                                             int mode, int p, int sorted, long long *idx_out, long long j)
{
    const int lane = lane_id();
    const int sh = 5 * p + 5;
    const long long i = entry_index(t, idx, j, mode);
    if (i < 0 || i >= t.size) return;
    const long long g = i >> sh;
    if (sorted && j > 0) {
        long long ip = entry_index(t, idx, j - 1, mode);
        if (ip >= 0 && ip < t.size && (ip >> sh) == g) return;  // the run leader does this line
    }
    const int d = t.L - 5 * p;
    const long long src = (1LL << d) + (g << 5) + lane;
    float vs = t.sum[src], vm = t.min[src];
    if (FUSED) {
        const float defp = (mode == MODE_EXTEND) ? default_priority(t) : 0.0f;
        float bm = 0.0f;
        bool touched = false;
        for (long long jj = j; jj < n; ++jj) {
            long long ii = entry_index(t, idx, jj, mode);
            if (ii < 0 || ii >= t.size || (ii >> 5) != g) break;
            float v = entry_leaf(t, val, jj, mode, defp);
            if (mode == MODE_PRIORITY) bm = fmaxf(bm, fabsf(val[jj]));
            if (lane == (int)(ii & 31)) { vs = v; vm = v; touched = true; }
            if (idx_out && lane == 0) idx_out[jj] = ii;
        }
        if (touched) { t.sum[src] = vs; t.min[src] = vm; }
        if (mode == MODE_PRIORITY && lane == 0 && bm > 0.0f)
            atomicMax(reinterpret_cast<int *>(&t.st->batch_max), __float_as_int(bm));
    }
    const long long e = (g << 5) + lane;
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        vs = op_sum(vs, __shfl_xor_sync(FULL, vs, 1 << s));
        vm = op_min(vm, __shfl_xor_sync(FULL, vm, 1 << s));
        if ((lane & ((2 << s) - 1)) == 0) {
            long long node = (1LL << (d - s - 1)) + (e >> (s + 1));
            t.sum[node] = vs;
            t.min[node] = vm;
        }
    }
}

template <bool FUSED>
__global__ void __launch_bounds__(256) upd_sparse_kernel(TreeView t, long long n, const long long *idx,
                                                         const float *val, int mode, int p, int sorted,
                                                         long long *idx_out)
{
    const long long j = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= n) return;
    sparse_entry<FUSED>(t, n, idx, val, mode, p, sorted, idx_out, j);
}

// ---------------------------------------------------------------------------------
// latency mode (n <= 64: the handful of new transitions of one learner iteration): the whole update -- dedup,
// leaves, every sparse phase, the top levels of both trees, the state block -- in ONE single-CTA launch,
// phases separated by __syncthreads instead of kernel boundaries.
// ---------------------------------------------------------------------------------
constexpr int SMALL_MAX = 64;      // beyond ~2 entries per warp the serial in-CTA loop loses to 3 wide launches (measured: 29 vs 17 us at n=256)
constexpr int SMALL_TOP = 12;      // both trees' top levels in shared memory: 2 * 2^(12+1) * 4 B = 64 KB

__global__ void __launch_bounds__(1024) upd_small_kernel(TreeView t, long long n, const long long *idx,
                                                         const float *val, int mode, int sorted, int P, int T,
                                                         long long *idx_out)
{
    extern __shared__ float sm[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool fused = sorted && P > 0;
    if (!fused) {
        float bm = 0.0f;
        for (long long j = tid; j < n; j += blockDim.x) {
            const long long i = entry_index(t, idx, j, mode);
            if (i >= 0 && i < t.size) {
                atomicMax(&t.owner[i], (int)j);
                if (mode == MODE_PRIORITY) bm = fmaxf(bm, fabsf(val[j]));
            }
        }
        if (mode == MODE_PRIORITY) {
            bm = warp_max(bm);
            if (lane_id() == 0 && bm > 0.0f) atomicMax(reinterpret_cast<int *>(&t.st->batch_max), __float_as_int(bm));
        }
        __syncthreads();
        const float defp = (mode == MODE_EXTEND) ? default_priority(t) : 0.0f;
        for (long long j = tid; j < n; j += blockDim.x) {
            const long long i = entry_index(t, idx, j, mode);
            if (idx_out) idx_out[j] = i;
            if (i < 0 || i >= t.size || t.owner[i] != (int)j) continue;
            const float v = entry_leaf(t, val, j, mode, defp);
            t.sum[t.cap + i] = v;
            t.min[t.cap + i] = v;
            t.owner[i] = -1;
        }
        __syncthreads();
    }
    for (int p = 0; p < P; ++p) {
        for (long long j = warp; j < n; j += (blockDim.x >> 5)) {
            if (p == 0 && fused) sparse_entry<true>(t, n, idx, val, mode, p, 1, idx_out, j);
            else sparse_entry<false>(t, n, idx, val, mode, p, fused ? 1 : 0, nullptr, j);
        }
        __syncthreads();
    }
    // top levels: threads [0, 512) own the sum tree, [512, 1024) the min tree
    const bool is_min = tid >= 512;
    const int gt = tid & 511;
    float *tree = is_min ? t.min : t.sum;
    float *s = sm + (is_min ? (2 << T) : 0);
    const int n_top = 1 << T;
    for (int i = gt; i < n_top; i += 512) s[n_top + i] = tree[n_top + i];
    __syncthreads();
    for (int d = T - 1; d >= 0; --d) {
        const int n_d = 1 << d;
        for (int i = gt; i < n_d; i += 512) {
            const float l = s[2 * (n_d + i)], r = s[2 * (n_d + i) + 1];
            s[n_d + i] = is_min ? op_min(l, r) : op_sum(l, r);
        }
        __syncthreads();
    }
    for (int i = gt + 1; i < n_top; i += 512) tree[i] = s[i];
    __syncthreads();
    if (tid >= 32) return;
    long long len = t.st->len;
    if (mode == MODE_EXTEND) len = len + n < t.size ? len + n : t.size;
    const float ps = tree_query_prefix<false>(t, len);
    const float pm = tree_query_prefix<true>(t, len);
    if (tid == 0) {
        pb_per_state *st = t.st;
        if (mode == MODE_EXTEND) st->seq += n;
        st->len = len;
        if (mode == MODE_PRIORITY) {
            const float bm = st->batch_max;
            if (bm > st->max_priority) st->max_priority = bm;
        }
        st->batch_max = 0.0f;
        st->p_sum = ps; st->p_min = pm;
    }
}

// ---------------------------------------------------------------------------------
// prefix-sum descent: one warp per sample, 5 levels per dependent 128-byte load.
// Returns the leaf index (or size when mass > root), uniform across the warp.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ long long warp_descend(const float *__restrict__ sum, int L, long long size,
                                                  float m)
{
    const int lane = lane_id();
    if (m > sum[1]) return size;
    long long node = 1;
    int rem = L;
    int c = rem % 5;          // short chunk first, so every deeper chunk is a full aligned line
    if (c == 0) c = 5;
    while (rem > 0) {
        const int cnt = 1 << c;
        const long long base = node << c;
        float s0 = lane < cnt ? sum[base + lane] : 0.0f;
        float s1 = op_sum(s0, __shfl_xor_sync(FULL, s0, 1));
        float s2 = op_sum(s1, __shfl_xor_sync(FULL, s1, 2));
        float s3 = op_sum(s2, __shfl_xor_sync(FULL, s2, 4));
        float s4 = op_sum(s3, __shfl_xor_sync(FULL, s3, 8));
        int pos = 0;
        if (c >= 5) { float l = __shfl_sync(FULL, s4, pos); if (m > l) { m = __fsub_rn(m, l); pos += 16; } }
        if (c >= 4) { float l = __shfl_sync(FULL, s3, pos); if (m > l) { m = __fsub_rn(m, l); pos += 8; } }
        if (c >= 3) { float l = __shfl_sync(FULL, s2, pos); if (m > l) { m = __fsub_rn(m, l); pos += 4; } }
        if (c >= 2) { float l = __shfl_sync(FULL, s1, pos); if (m > l) { m = __fsub_rn(m, l); pos += 2; } }
        {             float l = __shfl_sync(FULL, s0, pos); if (m > l) { m = __fsub_rn(m, l); pos += 1; } }
        node = base + pos;
        rem -= c;
        c = 5;
    }
    return node ^ (1LL << L);
}

// ---------------------------------------------------------------------------------
// throughput mode (large batches): one THREAD per sample, the top levels staged in shared
// memory.  Same comparisons in the same order as warp_descend (and as the reference loop), so
// the two modes return identical indices; with stratified (sorted) masses neighbouring threads
// walk neighbouring paths, so the per-level loads of a warp coalesce into a few sectors.
// ---------------------------------------------------------------------------------
constexpr int STAGE_LEVELS = 11;   // nodes [1, 2^11) = 8 KB of shared memory per CTA

__device__ __forceinline__ void stage_top(const float *__restrict__ sum, int L, float *sm)
{
    const int S = L < STAGE_LEVELS ? L : STAGE_LEVELS;
    const int n = 1 << S;
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = sum[i];
    __syncthreads();
}

__device__ __forceinline__ long long thread_descend(const float *__restrict__ sum, const float *sm, int L,
                                                    long long size, float m)
{
    if (m > sm[1]) return size;
    const int S = L < STAGE_LEVELS ? L : STAGE_LEVELS;
    const long long staged = 1LL << S;
    long long node = 1;
    for (int d = 0; d < L; ++d) {
        node <<= 1;
        const float left = node < staged ? sm[node] : __ldg(sum + node);
        if (m > left) { m = __fsub_rn(m, left); node |= 1; }
    }
    return node ^ (1LL << L);
}

__global__ void __launch_bounds__(256) tree_scan_thread_kernel(TreeView t, long long n, const float *mass,
                                                               long long *idx_out)
{
    __shared__ float sm[1 << STAGE_LEVELS];
    stage_top(t.sum, t.L, sm);
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    idx_out[k] = thread_descend(t.sum, sm, t.L, t.size, mass[k]);
}

__global__ void __launch_bounds__(256) tree_scan_kernel(TreeView t, long long n, const float *mass,
                                                        long long *idx_out)
{
    const long long k = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= n) return;
    long long i = warp_descend(t.sum, t.L, t.size, mass[k]);
    if (lane_id() == 0) idx_out[k] = i;
}

__device__ __forceinline__ float is_weight(float leaf, float p_min, float beta, const TreeView &t)
{
    float denom = t.weps ? __fadd_rn(p_min, t.eps32) : p_min;
    float ratio = __fdiv_rn(leaf, denom);
    // np.power(fp32, -beta) -> libm powf (< 1 ulp); evaluate in double and round once
    return (float)pow((double)ratio, -(double)beta);
}

// ---------------------------------------------------------------------------------
// device-side uniforms: Philox4x32-10 keyed by the shard seed, counter = (stratum, call number).
// Every rank of a sharded buffer draws the same u_k without communication; the call number lives in
// the state block and is advanced by the last CTA of the launch (ticket), so a replayed CUDA graph
// gets fresh numbers every iteration with no extra launch.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ double philox_uniform(unsigned seed, unsigned call, unsigned long long k)
{
    unsigned c0 = (unsigned)k, c1 = (unsigned)(k >> 32), c2 = call, c3 = 0x50455221u;
    unsigned k0 = seed, k1 = 0x9E3779B9u ^ seed;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    // 53 random bits -> [0, 1)
    return ((double)(c0 >> 5) * 67108864.0 + (double)(c1 >> 6)) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ void rng_advance(const TreeView &t)
{
    // call after a __syncthreads(): every thread of this CTA has read the call number
    if (threadIdx.x == 0) {
        __threadfence();
        const int ticket = atomicAdd(&t.st->pad[1], 1);
        if (ticket == (int)(gridDim.x * gridDim.y) - 1) {
            t.st->pad[2] += 1;
            t.st->pad[1] = 0;
        }
    }
}

__global__ void __launch_bounds__(256) tree_sample_kernel(TreeView t, long long n, const double *u, int mode,
                                                          float beta, long long *idx_out, float *w_out,
                                                          float *mass_out)
{
    const long long k = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = lane_id();
    const unsigned call = (unsigned)t.st->pad[2], seed = (unsigned)t.st->pad[3];
    if (k < n) {
        const long long len = t.st->len;
        const float p_sum = t.st->p_sum, p_min = t.st->p_min;
        int bad = 0;
        if (len <= 0) bad = PB_ST_EMPTY;
        else if (!(p_sum > 0.0f)) bad = PB_ST_PSUM_NONPOS;
        else if (!(p_min > 0.0f)) bad = PB_ST_PMIN_NONPOS;
        if (bad) {
            if (lane == 0) {
                idx_out[k] = 0; w_out[k] = 0.0f;
                if (mass_out) mass_out[k] = 0.0f;
                if (k == 0) atomicOr(&t.st->status, bad);
            }
        } else {
            const double uk = u ? u[k] : philox_uniform(seed, call, (unsigned long long)k);
            const double m64 = (mode == 0) ? (0.0 + ((double)p_sum - 0.0) * uk)
                                           : __dmul_rn(__ddiv_rn(__dadd_rn((double)k, uk), (double)n), (double)p_sum);
            const float m = (float)m64;
            long long i = warp_descend(t.sum, t.L, t.size, m);
            if (i > len - 1) i = len - 1;
            if (lane == 0) {
                float leaf = t.sum[t.cap + i];
                idx_out[k] = i;
                w_out[k] = is_weight(leaf, p_min, beta, t);
                if (mass_out) mass_out[k] = m;
            }
        }
    }
    if (!u) { __syncthreads(); rng_advance(t); }
}

__global__ void __launch_bounds__(256) tree_sample_thread_kernel(TreeView t, long long n, const double *u, int mode,
                                                                 float beta, long long *idx_out, float *w_out,
                                                                 float *mass_out)
{
    __shared__ float sm[1 << STAGE_LEVELS];
    const unsigned call = (unsigned)t.st->pad[2], seed = (unsigned)t.st->pad[3];
    stage_top(t.sum, t.L, sm);
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        const long long len = t.st->len;
        const float p_sum = t.st->p_sum, p_min = t.st->p_min;
        int bad = 0;
        if (len <= 0) bad = PB_ST_EMPTY;
        else if (!(p_sum > 0.0f)) bad = PB_ST_PSUM_NONPOS;
        else if (!(p_min > 0.0f)) bad = PB_ST_PMIN_NONPOS;
        if (bad) {
            idx_out[k] = 0; w_out[k] = 0.0f;
            if (mass_out) mass_out[k] = 0.0f;
            if (k == 0) atomicOr(&t.st->status, bad);
        } else {
            const double uk = u ? u[k] : philox_uniform(seed, call, (unsigned long long)k);
            const double m64 = (mode == 0) ? (0.0 + ((double)p_sum - 0.0) * uk)
                                           : __dmul_rn(__ddiv_rn(__dadd_rn((double)k, uk), (double)n), (double)p_sum);
            const float m = (float)m64;
            long long i = thread_descend(t.sum, sm, t.L, t.size, m);
            if (i > len - 1) i = len - 1;
            idx_out[k] = i;
            w_out[k] = is_weight(__ldg(t.sum + t.cap + i), p_min, beta, t);
            if (mass_out) mass_out[k] = m;
        }
    }
    if (!u) { __syncthreads(); rng_advance(t); }
}

// ---------------------------------------------------------------------------------
// sharded global stratified sampling: the G shard roots are the leaves of a virtual
// top tree, summed pairwise in fp32 (so shards concatenated == one big tree).
// ---------------------------------------------------------------------------------
constexpr int MAX_RANKS = 64;

struct GlobalTop {
    float psum[2 * MAX_RANKS];  // heap layout of the virtual top: leaves (shard p_sums) at [G, 2G)
    float pmin;
};

// every CTA rebuilds the virtual top from the all-gathered shard stats (device memory)
__device__ __forceinline__ void build_top(GlobalTop *g, const pb_per_state *all_state, int G)
{
    if (threadIdx.x < G) g->psum[G + threadIdx.x] = all_state[threadIdx.x].p_sum;
    __syncthreads();
    for (int n = G >> 1; n >= 1; n >>= 1) {
        if (threadIdx.x < n) g->psum[n + threadIdx.x] = op_sum(g->psum[2 * (n + threadIdx.x)], g->psum[2 * (n + threadIdx.x) + 1]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float m = all_state[0].p_min;
        for (int r = 1; r < G; ++r) m = op_min(m, all_state[r].p_min);
        g->pmin = m;
    }
    __syncthreads();
}

__device__ __forceinline__ int route_stratum(const GlobalTop &g, int G, long long k, long long n_global,
                                             double uk, float *residual)
{
    const float total = g.psum[1];
    const double m64 = __dmul_rn(__ddiv_rn(__dadd_rn((double)k, uk), (double)n_global), (double)total);
    float m = (float)m64;
    *residual = m;
    if (m > total) return G - 1;  // unreachable for u < 1; clamp like idx.clamp_max(len-1)
    int node = 1;
    while (node < G) {
        node <<= 1;
        float left = g.psum[node];
        if (m > left) { m = __fsub_rn(m, left); node |= 1; }
    }
    *residual = m;
    return node - G;
}

__global__ void __launch_bounds__(256) global_count_kernel(TreeView t, const pb_per_state *all_state,
                                                           int G, int rank, long long n_global, const double *u)
{
    __shared__ GlobalTop g;
    build_top(&g, all_state, G);
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int lo = 0, mine = 0;
    if (k < n_global) {
        float res;
        const double uk = u ? u[k] : philox_uniform((unsigned)t.st->pad[3], (unsigned)t.st->pad[2], (unsigned long long)k);
        int owner = route_stratum(g, G, k, n_global, uk, &res);
        lo = owner < rank;
        mine = owner == rank;
    }
    unsigned blo = __ballot_sync(FULL, lo), bmine = __ballot_sync(FULL, mine);
    if (lane_id() == 0) {
        if (blo) atomicAdd(&t.st->owned_lo, __popc(blo));
        if (bmine) atomicAdd(&t.st->owned_n, __popc(bmine));
    }
}

__global__ void __launch_bounds__(256) global_sample_kernel(TreeView t, const pb_per_state *all_state, int G, int rank,
                                                            long long n_global, const double *u, float beta,
                                                            long long *idx_out, float *w_out, long long *stratum_out)
{
    __shared__ GlobalTop g;
    build_top(&g, all_state, G);
    const long long k = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = lane_id();
    const unsigned call = (unsigned)t.st->pad[2], seed = (unsigned)t.st->pad[3];
    if (k < n_global) {
        const int lo = t.st->owned_lo, cnt = t.st->owned_n;
        if (k >= cnt && lane == 0) {  // padding rows of the static batch: idx -1 is skipped downstream
            idx_out[k] = -1; w_out[k] = 0.0f;
            if (stratum_out) stratum_out[k] = -1;
        }
        float res;
        const double uk = u ? u[k] : philox_uniform(seed, call, (unsigned long long)k);
        const int owner = route_stratum(g, G, k, n_global, uk, &res);
        if (owner == rank) {
            const long long len = all_state[rank].len;
            if (len <= 0 || !(g.psum[1] > 0.0f) || !(g.pmin > 0.0f)) {
                if (lane == 0 && k == lo)
                    atomicOr(&t.st->status, len <= 0 ? PB_ST_EMPTY : (!(g.psum[1] > 0.0f) ? PB_ST_PSUM_NONPOS : PB_ST_PMIN_NONPOS));
            }
            long long i = warp_descend(t.sum, t.L, t.size, res);
            if (i > len - 1) i = len - 1;
            if (i < 0) i = 0;
            if (lane == 0) {
                const long long pos = k - lo;
                float leaf = t.sum[t.cap + i];
                idx_out[pos] = i;
                w_out[pos] = is_weight(leaf, g.pmin, beta, t);
                if (stratum_out) stratum_out[pos] = k;
            }
        }
    }
    if (!u) { __syncthreads(); rng_advance(t); }
}

__global__ void global_reset_kernel(TreeView t) { t.st->owned_lo = 0; t.st->owned_n = 0; }

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
int make_view(const pb_tree *t, TreeView *v)
{
    if (!t || !t->sum || !t->min || !t->state) return PB_E_ARG;
    if (!pb_is_pow2(t->capacity) || t->size <= 0 || t->size > t->capacity) return PB_E_CAPACITY;
    if (t->capacity < 2 || t->capacity > (1LL << 30)) return PB_E_CAPACITY;
    v->sum = t->sum; v->min = t->min; v->owner = t->owner; v->st = t->state;
    v->cap = t->capacity; v->size = t->size; v->L = pb_ilog2(t->capacity);
    v->alpha = t->alpha; v->eps32 = t->eps_f32; v->eps64 = t->eps_f64;
    v->weps = t->weight_eps_in_denominator; v->dp64 = t->default_priority_fp64;
    return PB_OK;
}

int sparse_phases(int L) { return L <= TOP_MAX ? 0 : (L - TOP_MAX + 4) / 5; }

int launch_top(const TreeView &v, int T, int mode, long long n_new, long long set_len, void *stream)
{
    static PbPerDeviceOnce attr_set;
    size_t smem = sizeof(float) * (2ull << T);
    if (!attr_set.done()) {
        cudaError_t e = cudaFuncSetAttribute(tree_top_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(sizeof(float) * (2ull << TOP_MAX)));
        if (e != cudaSuccess) return (int)e;
        attr_set.mark();
    }
    PB_LAUNCH(tree_top_kernel, 2, 1024, smem, stream, v, T, mode, n_new, set_len);
    return PB_OK;
}

int launch_update(const pb_tree *t, long long n, const long long *idx, const float *val, int mode, int sorted,
                  long long *idx_out, void *stream)
{
    TreeView v;
    int rc = make_view(t, &v);
    if (rc) return rc;
    if (n < 0 || n >= (1LL << 31)) return PB_E_ARG;
    if (mode != MODE_EXTEND && (!idx || !val)) return n == 0 ? PB_OK : PB_E_ARG;
    if (n == 0) return PB_OK;
    if (mode == MODE_EXTEND) {
        if (n > v.size) return PB_E_ARG;
        sorted = (n + 64 <= v.size);  // contiguous run, except when it can wrap onto its own line
    }
    const int P = sparse_phases(v.L);
    // dense path: when the batch touches a sizeable part of the tree, scatter the leaves and rebuild
    // every level with the streaming build kernels (2 x 12 B x cap of traffic, independent of n)
    if (P > 0 && n * 128 >= v.cap && v.L >= 11) {
        const int nb = (int)((n + 255) / 256);
        if (sorted) {
            PB_LAUNCH(upd_leaf_sorted_kernel, nb, 256, 0, stream, v, n, idx, val, mode, idx_out);
        } else {
            if (!v.owner) return PB_E_ARG;
            PB_LAUNCH(upd_mark_kernel, nb, 256, 0, stream, v, n, idx, val, mode);
            PB_LAUNCH(upd_leaf_kernel, nb, 256, 0, stream, v, n, idx, val, mode, idx_out);
        }
        int d = v.L;
        while (d > TOP_MAX) {
            dim3 g2((unsigned)((1LL << d) / STAGE_TILE), 2);
            PB_LAUNCH(tree_reduce11_kernel<false>, g2, 256, 0, stream, v, d, (const float *)nullptr, 0LL);
            d -= 11;
        }
        return launch_top(v, d, mode, n, -1, stream);
    }
    if (n <= SMALL_MAX && v.L >= 5) {
        // latency mode: one single-CTA launch
        const int Ps = v.L <= SMALL_TOP ? 0 : (v.L - SMALL_TOP + 4) / 5;
        const int Ts = v.L - 5 * Ps;
        if (!(sorted && Ps > 0) && !v.owner) return PB_E_ARG;
        static PbPerDeviceOnce attr_set;
        if (!attr_set.done()) {
            cudaError_t e = cudaFuncSetAttribute(upd_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)(sizeof(float) * 2 * (2ull << SMALL_TOP)));
            if (e != cudaSuccess) return (int)e;
            attr_set.mark();
        }
        const size_t smem = sizeof(float) * 2 * (2ull << Ts);
        PB_LAUNCH(upd_small_kernel, 1, 1024, smem, stream, v, n, idx, val, mode, sorted, Ps, Ts, idx_out);
        return PB_OK;
    }
    const bool fused = sorted && P > 0;
    if (!fused) {
        if (!v.owner) return PB_E_ARG;
        const int nb = (int)((n + 255) / 256);
        PB_LAUNCH(upd_mark_kernel, nb, 256, 0, stream, v, n, idx, val, mode);
        PB_LAUNCH(upd_leaf_kernel, nb, 256, 0, stream, v, n, idx, val, mode, idx_out);
    }
    const int nbw = (int)((n + 7) / 8);
    for (int p = 0; p < P; ++p) {
        if (p == 0 && fused)
            PB_LAUNCH(upd_sparse_kernel<true>, nbw, 256, 0, stream, v, n, idx, val, mode, p, 1, idx_out);
        else
            PB_LAUNCH(upd_sparse_kernel<false>, nbw, 256, 0, stream, v, n, idx, val, mode, p, fused ? 1 : 0,
                      (long long *)nullptr);
    }
    return launch_top(v, v.L - 5 * P, mode, n, -1, stream);
}

}  // namespace

extern "C" {

int pb_tree_init(const pb_tree *t, void *stream)
{
    TreeView v;
    int rc = make_view(t, &v);
    if (rc) return rc;
    int nb = pb_sm_count() * 8;
    PB_LAUNCH(tree_init_kernel, nb, 256, 0, stream, v);
    return PB_OK;
}

int pb_tree_stats(const pb_tree *t, void *stream)
{
    TreeView v;
    int rc = make_view(t, &v);
    if (rc) return rc;
    PB_LAUNCH(tree_stats_kernel, 1, 32, 0, stream, v);
    return PB_OK;
}

int pb_tree_build(const pb_tree *t, const float *leaves, long long n, void *stream)
{
    TreeView v;
    int rc = make_view(t, &v);
    if (rc) return rc;
    if (!leaves || n < 0 || n > v.size) return PB_E_ARG;
    int d = v.L;
    if (d > TOP_MAX) {
        // stage 1 from the leaves, then further 11-level stages while the frontier is too deep
        if (d < 11) return PB_E_CAPACITY;
        dim3 grid((unsigned)(v.cap / STAGE_TILE), 2);
        PB_LAUNCH(tree_reduce11_kernel<true>, grid, 256, 0, stream, v, d, leaves, n);
        d -= 11;
        while (d > TOP_MAX) {
            dim3 g2((unsigned)((1LL << d) / STAGE_TILE), 2);
            PB_LAUNCH(tree_reduce11_kernel<false>, g2, 256, 0, stream, v, d, (const float *)nullptr, 0LL);
            d -= 11;
        }
    } else {
        int nb = (int)((v.cap + 255) / 256);
        PB_LAUNCH(tree_fill_leaves_kernel, nb, 256, 0, stream, v, leaves, n);
    }
    return launch_top(v, d, MODE_RAW, 0, n, stream);
}

int pb_tree_set_leaves(const pb_tree *t, long long n, const long long *idx, const float *leaves, int sorted,
                       void *stream)
{
    return launch_update(t, n, idx, leaves, MODE_RAW, sorted, nullptr, stream);
}

int pb_tree_update_priority(const pb_tree *t, long long n, const long long *idx, const float *priority,
                            int sorted, void *stream)
{
    return launch_update(t, n, idx, priority, MODE_PRIORITY, sorted, nullptr, stream);
}

int pb_tree_extend(const pb_tree *t, long long n, long long *idx_out, void *stream)
{
    return launch_update(t, n, nullptr, nullptr, MODE_EXTEND, 1, idx_out, stream);
}

int pb_tree_scan(const pb_tree *t, long long n, const float *mass, long long *idx_out, void *stream)
{
    TreeView v;
    int rc = make_view(t, &v);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!mass || !idx_out))) return PB_E_ARG;
    if (n == 0) return PB_OK;
    if (n >= THREAD_MODE_MIN)
        PB_LAUNCH(tree_scan_thread_kernel, (unsigned)((n + 255) / 256), 256, 0, stream, v, n, mass, idx_out);
    else
        PB_LAUNCH(tree_scan_kernel, (unsigned)((n + 7) / 8), 256, 0, stream, v, n, mass, idx_out);
    return PB_OK;
}

int pb_tree_sample(const pb_tree *t, long long n, const double *u, int mode, float beta, long long *idx_out,
                   float *weight_out, float *mass_out, void *stream)
{
    TreeView v;
    int rc = make_view(t, &v);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!idx_out || !weight_out)) || (mode != 0 && mode != 1)) return PB_E_ARG;
    if (n == 0) return PB_OK;
    if (n >= THREAD_MODE_MIN)
        PB_LAUNCH(tree_sample_thread_kernel, (unsigned)((n + 255) / 256), 256, 0, stream, v, n, u, mode, beta,
                  idx_out, weight_out, mass_out);
    else
        PB_LAUNCH(tree_sample_kernel, (unsigned)((n + 7) / 8), 256, 0, stream, v, n, u, mode, beta, idx_out,
                  weight_out, mass_out);
    return PB_OK;
}

int pb_tree_sample_global(const pb_tree *t, int n_ranks, int rank, const pb_per_state *all_state,
                          long long n_global, const double *u, float beta, long long *idx_out,
                          float *weight_out, long long *stratum_out, void *stream)
{
    TreeView v;
    int rc = make_view(t, &v);
    if (rc) return rc;
    if (n_ranks < 1 || n_ranks > MAX_RANKS || !pb_is_pow2(n_ranks) || rank < 0 || rank >= n_ranks) return PB_E_ARG;
    if (!all_state || n_global < 0) return PB_E_ARG;
    if (n_global > 0 && (!idx_out || !weight_out)) return PB_E_ARG;
    PB_LAUNCH(global_reset_kernel, 1, 1, 0, stream, v);
    if (n_global == 0) return PB_OK;
    PB_LAUNCH(global_count_kernel, (unsigned)((n_global + 255) / 256), 256, 0, stream, v, all_state,
              n_ranks, rank, n_global, u);
    PB_LAUNCH(global_sample_kernel, (unsigned)((n_global + 7) / 8), 256, 0, stream, v, all_state,
              n_ranks, rank, n_global, u, beta, idx_out, weight_out, stratum_out);
    return PB_OK;
}

}  // extern "C"
