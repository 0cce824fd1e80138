// per_tree.cu -- sum-tree / min-tree priority store for prioritized replay (sm_100a).
//
// Replaces the segment tree + PrioritizedSampler that the reference reaches through
// torchrl (prism/factory/exp_buffer_factory.py:22-28; call sites
// prism/experience/timestep_buffer.py:33,37,54; prism/learner.py:100,120).
//
// Design (B200-first, not a port of a pointer-walking CPU tree):
//  * every node is fl32(left + right) / min(left, right) of a binary tree over the leaves -- the reference add
//    order, so sampled indices are bit-exact whatever order updates arrive in.
//  * one primitive everywhere: 32 sibling nodes (one aligned 128-byte line) are reduced over 5 levels in registers
//    and shuffles.  The rebuilt intermediate values are bit-identical to the nodes a pointer-walking tree would
//    store, so they are NOT stored: the tree keeps only every 5th level below its top ("compact" layout)
//        levels 0 .. TL          one level-ordered heap array (TL <= 9: <= 4 KB)
//        levels TL+5, TL+10, .., L   one array per level; L = leaves
//    and the min tree shares the leaf array of the sum tree (a leaf is written with the same value in both;
//    slots that were never written -- index >= len -- read as +inf on the min side).
//    16M leaves: 64 MiB + 2 x 2.1 MiB instead of 2 x 128 MiB, i.e. the whole store fits the 126 MB L2.
//      - sampling descends 5 levels per dependent 128-byte load: 8 lanes per sample for one learner batch (latency),
//        a thread per sample over a shared-memory copy of the top 14 levels for many batches in flight (throughput);
//      - a sorted priority update is ONE launch: the leader of every touched line applies its leaves, reduces the
//        line and climbs; lines of the next stored level are finished by whichever child arrives last (arrival
//        counters, no spinning), and the last CTA (ticket) rebuilds the top heap and the state block;
//      - any other batch (unsorted, duplicates, K batches in flight) is three launches: a mark pass (dedup tag on the
//        leaf slot + one bit per touched leaf line), the leaf scatter, and a sparse rebuild whose CTAs own 32768-leaf
//        spans (lane per touched leaf line) and climb three stored levels without any cross-CTA dependency;
//      - batches beyond cap/16 entries scatter their leaves and rebuild every line with one streaming pass over the
//        leaf array (4.3 B per leaf of traffic, independent of the batch size);
//      - the bulk build is that same streaming pass.
//  * duplicates in an update batch: last occurrence wins (sequential reference loop), resolved deterministically
//    (adjacency when sorted; otherwise atomicMax of a NaN-tagged entry number on the leaf slot itself -- no side array).
//  * no host sync anywhere: len / cursor / max_priority / p_sum / p_min live in a 64-byte device state block, so
//    the whole sample->update loop is CUDA-graph capturable.
#include "common.cuh"
#include <math.h>

namespace {

using namespace pb;

constexpr int TOP_MAX = 9;         // the top heap holds levels [0, TL], TL <= 9: rebuilt whole by ONE CTA after every update
constexpr int MAX_DEEP = 8;        // stored levels below the heap (capacity <= 2^30 -> at most 4)

struct TreeView {
    float *sum, *min;
    int *owner, *cnt;
    pb_per_state *st;
    long long cap, size;
    int L, TL, P;                  // P = (L - TL) / 5 deep levels: L, L-5, .., TL+5
    long long off[MAX_DEEP];       // off[m]: float offset of level TL + 5m in the sum / min store (m >= 1)
    long long coff[MAX_DEEP];      // coff[m]: int offset of the arrival counters of the lines at level TL + 5m (1 <= m < P)
    unsigned *bitmap;              // one bit per leaf line (cap / 32 bits), all zero between calls: lines an update touched
    float alpha, eps32;
    double eps64;
    int weps, dp64;
};

enum { MODE_RAW = 0, MODE_PRIORITY = 1, MODE_EXTEND = 2 };
constexpr float INF = __builtin_huge_valf();

__device__ __forceinline__ float op_sum(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float op_min(float a, float b) { return fminf(a, b); }

// level s of the sum tree (node i of that level at [i])
__device__ __forceinline__ float *sum_level(const TreeView &t, int s)
{
    return s <= t.TL ? t.sum + (1LL << s) : t.sum + t.off[(s - t.TL) / 5];
}
// level s < L of the min tree (the leaf level is shared with the sum tree, see min_of_leaf)
__device__ __forceinline__ float *min_level(const TreeView &t, int s)
{
    return s <= t.TL ? t.min + (1LL << s) : t.min + t.off[(s - t.TL) / 5];
}
__device__ __forceinline__ float *leaf_ptr(const TreeView &t) { return sum_level(t, t.L); }
__device__ __forceinline__ float min_of_leaf(float v, long long i, long long len) { return i < len ? v : INF; }

__device__ __forceinline__ float ldcg(const float *p) { return __ldcg(p); }

// arrival-counter update with release (my node store is visible to whoever reads the counter after me) and acquire
// (I see the node stores of everyone who updated it before me) semantics in ONE instruction -- no MEMBAR.SC
__device__ __forceinline__ int atom_add_acq_rel(int *p, int v)
{
    int old;
    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}

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

__device__ __forceinline__ long long entry_index(const TreeView &t, const long long *idx, long long j, int mode,
                                                 long long seq0)
{
    if (mode == MODE_EXTEND) return (seq0 + j) % t.size;
    return idx[j];
}

__device__ __forceinline__ float entry_leaf(const TreeView &t, const float *val, long long j, int mode, float defp)
{
    if (mode == MODE_RAW) return val[j];
    if (mode == MODE_PRIORITY) return pow_leaf(fabsf(val[j]), t);   // learner.py:120 passes |td|; idempotent
    return defp;
}

__device__ __forceinline__ long long len_after(const TreeView &t, int mode, long long n_new, long long set_len)
{
    long long len = t.st->len;
    if (mode == MODE_EXTEND) len = len + n_new < t.size ? len + n_new : t.size;
    else if (set_len >= 0) len = set_len < t.size ? set_len : t.size;
    return len;
}

// reduce one line: every lane holds one of 32 sibling nodes; returns the 5-level pairwise total in every lane
__device__ __forceinline__ void line_reduce(float &vs, float &vm)
{
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        vs = op_sum(vs, __shfl_xor_sync(FULL, vs, 1 << s));
        vm = op_min(vm, __shfl_xor_sync(FULL, vm, 1 << s));
    }
}

// ---------------------------------------------------------------------------------
// value of ANY node (level d, index i), stored or not: an unstored level is rebuilt pairwise from the next
// stored level below it (<= 16 nodes).
// ---------------------------------------------------------------------------------
template <bool IS_MIN>
__device__ float node_value(const TreeView &t, int d, long long i, long long len)
{
    const float ident = IS_MIN ? INF : 0.0f;
    const int r = d <= t.TL ? 0 : (t.L - d) % 5;     // levels to the next stored level below
    const int s = d + r;
    const long long base = i << r;
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        if (k < (1 << r)) {
            // L2 loads: the state block is finalised by the last CTA of a launch whose other CTAs wrote these nodes
            if (s == t.L) {
                float x = ldcg(leaf_ptr(t) + base + k);
                v[k] = IS_MIN ? min_of_leaf(x, base + k, len) : x;
            } else {
                v[k] = IS_MIN ? ldcg(min_level(t, s) + base + k) : ldcg(sum_level(t, s) + base + k);
            }
        } else {
            v[k] = ident;
        }
    }
#pragma unroll
    for (int w = 8; w >= 1; w >>= 1) {
        if (w < (1 << r)) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < w) v[k] = IS_MIN ? op_min(v[2 * k], v[2 * k + 1]) : op_sum(v[2 * k], v[2 * k + 1]);
        }
    }
    return v[0];
}

// ---------------------------------------------------------------------------------
// query(0, len) with torchrl's interval-walk association (one warp)
// ---------------------------------------------------------------------------------
template <bool IS_MIN>
__device__ float tree_query_prefix(const TreeView &t, long long len)
{
    const float ident = IS_MIN ? INF : 0.0f;
    if (len >= t.size) return IS_MIN ? ldcg(t.min + 1) : ldcg(t.sum + 1);
    if (len <= 0) return ident;
    const int k = lane_id();
    float v = ident;
    // bit k of len set: the walk takes node (len >> k) - 1 of level L - k (a node fully inside [0, len))
    if (k < t.L && ((len >> k) & 1)) v = node_value<IS_MIN>(t, t.L - k, (len >> k) - 1, len);
    float ret = ident;
    for (int b = 0; b < t.L; ++b) {
        float vb = __shfl_sync(FULL, v, b);
        if ((len >> b) & 1) ret = IS_MIN ? op_min(ret, vb) : op_sum(ret, vb);
    }
    return ret;
}

// one warp: advance len / seq, merge max_priority, refresh p_sum / p_min
__device__ void finalize_state(const TreeView &t, int mode, long long n_new, long long set_len)
{
    const long long len = len_after(t, mode, n_new, set_len);
    const float ps = tree_query_prefix<false>(t, len);
    const float pm = tree_query_prefix<true>(t, len);
    if (lane_id() == 0) {
        pb_per_state *s = t.st;
        if (mode == MODE_EXTEND) s->seq += n_new;
        else if (set_len >= 0) s->seq = set_len;
        s->len = len;
        if (mode == MODE_PRIORITY) {
            const float bm = s->batch_max;
            if (bm > s->max_priority) s->max_priority = bm;
        }
        s->batch_max = 0.0f;
        s->p_sum = ps; s->p_min = pm;
        s->pad[0] = 0;
    }
}

// ---------------------------------------------------------------------------------
// top phase: rebuild heap levels TL-1 .. 0 of BOTH trees from level TL (32 .. 512 nodes).  Called by every thread of ONE
// CTA (blockDim.x threads, a multiple of 64; the first half works on the sum tree, the second on the min tree).
// Every warp reduces lines of 32 level-TL nodes with shuffles, storing the intermediate levels as it goes; the
// <= 16 line totals (level TL-5) meet in shared memory and one warp per tree finishes the last levels.
// ---------------------------------------------------------------------------------
constexpr int TOP_SM_FLOATS = 32;                   // per tree: the line totals

__device__ void top_rebuild(const TreeView &t, float *sm, long long len)
{
    const int half = blockDim.x >> 1;
    const bool is_min = threadIdx.x >= half;
    const int tid = is_min ? threadIdx.x - half : threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5, n_warps = half >> 5;
    float *heap = is_min ? t.min : t.sum;
    float *s = sm + (is_min ? TOP_SM_FLOATS : 0);
    const int TL = t.TL;                                        // 5 .. 9
    const bool shared_leaves = is_min && TL == t.L;             // tiny trees: level TL is the shared leaf level
    const float *src = shared_leaves ? t.sum + (1 << TL) : heap + (1 << TL);
    const int n_lines = 1 << (TL - 5);
    for (int line = warp; line < n_lines; line += n_warps) {
        const int e = (line << 5) + lane;
        float v = ldcg(src + e);
        if (shared_leaves) v = min_of_leaf(v, e, len);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const float o = __shfl_xor_sync(FULL, v, 1 << k);
            v = is_min ? op_min(v, o) : op_sum(v, o);
            if ((lane & ((2 << k) - 1)) == 0) heap[(1 << (TL - 1 - k)) + (e >> (k + 1))] = v;
        }
        if (lane == 0) s[line] = v;                              // node `line` of level TL - 5 (already stored above)
    }
    __syncthreads();
    if (warp == 0 && TL > 5) {
        float v = lane < n_lines ? s[lane] : (is_min ? INF : 0.0f);
        for (int k = 0; k < TL - 5; ++k) {
            const float o = __shfl_xor_sync(FULL, v, 1 << k);
            v = is_min ? op_min(v, o) : op_sum(v, o);
            if (lane < n_lines && (lane & ((2 << k) - 1)) == 0) heap[(1 << (TL - 6 - k)) + (lane >> (k + 1))] = v;
        }
    }
}

// last-CTA ticket: returns true in every thread of the CTA that arrives last.  pad[0] is reset by finalize_state.
__device__ bool last_cta(const TreeView &t)
{
    __shared__ int is_last;
    __syncthreads();                                  // the CTA's stores happen-before thread 0's release below
    if (threadIdx.x == 0) {
        const int ticket = atom_add_acq_rel(&t.st->pad[0], 1);
        is_last = (ticket == (int)gridDim.x - 1);
    }
    __syncthreads();
    return is_last != 0;
}

__device__ void top_and_finalize(const TreeView &t, float *sm, int mode, long long n_new, long long set_len)
{
    const long long len = len_after(t, mode, n_new, set_len);
    top_rebuild(t, sm, len);
    __threadfence();
    __syncthreads();
    if (threadIdx.x < 32) finalize_state(t, mode, n_new, set_len);
}

// ---------------------------------------------------------------------------------
// init
// ---------------------------------------------------------------------------------
__global__ void tree_init_kernel(TreeView t, long long n_sum, long long n_min, long long n_cnt)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = i; k < n_sum; k += stride) t.sum[k] = 0.0f;
    for (long long k = i; k < n_min; k += stride) t.min[k] = INF;
    for (long long k = i; k < n_cnt; k += stride) t.cnt[k] = 0;
    if (i == 0) {
        pb_per_state s;
        s.len = 0; s.seq = 0; s.max_priority = 1.0f; s.p_sum = 0.0f; s.p_min = INF; s.status = 0;
        s.batch_max = 0.0f; s.owned_lo = 0; s.owned_n = 0;
        for (int k = 0; k < 5; ++k) s.pad[k] = 0;
        *t.st = s;
    }
}

__global__ void tree_stats_kernel(TreeView t)
{
    const long long len = t.st->len;
    const float ps = tree_query_prefix<false>(t, len);
    const float pm = tree_query_prefix<true>(t, len);
    if (threadIdx.x == 0) { t.st->p_sum = ps; t.st->p_min = pm; }
}

// ---------------------------------------------------------------------------------
// streaming rebuild: every line of level s is reduced to its node of level s-5 and, when that level is stored
// too, on to level s-10.  Warp-autonomous, no shared memory: per iteration a warp streams 1024 consecutive source
// nodes with 8 coalesced float4 loads per lane in flight (8 lanes per 128-byte line), reduces every line in
// registers + 3 shuffle levels, transposes the 32 line totals into one lane each with shuffles (= one coalesced
// 128-byte store of level s-5) and reduces those 5 more levels for the node of level s-10.  Persistent grid.
// s == L reads the shared leaf array (optionally first loading it from `ext`, the bulk build).
// FUSE_TOP: the last CTA (ticket) rebuilds the top heap and the state block.
// ---------------------------------------------------------------------------------
constexpr int RB_THREADS = 256;
constexpr int RB_TILE = 1024;                                    // source nodes per warp iteration

// n_lv = 1, 2 or 3 levels per pass (s-5, s-10, s-15).  The third level needs 32 warp tiles in one CTA: a CTA then takes
// 32768 consecutive source nodes per iteration (4 warp tiles per warp) and combines their level s-10 nodes through
// shared memory.
__global__ void __launch_bounds__(RB_THREADS) tree_rebuild_kernel(TreeView t, int s, int n_lv, const float *ext,
                                                                  long long n_ext, int fuse_top, int mode,
                                                                  long long n_new, long long set_len)
{
    __shared__ float sm_top[2 * TOP_SM_FLOATS];
    __shared__ float l2s[32], l2m[32];
    const int two_levels = n_lv >= 2;
    const long long len = len_after(t, mode, n_new, set_len);
    const bool leaves = (s == t.L);
    float *src_s = sum_level(t, s);
    const float *src_m = leaves ? src_s : min_level(t, s);
    float *dst1_s = sum_level(t, s - 5), *dst1_m = min_level(t, s - 5);
    const long long n_tiles = (1LL << s) / RB_TILE;
    const int lane = lane_id();
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    // n_lv < 3: warps stride over the warp tiles on their own.  n_lv == 3: CTAs stride over groups of 32 warp tiles.
    const int wic = threadIdx.x >> 5;                            // warp in CTA (RB_THREADS / 32 = 8 of them)
    const long long n_outer = n_lv == 3 ? n_tiles / 32 : 1;
    for (long long outer = n_lv == 3 ? blockIdx.x : 0; outer < n_outer; outer += n_lv == 3 ? gridDim.x : 1) {
    const long long t_first = n_lv == 3 ? outer * 32 + wic : (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const long long t_step = n_lv == 3 ? RB_THREADS / 32 : n_warps;
    const long long t_end = n_lv == 3 ? outer * 32 + 32 : n_tiles;
    for (long long tile = t_first; tile < t_end; tile += t_step) {
        const long long base = tile * RB_TILE + lane * 4;        // load q covers nodes base + 128 q .. + 3
        float4 a[8], m[8];
        if (ext) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const long long e = base + 128 * q;
                if (e + 4 <= n_ext) a[q] = __ldcs(reinterpret_cast<const float4 *>(ext + e));
                else {
                    a[q].x = e + 0 < n_ext ? ext[e + 0] : 0.0f; a[q].y = e + 1 < n_ext ? ext[e + 1] : 0.0f;
                    a[q].z = e + 2 < n_ext ? ext[e + 2] : 0.0f; a[q].w = e + 3 < n_ext ? ext[e + 3] : 0.0f;
                }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) *reinterpret_cast<float4 *>(src_s + base + 128 * q) = a[q];
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) a[q] = __ldcg(reinterpret_cast<const float4 *>(src_s + base + 128 * q));
        }
        if (leaves) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const long long e = base + 128 * q;
                m[q] = a[q];
                if (e + 4 > len) {
                    m[q].x = min_of_leaf(a[q].x, e + 0, len); m[q].y = min_of_leaf(a[q].y, e + 1, len);
                    m[q].z = min_of_leaf(a[q].z, e + 2, len); m[q].w = min_of_leaf(a[q].w, e + 3, len);
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) m[q] = __ldcg(reinterpret_cast<const float4 *>(src_m + base + 128 * q));
        }
        // line totals: load q holds 4 lines (8 lanes each); afterwards lane l owns line l of the tile
        float ps = 0.0f, pm = 0.0f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float vs = op_sum(op_sum(a[q].x, a[q].y), op_sum(a[q].z, a[q].w));
            float vm = op_min(op_min(m[q].x, m[q].y), op_min(m[q].z, m[q].w));
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                vs = op_sum(vs, __shfl_xor_sync(FULL, vs, 1 << k));
                vm = op_min(vm, __shfl_xor_sync(FULL, vm, 1 << k));
            }
            // line 4q + g sits in the lanes of group g: hand it to lane 4q + g
            const float gs = __shfl_sync(FULL, vs, (lane & 3) << 3);
            const float gm = __shfl_sync(FULL, vm, (lane & 3) << 3);
            if ((lane >> 2) == q) { ps = gs; pm = gm; }
        }
        dst1_s[tile * 32 + lane] = ps;                           // 32 nodes of level s-5: one coalesced line
        dst1_m[tile * 32 + lane] = pm;
        if (two_levels) {
            line_reduce(ps, pm);
            if (lane == 0) {
                sum_level(t, s - 10)[tile] = ps; min_level(t, s - 10)[tile] = pm;
                if (n_lv == 3) { l2s[tile & 31] = ps; l2m[tile & 31] = pm; }
            }
        }
    }
    if (n_lv == 3) {
        __syncthreads();
        if (wic == 0) {
            float ps = l2s[lane], pm = l2m[lane];
            line_reduce(ps, pm);
            if (lane == 0) { sum_level(t, s - 15)[outer] = ps; min_level(t, s - 15)[outer] = pm; }
        }
        __syncthreads();
    }
    }
    if (!fuse_top) return;
    if (!last_cta(t)) return;
    top_and_finalize(t, sm_top, mode, n_new, set_len);
}

// trees without deep levels (L <= 14): copy the leaves (identity padded) into the heap's leaf level
__global__ void tree_fill_leaves_kernel(TreeView t, const float *leaves, long long n_leaves)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t.cap) return;
    leaf_ptr(t)[i] = i < n_leaves ? leaves[i] : 0.0f;
}

// standalone top phase (one CTA)
__global__ void __launch_bounds__(512) tree_top_kernel(TreeView t, int mode, long long n_new, long long set_len)
{
    __shared__ float sm_top[2 * TOP_SM_FLOATS];
    top_and_finalize(t, sm_top, mode, n_new, set_len);
}

// ---------------------------------------------------------------------------------
// leaf scatter, general (unsorted) path.  Duplicates: the LAST occurrence wins (sequential reference loop).  No side
// array: the leaf slot itself is the scratch.  A) every entry j does atomicMax(int view of leaf[i], TAG + j): TAG + j
// is a NaN bit pattern, as a signed int larger than every non-NaN float's, so the slot ends up holding the tag of the
// highest j that targets it.  B) the entry whose tag it finds there writes the value.  Dense path, sorted input: the
// last entry of every run of equal indices writes the leaf.
// ---------------------------------------------------------------------------------
constexpr int LEAF_TAG = 0x7F800001;                 // first NaN pattern; TAG + j stays a positive int for j < 2^23 - 1
constexpr long long LEAF_TAG_MAX_N = (1LL << 23) - 2;

__global__ void upd_mark_kernel(TreeView t, long long n, const long long *idx, const float *val, int mode,
                                long long *idx_out, int set_bits)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long seq0 = t.st->seq;
    float bm = 0.0f;
    if (j < n) {
        const long long i = entry_index(t, idx, j, mode, seq0);
        if (idx_out) idx_out[j] = i;
        if (i >= 0 && i < t.size) {
            atomicMax(reinterpret_cast<int *>(leaf_ptr(t) + i), LEAF_TAG + (int)j);
            if (set_bits) atomicOr(t.bitmap + (i >> 10), 1u << ((i >> 5) & 31));      // leaf line i >> 5 is touched
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
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const long long i = entry_index(t, idx, j, mode, t.st->seq);
    if (idx_out) idx_out[j] = i;
    if (i < 0 || i >= t.size) return;
    if (__ldcg(reinterpret_cast<const int *>(leaf_ptr(t) + i)) != LEAF_TAG + (int)j) return;
    const float defp = (mode == MODE_EXTEND) ? default_priority(t) : 0.0f;
    leaf_ptr(t)[i] = entry_leaf(t, val, j, mode, defp);
}

__global__ void upd_leaf_sorted_kernel(TreeView t, long long n, const long long *idx, const float *val, int mode,
                                       long long *idx_out)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long seq0 = t.st->seq;
    float bm = 0.0f;
    if (j < n) {
        const long long i = entry_index(t, idx, j, mode, seq0);
        if (idx_out) idx_out[j] = i;
        if (i >= 0 && i < t.size) {
            if (mode == MODE_PRIORITY) bm = fabsf(val[j]);
            const long long nx = (j + 1 < n) ? entry_index(t, idx, j + 1, mode, seq0) : -1;
            if (nx != i) {
                const float defp = (mode == MODE_EXTEND) ? default_priority(t) : 0.0f;
                leaf_ptr(t)[i] = entry_leaf(t, val, j, mode, defp);
            }
        }
    }
    if (mode == MODE_PRIORITY) {
        bm = warp_max(bm);
        if (lane_id() == 0 && bm > 0.0f) atomicMax(reinterpret_cast<int *>(&t.st->batch_max), __float_as_int(bm));
    }
}

// ---------------------------------------------------------------------------------
// sparse rebuild (the last launch of a general update, after upd_mark_kernel + upd_leaf_kernel).  A CTA owns a span of
// 32768 leaves = 32 words of the touched-line bitmap = one line of level L-10; a warp takes 2 of those words.  For a
// word, LANE l owns leaf line 32 w + l: if its bit is set it streams the line (8 x 128-bit loads, all in flight),
// reduces it in registers in tree order and stores the node of level L-5; otherwise it just loads the node that is
// there.  The warp then holds the whole line of level L-5 in its lanes: 5 shuffle levels give the node of level L-10.
// Finally warp 0 reduces the CTA's line of level L-10 to the node of level L-15.  No cross-CTA dependency below the
// top heap; ~130 warp instructions per bitmap word whatever its population.  The bitmap words are cleared on the
// way.  The last CTA (ticket) rebuilds the top heap and the state block when level L-15 is (or lies above) the
// heap's bottom level.
// ---------------------------------------------------------------------------------
constexpr int SPR_THREADS = 512;                 // 16 warps, 2 bitmap words (64 leaf lines) each

__global__ void __launch_bounds__(SPR_THREADS) tree_rebuild_sparse_kernel(TreeView t, int mode, long long n_new,
                                                                          int fuse_top)
{
    __shared__ float sm_top[2 * TOP_SM_FLOATS];
    const int lane = lane_id(), wic = threadIdx.x >> 5;
    constexpr int WPW = 32 / (SPR_THREADS / 32);                 // bitmap words per warp
    const long long len = len_after(t, mode, n_new, -1);
    const float *leaf = leaf_ptr(t);
    float *n1s = sum_level(t, t.L - 5), *n1m = min_level(t, t.L - 5);
    const long long n_words = t.cap >= 1024 ? t.cap >> 10 : 1;
    const int P = t.P;
    for (long long span = blockIdx.x; span * 32 < n_words; span += gridDim.x) {
        int touched_any = 0;
#pragma unroll
        for (int q = 0; q < WPW; ++q) {
            const long long w = span * 32 + wic * WPW + q;
            unsigned bits = 0;
            if (w < n_words && lane == 0) { bits = t.bitmap[w]; if (bits) t.bitmap[w] = 0u; }
            bits = __shfl_sync(FULL, bits, 0);
            if (!bits) continue;                                  // warp-uniform
            touched_any = 1;
            const long long line = (w << 5) + lane;
            float vs, vm;
            if ((bits >> lane) & 1u) {
                const float4 *src = reinterpret_cast<const float4 *>(leaf + (line << 5));
                float4 v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = __ldcg(src + k);
                float ps[8], pm[8];
                const bool whole = (line << 5) + 32 <= len;       // every leaf of the line is a filled slot
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float4 m = v[k];
                    if (!whole) {
                        const long long e = (line << 5) + 4 * k;
                        m.x = min_of_leaf(m.x, e + 0, len); m.y = min_of_leaf(m.y, e + 1, len);
                        m.z = min_of_leaf(m.z, e + 2, len); m.w = min_of_leaf(m.w, e + 3, len);
                    }
                    ps[k] = op_sum(op_sum(v[k].x, v[k].y), op_sum(v[k].z, v[k].w));
                    pm[k] = op_min(op_min(m.x, m.y), op_min(m.z, m.w));
                }
                vs = op_sum(op_sum(op_sum(ps[0], ps[1]), op_sum(ps[2], ps[3])), op_sum(op_sum(ps[4], ps[5]), op_sum(ps[6], ps[7])));
                vm = op_min(op_min(op_min(pm[0], pm[1]), op_min(pm[2], pm[3])), op_min(op_min(pm[4], pm[5]), op_min(pm[6], pm[7])));
                n1s[line] = vs; n1m[line] = vm;
            } else {
                vs = ldcg(n1s + line); vm = ldcg(n1m + line);
            }
            if (P >= 2) {
                line_reduce(vs, vm);                              // the lanes hold the word's line of level L-5
                if (lane == 0) { sum_level(t, t.L - 10)[w] = vs; min_level(t, t.L - 10)[w] = vm; }
            }
        }
        if (P >= 3) {
            const int any = __syncthreads_or(touched_any);        // also orders the warps' level L-10 stores
            if (any && wic == 0) {
                float vs = ldcg(sum_level(t, t.L - 10) + (span << 5) + lane);
                float vm = ldcg(min_level(t, t.L - 10) + (span << 5) + lane);
                line_reduce(vs, vm);
                if (lane == 0) { sum_level(t, t.L - 15)[span] = vs; min_level(t, t.L - 15)[span] = vm; }
            }
        }
    }
    if (!fuse_top) return;
    if (!last_cta(t)) return;
    top_and_finalize(t, sm_top, mode, n_new, -1);
}

// ---------------------------------------------------------------------------------
// sorted input, ONE launch: warp per batch entry (grid-stride).  Entries are grouped by the 1024-leaf SPAN they fall
// into (one line of level L-5; 32 leaf lines).  The first entry of a span's run is its leader: it writes the run's
// leaves (last of equal indices wins), reduces the touched leaf lines (several loads in flight) to their nodes of
// level L-5, then -- being the only writer of that line -- reduces the span's line of level L-5 to the node of
// level L-10.  Lines of the levels above are finished by whichever child arrives last: the first entry of a line's
// run registers the number of touched children in the line's arrival counter (+c), every finished child subtracts
// one (acq_rel), and the warp whose atomic brings the counter back to zero reduces the line and climbs on (the
// counter is zero again for the next call; nobody spins).  The last CTA (ticket) rebuilds the top heap and the
// state block.  Trees with a single deep level (P == 1) use the leaf line itself as the group.
// ---------------------------------------------------------------------------------
constexpr int CHAIN_THREADS = 512;
constexpr long long CHAIN_MAX_N = 8192;      // beyond this the per-warp serial chain loses to mark + sparse rebuild
constexpr int CHAIN_U = 4;                   // touched leaf lines in flight per leader

struct RunScan { long long seq0; const long long *idx; long long n; int mode; };

__device__ __forceinline__ bool valid_at(const TreeView &t, const RunScan &r, long long j, long long &i)
{
    if (j < 0 || j >= r.n) return false;
    i = entry_index(t, r.idx, j, r.mode, r.seq0);
    return i >= 0 && i < t.size;
}

// number of distinct values of (index >> child_shift) among the run of entries starting at j whose
// (index >> shift) equals key (warp-uniform result)
__device__ int count_children(const TreeView &t, const RunScan &r, long long j, int shift, int child_shift,
                              long long key)
{
    const int lane = lane_id();
    int c = 0;
    for (long long base = j;; base += 32) {
        const long long jj = base + lane;
        long long ii = -1, ip = -1;
        const bool ok = valid_at(t, r, jj, ii) && (ii >> shift) == key;
        const unsigned bad = ~__ballot_sync(FULL, ok);
        const int run = bad ? __ffs(bad) - 1 : 32;
        bool fresh = false;
        if (lane < run) fresh = (jj == j) || !valid_at(t, r, jj - 1, ip) || (ip >> child_shift) != (ii >> child_shift);
        c += __popc(__ballot_sync(FULL, fresh));
        if (run < 32) break;
    }
    return c;
}

__global__ void __launch_bounds__(CHAIN_THREADS) upd_chain_kernel(TreeView t, long long n, const long long *idx,
                                                                  const float *val, int mode, long long *idx_out)
{
    __shared__ float sm_top[2 * TOP_SM_FLOATS];
    const int lane = lane_id();
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long w0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    RunScan r;
    r.seq0 = t.st->seq; r.idx = idx; r.n = n; r.mode = mode;
    const long long len = len_after(t, mode, n, -1);
    const float defp = (mode == MODE_EXTEND) ? default_priority(t) : 0.0f;
    float *leaf = leaf_ptr(t);
    const int P = t.P;
    const int gs = P >= 2 ? 10 : 5;                               // group = span of 2^gs leaves
    for (long long j = w0; j < n; j += n_warps) {
        long long i, ip = -1;
        if (!valid_at(t, r, j, i)) continue;
        const bool prev_ok = valid_at(t, r, j - 1, ip);
        if (P == 0) {
            // no deep level: the heap's leaf level is written directly, the top phase does the rest
            float bm = 0.0f;
            long long inx = -1;
            const bool last_of_idx = !valid_at(t, r, j + 1, inx) || inx != i;
            if (lane == 0) {
                if (idx_out) idx_out[j] = i;
                if (last_of_idx) leaf[i] = entry_leaf(t, val, j, mode, defp);
                if (mode == MODE_PRIORITY) bm = fabsf(val[j]);
                if (bm > 0.0f) atomicMax(reinterpret_cast<int *>(&t.st->batch_max), __float_as_int(bm));
            }
            continue;
        }
        const long long grp = i >> gs;
        if (prev_ok && (ip >> gs) == grp) continue;               // the group's leader does this span
        // ---- registration: lines of the levels above the group whose run starts at this entry.  The group's node
        // lives at level L - gs = TL + 5 mg; the lines above are those of levels TL + 5m, m = mg .. 1.
        const int mg = P - gs / 5;
        for (int m = mg; m >= 1; --m) {
            const int shift = 5 * (P - m) + 5;                    // index >> shift = line index at level TL + 5m
            const long long pl = i >> shift;
            if (prev_ok && (ip >> shift) == pl) break;            // not the first entry of that line (nor of any above)
            const int c = count_children(t, r, j, shift, shift - 5, pl);
            if (lane == 0) atomicAdd(&t.cnt[t.coff[m] + pl], c);
        }
        // ---- the group's run: write its leaves, 32 entries at a time; collect the touched leaf lines
        float bm = 0.0f;
        unsigned linebits = 0;
        for (long long base = j;; base += 32) {
            const long long jj = base + lane;
            long long ii = -1, inx = -1;
            const bool ok = valid_at(t, r, jj, ii) && (ii >> gs) == grp;
            const unsigned bad = ~__ballot_sync(FULL, ok);
            const int run = bad ? __ffs(bad) - 1 : 32;
            unsigned mine = 0;
            if (lane < run) {
                if (idx_out) idx_out[jj] = ii;
                if (mode == MODE_PRIORITY) bm = fmaxf(bm, fabsf(val[jj]));
                if (!valid_at(t, r, jj + 1, inx) || inx != ii) leaf[ii] = entry_leaf(t, val, jj, mode, defp);
                mine = 1u << ((ii >> 5) & (gs == 10 ? 31 : 0));
            }
            linebits |= __reduce_or_sync(FULL, mine);
            if (run < 32) break;
        }
        if (mode == MODE_PRIORITY) {
            bm = warp_max(bm);
            if (lane == 0 && bm > 0.0f) atomicMax(reinterpret_cast<int *>(&t.st->batch_max), __float_as_int(bm));
        }
        __syncwarp();
        // ---- touched leaf lines -> nodes of level L-5, CHAIN_U lines per round
        float *n1s = sum_level(t, t.L - 5), *n1m = min_level(t, t.L - 5);
        const long long line0 = gs == 10 ? grp << 5 : grp;        // first leaf line of the group
        float vs = 0.0f, vm = 0.0f;
        while (linebits) {
            long long line[CHAIN_U];
            float x[CHAIN_U];
#pragma unroll
            for (int u = 0; u < CHAIN_U; ++u) {
                line[u] = -1;
                if (linebits) { line[u] = line0 + (__ffs(linebits) - 1); linebits &= linebits - 1; }
            }
#pragma unroll
            for (int u = 0; u < CHAIN_U; ++u) x[u] = line[u] >= 0 ? ldcg(leaf + (line[u] << 5) + lane) : 0.0f;
#pragma unroll
            for (int u = 0; u < CHAIN_U; ++u) {
                if (line[u] < 0) continue;                        // warp-uniform
                vs = x[u]; vm = min_of_leaf(vs, (line[u] << 5) + lane, len);
                line_reduce(vs, vm);
                if (lane == 0) { n1s[line[u]] = vs; n1m[line[u]] = vm; }
            }
        }
        long long g = grp;                                        // node index at the group's level
        if (gs == 10) {
            // the span's line of level L-5: this warp is its only writer in this launch
            __syncwarp();
            vs = ldcg(n1s + (grp << 5) + lane);
            vm = ldcg(n1m + (grp << 5) + lane);
            line_reduce(vs, vm);
            if (lane == 0) { sum_level(t, t.L - 10)[grp] = vs; min_level(t, t.L - 10)[grp] = vm; }
        }
        // ---- climb: lines of level TL + 5m, m = mg .. 1
        for (int m = mg; m >= 1; --m) {
            const long long pl = g >> 5;
            int newv = 0;
            if (lane == 0) newv = atom_add_acq_rel(&t.cnt[t.coff[m] + pl], -1) - 1;     // releases lane 0's node store
            newv = __shfl_sync(FULL, newv, 0);
            if (newv != 0) break;                                 // another child of this line is still on its way
            __syncwarp();                                         // lane 0's acquire orders the whole warp's loads below
            const int s = t.TL + 5 * m;
            vs = ldcg(sum_level(t, s) + (pl << 5) + lane);
            vm = ldcg(min_level(t, s) + (pl << 5) + lane);
            line_reduce(vs, vm);
            if (lane == 0) { sum_level(t, s - 5)[pl] = vs; min_level(t, s - 5)[pl] = vm; }
            g = pl;
        }
    }
    if (!last_cta(t)) return;
    top_and_finalize(t, sm_top, mode, n, -1);
}

// ---------------------------------------------------------------------------------
// prefix-sum descent: 8 lanes per sample (4 samples per warp), 5 levels per dependent 128-byte load: every lane
// loads 4 of the line's 32 nodes (one float4), two levels are reduced in registers and three with shuffles.  Same
// comparisons in the same order as the reference loop (go right and subtract iff mass > left).  All 32 lanes of the
// warp must call this together.  Returns the leaf index (or size when mass > root) and the leaf's value.
// ---------------------------------------------------------------------------------
template <int U>
__device__ __forceinline__ void group_descend_multi(const TreeView &t, float (&m)[U], long long (&node)[U],
                                                    float (&leafv)[U])
{
    // U independent descents per 8-lane group, interleaved level by level: U 128-byte loads in flight per group
    const int gl = threadIdx.x & 7;
    const float root = t.sum[1];
    bool over[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        over[u] = m[u] > root;
        if (over[u]) m[u] = 0.0f;                                 // stay on real nodes; the result is discarded
        node[u] = 0;                                              // index within the current level
        leafv[u] = 0.0f;
    }
    int d = 0;
    int c = t.L % 5;                                              // short chunk first: deeper chunks are full aligned lines
    if (c == 0) c = 5;
    while (d < t.L) {
        const int s = d + c;
        const float *lvl = sum_level(t, s);
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float *p = lvl + (node[u] << c);
            if (c == 5) {
                v[u] = *reinterpret_cast<const float4 *>(p + 4 * gl);
            } else {
                const int cnt = 1 << c, b = 4 * gl;
                v[u].x = b + 0 < cnt ? p[b + 0] : 0.0f; v[u].y = b + 1 < cnt ? p[b + 1] : 0.0f;
                v[u].z = b + 2 < cnt ? p[b + 2] : 0.0f; v[u].w = b + 3 < cnt ? p[b + 3] : 0.0f;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float a = op_sum(v[u].x, v[u].y), b2 = op_sum(v[u].z, v[u].w);
            const float c4 = op_sum(a, b2);
            const float s1 = op_sum(c4, __shfl_xor_sync(FULL, c4, 1, 8));
            const float s2 = op_sum(s1, __shfl_xor_sync(FULL, s1, 2, 8));
            float mm = m[u];
            int pos = 0;
            float l = __shfl_sync(FULL, s2, 0, 8);
            if (mm > l) { mm = __fsub_rn(mm, l); pos = 4; }
            l = __shfl_sync(FULL, s1, pos, 8);
            if (mm > l) { mm = __fsub_rn(mm, l); pos += 2; }
            l = __shfl_sync(FULL, c4, pos, 8);
            if (mm > l) { mm = __fsub_rn(mm, l); pos += 1; }
            const float la = __shfl_sync(FULL, a, pos, 8);
            const float lx = __shfl_sync(FULL, v[u].x, pos, 8), ly = __shfl_sync(FULL, v[u].y, pos, 8);
            const float lz = __shfl_sync(FULL, v[u].z, pos, 8), lw = __shfl_sync(FULL, v[u].w, pos, 8);
            int sub = 0;
            float l0 = lx, l1 = ly;
            if (mm > la) { mm = __fsub_rn(mm, la); sub = 2; l0 = lz; l1 = lw; }
            float lf = l0;
            if (mm > l0) { mm = __fsub_rn(mm, l0); sub += 1; lf = l1; }
            m[u] = mm;
            leafv[u] = lf;
            node[u] = (c == 5 ? (node[u] << 5) : 0) + 4 * pos + sub;
        }
        d = s;
        c = 5;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) if (over[u]) node[u] = t.size;
}

__device__ __forceinline__ long long group_descend(const TreeView &t, float m, float *leaf_out)
{
    float mm[1] = {m}, lf[1];
    long long nd[1];
    group_descend_multi<1>(t, mm, nd, lf);
    *leaf_out = lf[0];
    return nd[0];
}

// R sub-rounds of 4 descents each (one per 8-lane group): lane l = 4 sub + g hands its mass to group g in sub-round
// `sub` and collects the result.  R = 8 runs the sub-rounds four at a time (interleaved descents).
__device__ __forceinline__ void warp_descend_rounds(const TreeView &t, float m, int R, long long &mine, float &my_leaf)
{
    const int lane = lane_id();
    mine = 0; my_leaf = 0.0f;
    if (R == 8) {
#pragma unroll 1
        for (int sub = 0; sub < 8; sub += 4) {
            float mg[4], lf[4];
            long long nd[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) mg[u] = __shfl_sync(FULL, m, 4 * (sub + u) + (lane >> 3));
            group_descend_multi<4>(t, mg, nd, lf);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long gi = __shfl_sync(FULL, nd[u], (lane & 3) << 3);
                const float gf = __shfl_sync(FULL, lf[u], (lane & 3) << 3);
                if ((lane >> 2) == sub + u) { mine = gi; my_leaf = gf; }
            }
        }
    } else {
#pragma unroll 1
        for (int sub = 0; sub < R; ++sub) {
            const float mg = __shfl_sync(FULL, m, 4 * sub + (lane >> 3));
            float leafv;
            const long long i = group_descend(t, mg, &leafv);
            const long long gi = __shfl_sync(FULL, i, (lane & 3) << 3);
            const float gf = __shfl_sync(FULL, leafv, (lane & 3) << 3);
            if ((lane >> 2) == sub) { mine = gi; my_leaf = gf; }
        }
    }
}

__global__ void __launch_bounds__(256) tree_scan_kernel(TreeView t, long long n, const float *mass, long long *idx_out)
{
    const int lane = lane_id();
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long w0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long rounds = (n + 32 * n_warps - 1) / (32 * n_warps);
    for (long long rd = 0; rd < rounds; ++rd) {
        const long long k = (w0 + rd * n_warps) * 32 + lane;
        const float m = k < n ? mass[k] : 0.0f;
        long long mine;
        float my_leaf;
        warp_descend_rounds(t, m, 8, mine, my_leaf);
        if (k < n) idx_out[k] = mine;
    }
}

__device__ __forceinline__ float is_weight(float leaf, float p_min, float beta, const TreeView &t)
{
    float denom = t.weps ? __fadd_rn(p_min, t.eps32) : p_min;
    float ratio = __fdiv_rn(leaf, denom);
    // np.power(fp32, -beta) -> libm powf (< 1 ulp); evaluate in double and round once
    return (float)pow((double)ratio, -(double)beta);
}

// ---------------------------------------------------------------------------------
// device-side uniforms: Philox4x32-10 keyed by the shard seed, counter = (sample number, call number).
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

__device__ __forceinline__ float stratified_mass(long long k, double uk, long long n, float total)
{
    return (float)__dmul_rn(__ddiv_rn(__dadd_rn((double)k, uk), (double)n), (double)total);
}

// n = n_batches * batch samples; sample k belongs to batch k / batch, stratum k % batch of its batch (every batch
// is stratified on its own, all against the same tree state).  Persistent grid.  A warp takes 32 consecutive samples
// per round: every lane prepares ONE sample (uniform, fp64 mass), the descents run 4 at a time (8 lanes each, 8
// sub-rounds, masses and results handed over by shuffles), then every lane finishes its own sample (clamp, fp64
// importance weight) and the 32 results leave as coalesced stores.
// R sub-rounds per warp round (4 R samples per warp): 8 for large n (throughput), 1 for small batches (a warp then
// runs ONE descent chain, the latency of a single learner batch).
__global__ void __launch_bounds__(256) tree_sample_kernel(TreeView t, long long n, long long batch, const double *u,
                                                          int mode, float beta, long long *idx_out, float *w_out,
                                                          float *mass_out, int R)
{
    const unsigned call = (unsigned)t.st->pad[2], seed = (unsigned)t.st->pad[3];
    const long long len = t.st->len;
    const float p_sum = t.st->p_sum, p_min = t.st->p_min;
    int bad = 0;
    if (len <= 0) bad = PB_ST_EMPTY;
    else if (!(p_sum > 0.0f)) bad = PB_ST_PSUM_NONPOS;
    else if (!(p_min > 0.0f)) bad = PB_ST_PMIN_NONPOS;
    const int lane = lane_id();
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long w0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int per_warp = 4 * R;
    const long long rounds = (n + per_warp * n_warps - 1) / (per_warp * n_warps);
    for (long long rd = 0; rd < rounds; ++rd) {
        const long long k = (w0 + rd * n_warps) * per_warp + lane;
        const bool live = lane < per_warp && k < n;
        if (bad) {
            if (live) {
                idx_out[k] = 0; w_out[k] = 0.0f;
                if (mass_out) mass_out[k] = 0.0f;
                if (k == 0) atomicOr(&t.st->status, bad);
            }
            continue;
        }
        float m = 0.0f;
        if (live) {
            const double uk = u ? u[k] : philox_uniform(seed, call, (unsigned long long)k);
            m = (mode == 0) ? (float)(0.0 + ((double)p_sum - 0.0) * uk) : stratified_mass(k % batch, uk, batch, p_sum);
        }
        long long mine;
        float my_leaf;
        warp_descend_rounds(t, m, R, mine, my_leaf);
        if (live) {
            if (mine > len - 1) { mine = len - 1; my_leaf = leaf_ptr(t)[mine]; }
            idx_out[k] = mine;
            w_out[k] = is_weight(my_leaf, p_min, beta, t);
            if (mass_out) mass_out[k] = m;
        }
    }
    if (!u) { __syncthreads(); rng_advance(t); }
}

// ---------------------------------------------------------------------------------
// throughput mode (many batches in flight): one THREAD per sample.  The CTA stages the top of the tree as a full
// level-ordered heap in shared memory -- heap levels 0..TL copied, the stored level S <= 14 copied and the levels in
// between rebuilt pairwise (bit-identical to a stored tree) -- so the first S levels cost one shared-memory load
// each; every deeper stored level costs the thread one 128-byte line (8 x 128-bit loads in flight) whose 5 levels
// it resolves in registers.  Same comparisons in the same order as the reference loop and as group_descend, so
// both modes return identical indices.  ~10x fewer warp instructions per sample than the 8-lane descent.
// ---------------------------------------------------------------------------------
constexpr int ST_THREADS = 512;

// one line (32 nodes of a stored level, children of `node`): go right and subtract iff mass > left, 5 times
__device__ __forceinline__ void thread_line_descend(const float *__restrict__ lvl, long long &node, float &m, float &leaf)
{
    const float4 *p = reinterpret_cast<const float4 *>(lvl + (node << 5));
    float v[32];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float4 q = __ldg(p + k);
        v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
    }
    float t16[16], t8[8], t4[4], t2[2];
#pragma unroll
    for (int i = 0; i < 16; ++i) t16[i] = op_sum(v[2 * i], v[2 * i + 1]);
#pragma unroll
    for (int i = 0; i < 8; ++i) t8[i] = op_sum(t16[2 * i], t16[2 * i + 1]);
#pragma unroll
    for (int i = 0; i < 4; ++i) t4[i] = op_sum(t8[2 * i], t8[2 * i + 1]);
#pragma unroll
    for (int i = 0; i < 2; ++i) t2[i] = op_sum(t4[2 * i], t4[2 * i + 1]);
    int pos = 0;
    float l = t2[0];
    if (m > l) { m = __fsub_rn(m, l); pos = 1; }
    l = pos ? t4[2] : t4[0];
    pos <<= 1;
    if (m > l) { m = __fsub_rn(m, l); pos |= 1; }
    {   // left child of node `pos` at the 4-node level: t8[2 pos]
        const float a = (pos & 1) ? t8[2] : t8[0], b = (pos & 1) ? t8[6] : t8[4];
        l = (pos & 2) ? b : a;
    }
    pos <<= 1;
    if (m > l) { m = __fsub_rn(m, l); pos |= 1; }
    {   // t16[2 pos], pos in 0..7
        const float a0 = (pos & 1) ? t16[2] : t16[0], a1 = (pos & 1) ? t16[6] : t16[4];
        const float a2 = (pos & 1) ? t16[10] : t16[8], a3 = (pos & 1) ? t16[14] : t16[12];
        const float b0 = (pos & 2) ? a1 : a0, b1 = (pos & 2) ? a3 : a2;
        l = (pos & 4) ? b1 : b0;
    }
    pos <<= 1;
    if (m > l) { m = __fsub_rn(m, l); pos |= 1; }
    float lo, hi;                                                 // v[2 pos], v[2 pos + 1], pos in 0..15
    {
        float e[8], o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { e[i] = (pos & 1) ? v[4 * i + 2] : v[4 * i]; o[i] = (pos & 1) ? v[4 * i + 3] : v[4 * i + 1]; }
        float e2[4], o2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { e2[i] = (pos & 2) ? e[2 * i + 1] : e[2 * i]; o2[i] = (pos & 2) ? o[2 * i + 1] : o[2 * i]; }
        const float e3a = (pos & 4) ? e2[1] : e2[0], e3b = (pos & 4) ? e2[3] : e2[2];
        const float o3a = (pos & 4) ? o2[1] : o2[0], o3b = (pos & 4) ? o2[3] : o2[2];
        lo = (pos & 8) ? e3b : e3a;
        hi = (pos & 8) ? o3b : o3a;
    }
    pos <<= 1;
    leaf = lo;
    if (m > lo) { m = __fsub_rn(m, lo); pos |= 1; leaf = hi; }
    node = (node << 5) + pos;
}

__global__ void __launch_bounds__(ST_THREADS) tree_sample_thread_kernel(TreeView t, long long n, long long batch,
                                                                        const double *u, int mode, float beta,
                                                                        long long *idx_out, float *w_out,
                                                                        float *mass_out, int S)
{
    extern __shared__ float sh[];                                 // heap of levels 0..S: node i of level d at [2^d + i]
    const unsigned call = (unsigned)t.st->pad[2], seed = (unsigned)t.st->pad[3];
    const int TL = t.TL;
    for (int i = threadIdx.x; i < (2 << TL); i += blockDim.x) sh[i] = i ? t.sum[i] : 0.0f;
    if (S > TL) {
        const float4 *lv = reinterpret_cast<const float4 *>(sum_level(t, S));
        float4 *dst = reinterpret_cast<float4 *>(sh + (1 << S));
        for (int i = threadIdx.x; i < (1 << (S - 2)); i += blockDim.x) dst[i] = lv[i];
        for (int d = S - 1; d > TL; --d) {
            __syncthreads();
            for (int i = threadIdx.x; i < (1 << d); i += blockDim.x)
                sh[(1 << d) + i] = op_sum(sh[2 * ((1 << d) + i)], sh[2 * ((1 << d) + i) + 1]);
        }
    }
    __syncthreads();
    const long long len = t.st->len;
    const float p_sum = t.st->p_sum, p_min = t.st->p_min;
    int bad = 0;
    if (len <= 0) bad = PB_ST_EMPTY;
    else if (!(p_sum > 0.0f)) bad = PB_ST_PSUM_NONPOS;
    else if (!(p_min > 0.0f)) bad = PB_ST_PMIN_NONPOS;
    const float root = sh[1];
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        if (bad) {
            idx_out[k] = 0; w_out[k] = 0.0f;
            if (mass_out) mass_out[k] = 0.0f;
            if (k == 0) atomicOr(&t.st->status, bad);
            continue;
        }
        const double uk = u ? u[k] : philox_uniform(seed, call, (unsigned long long)k);
        const float m0 = (mode == 0) ? (float)(0.0 + ((double)p_sum - 0.0) * uk) : stratified_mass(k % batch, uk, batch, p_sum);
        float m = m0;
        const bool over = m > root;
        if (over) m = 0.0f;
        int hn = 1;                                               // heap index in shared memory
        for (int d = 0; d < S; ++d) {
            hn <<= 1;
            const float left = sh[hn];
            if (m > left) { m = __fsub_rn(m, left); hn |= 1; }
        }
        long long node = hn - (1 << S);                           // index within level S
        float leafv = sh[hn];
        for (int s = S + 5; s <= t.L; s += 5) thread_line_descend(sum_level(t, s), node, m, leafv);
        long long i = over ? t.size : node;
        if (i > len - 1) { i = len - 1; leafv = leaf_ptr(t)[i]; }
        idx_out[k] = i;
        w_out[k] = is_weight(leafv, p_min, beta, t);
        if (mass_out) mass_out[k] = m0;
    }
    if (!u) { __syncthreads(); rng_advance(t); }
}

// ---------------------------------------------------------------------------------
// sharded global stratified sampling: the G shard roots are the leaves of a virtual
// top tree, summed pairwise in fp32 (so shards concatenated == one big tree).  ONE launch: masses are
// non-decreasing in the stratum number and the descent is monotone, so the strata a rank owns are a contiguous
// range [lo, lo + cnt) that every CTA finds by a warp-parallel 32-ary search (no counting pass, no atomics).
// ---------------------------------------------------------------------------------
constexpr int MAX_RANKS = 64;

struct GlobalTop {
    float psum[2 * MAX_RANKS];  // heap layout of the virtual top: leaves (shard p_sums) at [G, 2G)
    float pmin;
    long long lo, hi;           // this rank owns strata [lo, hi)
};

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
    float m = stratified_mass(k, uk, n_global, total);
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

// first stratum k in [0, n] whose owner is >= rank_bound (n when there is none); one warp, 32-ary search over the
// monotone predicate owner(k) >= rank_bound
__device__ long long first_owned(const GlobalTop &g, int G, long long n, int rank_bound, const double *u,
                                 unsigned seed, unsigned call)
{
    const int lane = lane_id();
    long long lo = 0, hi = n;                                     // the answer lies in [lo, hi]
    while (lo < hi) {
        const long long step = (hi - lo + 31) / 32;
        const long long k = lo + (long long)lane * step;          // 32 probes, lo first
        bool ge = true;                                           // probes at or past hi count as "owned"
        if (k < hi) {
            float res;
            const double uk = u ? u[k] : philox_uniform(seed, call, (unsigned long long)k);
            ge = route_stratum(g, G, k, n, uk, &res) >= rank_bound;
        }
        const unsigned m = __ballot_sync(FULL, ge);
        if (m == 0u) { lo = lo + 31 * step + 1; continue; }       // all 32 probes below hi and none owned yet
        const int first = __ffs(m) - 1;
        if (first == 0) { hi = lo; break; }
        const long long k_first = lo + (long long)first * step;
        lo = k_first - step + 1;                                  // probe first-1 was not owned
        hi = k_first < hi ? k_first : hi;
    }
    return hi;
}

__global__ void __launch_bounds__(256) global_sample_kernel(TreeView t, const pb_per_state *all_state, int G, int rank,
                                                            long long n_global, const double *u, float beta,
                                                            long long *idx_out, float *w_out, long long *stratum_out,
                                                            int R)
{
    __shared__ GlobalTop g;
    build_top(&g, all_state, G);
    const unsigned call = (unsigned)t.st->pad[2], seed = (unsigned)t.st->pad[3];
    if (threadIdx.x < 32) {
        const long long lo = first_owned(g, G, n_global, rank, u, seed, call);
        const long long hi = first_owned(g, G, n_global, rank + 1, u, seed, call);
        if (threadIdx.x == 0) { g.lo = lo; g.hi = hi; }
    }
    __syncthreads();
    const long long lo = g.lo, cnt = g.hi - g.lo;
    const long long len = all_state[rank].len;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        t.st->owned_lo = (int)lo; t.st->owned_n = (int)cnt;
        if (cnt > 0 && (len <= 0 || !(g.psum[1] > 0.0f) || !(g.pmin > 0.0f)))
            atomicOr(&t.st->status, len <= 0 ? PB_ST_EMPTY : (!(g.psum[1] > 0.0f) ? PB_ST_PSUM_NONPOS : PB_ST_PMIN_NONPOS));
    }
    const int lane = lane_id();
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long w0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int per_warp = 4 * R;
    const long long rounds = (n_global + per_warp * n_warps - 1) / (per_warp * n_warps);
    for (long long rd = 0; rd < rounds; ++rd) {
        const long long pos = (w0 + rd * n_warps) * per_warp + lane;    // output row
        const bool mine_row = lane < per_warp;
        const bool live = mine_row && pos < cnt;
        float res = 0.0f;
        if (live) {
            const long long k = lo + pos;
            const double uk = u ? u[k] : philox_uniform(seed, call, (unsigned long long)k);
            route_stratum(g, G, k, n_global, uk, &res);
        }
        long long mine;
        float my_leaf;
        warp_descend_rounds(t, res, R, mine, my_leaf);
        if (mine_row && pos < n_global) {
            if (live) {
                if (mine > len - 1) { mine = len - 1; if (mine < 0) mine = 0; my_leaf = leaf_ptr(t)[mine]; }
                idx_out[pos] = mine;
                w_out[pos] = is_weight(my_leaf, g.pmin, beta, t);
                if (stratum_out) stratum_out[pos] = lo + pos;
            } else {                                              // padding rows of the static batch: skipped downstream
                idx_out[pos] = -1; w_out[pos] = 0.0f;
                if (stratum_out) stratum_out[pos] = -1;
            }
        }
    }
    if (!u) { __syncthreads(); rng_advance(t); }
}

// ---------------------------------------------------------------------------------
// export: the full level-ordered arrays (2 * capacity floats per tree) a pointer-walking tree would hold --
// parity tests and checkpoints.  One launch per level.
// ---------------------------------------------------------------------------------
__global__ void tree_export_kernel(TreeView t, int d, float *sum_heap, float *min_heap)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1LL << d)) return;
    const long long len = t.st->len;
    if (sum_heap) sum_heap[(1LL << d) + i] = node_value<false>(t, d, i, len);
    if (min_heap) min_heap[(1LL << d) + i] = node_value<true>(t, d, i, len);
    if (d == 0 && i == 0) {
        if (sum_heap) sum_heap[0] = 0.0f;
        if (min_heap) min_heap[0] = INF;
    }
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
struct Layout { int L, TL, P; long long off[MAX_DEEP], coff[MAX_DEEP]; long long n_sum, n_min, n_cnt, bitmap_off; };

void make_layout(long long cap, Layout *y)
{
    y->L = pb_ilog2(cap);
    y->P = y->L <= TOP_MAX ? 0 : (y->L - TOP_MAX + 4) / 5;
    y->TL = y->L - 5 * y->P;
    long long off = 2LL << y->TL, coff = 0;
    for (int m = 0; m < MAX_DEEP; ++m) { y->off[m] = 0; y->coff[m] = 0; }
    y->n_min = off;
    for (int m = 1; m <= y->P; ++m) {
        y->off[m] = off;
        off += 1LL << (y->TL + 5 * m);
        if (m < y->P) {
            y->n_min = off;                                       // the min store ends before the leaf level
            y->coff[m] = coff;
            coff += 1LL << (y->TL + 5 * m - 5);                   // one counter per line of level TL + 5m
        }
    }
    y->n_sum = off;
    y->bitmap_off = coff;                                         // touched-line bitmap behind the arrival counters
    y->n_cnt = coff + (cap >= 1024 ? cap / 1024 : 1);
}

int make_view(const pb_tree *t, TreeView *v)
{
    if (!t || !t->sum || !t->min || !t->state) return PB_E_ARG;
    if (!pb_is_pow2(t->capacity) || t->size <= 0 || t->size > t->capacity) return PB_E_CAPACITY;
    if (t->capacity < 32 || t->capacity > (1LL << 30)) return PB_E_CAPACITY;     // at least one 32-leaf line
    Layout y;
    make_layout(t->capacity, &y);
    if (!t->counters) return PB_E_ARG;
    v->sum = t->sum; v->min = t->min; v->owner = t->owner; v->cnt = t->counters; v->st = t->state;
    v->bitmap = reinterpret_cast<unsigned *>(t->counters + y.bitmap_off);
    v->cap = t->capacity; v->size = t->size; v->L = y.L; v->TL = y.TL; v->P = y.P;
    for (int m = 0; m < MAX_DEEP; ++m) { v->off[m] = y.off[m]; v->coff[m] = y.coff[m]; }
    v->alpha = t->alpha; v->eps32 = t->eps_f32; v->eps64 = t->eps_f64;
    v->weps = t->weight_eps_in_denominator; v->dp64 = t->default_priority_fp64;
    return PB_OK;
}

int persistent_grid(long long work_items, int per_cta, int ctas_per_sm)
{
    long long need = (work_items + per_cta - 1) / per_cta;
    long long cap = (long long)pb_sm_count() * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// every line of every deep level from the leaves up, then the top heap + state (streaming; ext: bulk build source)
int launch_rebuild(const TreeView &v, const float *ext, long long n_ext, int mode, long long n_new, long long set_len,
                   void *stream)
{
    if (v.P == 0) {
        if (ext) PB_LAUNCH(tree_fill_leaves_kernel, (unsigned)((v.cap + 255) / 256), 256, 0, stream, v, ext, n_ext);
        PB_LAUNCH(tree_top_kernel, 1, 512, 0, stream, v, mode, n_new, set_len);
        return PB_OK;
    }
    for (int s = v.L; s > v.TL;) {
        int n_lv = (s - v.TL) / 5;
        if (n_lv > 3) n_lv = 3;
        const int fuse = (s - 5 * n_lv <= v.TL) ? 1 : 0;
        const int grid = n_lv == 3 ? persistent_grid(1LL << s, 32 * RB_TILE, 4)
                                   : persistent_grid(1LL << s, RB_TILE * (RB_THREADS / 32), 8);
        PB_LAUNCH(tree_rebuild_kernel, grid, RB_THREADS, 0, stream, v, s, n_lv, s == v.L ? ext : (const float *)nullptr,
                  n_ext, fuse, mode, n_new, set_len);
        s -= 5 * n_lv;
    }
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
        sorted = (2 * n <= v.size);     // a contiguous run of slots; a long one may wrap onto lines it already touched
    }
    // Three regimes.  (1) small sorted batches (one learner batch, the ring's extends): ONE launch, leaders + arrival
    // counters.  (2) anything else up to cap/16 entries: mark (dedup tags + touched-line bitmap), leaf scatter, sparse
    // rebuild -- cost proportional to the batch.  (3) beyond that: scatter the leaves and stream the whole leaf array.
    if (v.P == 0) {
        if (sorted) {
            const int grid = persistent_grid(n, CHAIN_THREADS / 32, 2);
            PB_LAUNCH(upd_chain_kernel, grid, CHAIN_THREADS, 0, stream, v, n, idx, val, mode, idx_out);
            return PB_OK;
        }
        if (n > LEAF_TAG_MAX_N) return PB_E_UNSUPPORTED;
        const int nb0 = (int)((n + 255) / 256);
        PB_LAUNCH(upd_mark_kernel, nb0, 256, 0, stream, v, n, idx, val, mode, idx_out, 0);
        PB_LAUNCH(upd_leaf_kernel, nb0, 256, 0, stream, v, n, idx, val, mode, (long long *)nullptr);
        PB_LAUNCH(tree_top_kernel, 1, 512, 0, stream, v, mode, n, -1LL);
        return PB_OK;
    }
    if (sorted && n <= CHAIN_MAX_N) {
        const int grid = persistent_grid(n, CHAIN_THREADS / 32, 2);
        PB_LAUNCH(upd_chain_kernel, grid, CHAIN_THREADS, 0, stream, v, n, idx, val, mode, idx_out);
        return PB_OK;
    }
    const int nb = (int)((n + 255) / 256);
    if (n < v.cap / 16 && n <= LEAF_TAG_MAX_N) {
        PB_LAUNCH(upd_mark_kernel, nb, 256, 0, stream, v, n, idx, val, mode, idx_out, 1);
        PB_LAUNCH(upd_leaf_kernel, nb, 256, 0, stream, v, n, idx, val, mode, (long long *)nullptr);
        const long long spans = v.cap >= 32768 ? v.cap >> 15 : 1;
        const int fuse = v.P <= 3 ? 1 : 0;
        PB_LAUNCH(tree_rebuild_sparse_kernel, (int)(spans < 2048 ? spans : 2048), SPR_THREADS, 0, stream, v, mode, n, fuse);
        if (!fuse) {
            // deeper trees: levels L-15 .. TL by the streaming pass over the (small) array of level L-15
            TreeView u2 = v;
            for (int s2 = v.L - 15; s2 > v.TL;) {
                int n_lv = (s2 - v.TL) / 5;
                if (n_lv > 3) n_lv = 3;
                const int f2 = (s2 - 5 * n_lv <= v.TL) ? 1 : 0;
                const int grid = n_lv == 3 ? persistent_grid(1LL << s2, 32 * RB_TILE, 4)
                                           : persistent_grid(1LL << s2, RB_TILE * (RB_THREADS / 32), 8);
                PB_LAUNCH(tree_rebuild_kernel, grid, RB_THREADS, 0, stream, u2, s2, n_lv, (const float *)nullptr, 0LL, f2,
                          mode, n, -1LL);
                s2 -= 5 * n_lv;
            }
        }
        return PB_OK;
    }
    if (sorted) {
        PB_LAUNCH(upd_leaf_sorted_kernel, nb, 256, 0, stream, v, n, idx, val, mode, idx_out);
    } else {
        if (n > LEAF_TAG_MAX_N) return PB_E_UNSUPPORTED;        // the dedup tags are NaN payloads: 2^23 - 2 of them
        PB_LAUNCH(upd_mark_kernel, nb, 256, 0, stream, v, n, idx, val, mode, idx_out, 0);
        PB_LAUNCH(upd_leaf_kernel, nb, 256, 0, stream, v, n, idx, val, mode, (long long *)nullptr);
    }
    return launch_rebuild(v, nullptr, 0, mode, n, -1, stream);
}

}  // namespace

extern "C" {

int pb_tree_layout(long long capacity, long long *sum_floats, long long *min_floats, long long *counter_ints,
                   long long *leaf_offset, int *top_level)
{
    if (!pb_is_pow2(capacity) || capacity < 32 || capacity > (1LL << 30)) return PB_E_CAPACITY;
    Layout y;
    make_layout(capacity, &y);
    if (sum_floats) *sum_floats = y.n_sum;
    if (min_floats) *min_floats = y.n_min;
    if (counter_ints) *counter_ints = y.n_cnt;
    if (leaf_offset) *leaf_offset = y.P == 0 ? (1LL << y.L) : y.off[y.P];
    if (top_level) *top_level = y.TL;
    return PB_OK;
}

int pb_tree_init(const pb_tree *t, void *stream)
{
    TreeView v;
    int rc = make_view(t, &v);
    if (rc) return rc;
    if (!t->counters) return PB_E_ARG;
    Layout y;
    make_layout(t->capacity, &y);
    int nb = pb_sm_count() * 8;
    PB_LAUNCH(tree_init_kernel, nb, 256, 0, stream, v, y.n_sum, y.n_min, y.n_cnt);
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
    return launch_rebuild(v, leaves, n, MODE_RAW, 0, n, stream);
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
    PB_LAUNCH(tree_scan_kernel, persistent_grid(n, 256, 6), 256, 0, stream, v, n, mass, idx_out);
    return PB_OK;
}

int pb_tree_sample_batches(const pb_tree *t, long long n_batches, long long batch, const double *u, int mode,
                           float beta, long long *idx_out, float *weight_out, float *mass_out, void *stream)
{
    TreeView v;
    int rc = make_view(t, &v);
    if (rc) return rc;
    if (n_batches < 0 || batch < 0 || (mode != 0 && mode != 1)) return PB_E_ARG;
    const long long n = n_batches * batch;
    if (n >= (1LL << 31) || (n > 0 && (!idx_out || !weight_out))) return PB_E_ARG;
    if (n == 0) return PB_OK;
    if (n > 16384 && v.L >= 10) {
        // throughput mode: a thread per sample over a shared-memory heap of the top S levels (S = the deepest stored
        // level <= 14)
        int S = v.TL;
        while (S + 5 <= v.L && S + 5 <= 14) S += 5;
        const size_t smem = sizeof(float) * (2ull << S);
        static PbPerDeviceOnce attr_set;
        if (!attr_set.done()) {
            cudaError_t e = cudaFuncSetAttribute(tree_sample_thread_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)(sizeof(float) * (2ull << 14)));
            if (e != cudaSuccess) return (int)e;
            attr_set.mark();
        }
        const int per_sm = smem > 100 * 1024 ? 1 : 2;
        PB_LAUNCH(tree_sample_thread_kernel, persistent_grid(n, ST_THREADS, per_sm), ST_THREADS, smem, stream, v, n, batch, u,
                  mode, beta, idx_out, weight_out, mass_out, S);
        return PB_OK;
    }
    // one learner batch: one 8-lane descent chain per group, 4 samples per warp (latency)
    const int R = 1;
    const int sgrid = persistent_grid(n, 8 * 4 * R, 6);
    PB_LAUNCH(tree_sample_kernel, sgrid, 256, 0, stream, v, n, batch, u, mode, beta, idx_out, weight_out, mass_out, R);
    return PB_OK;
}

int pb_tree_sample(const pb_tree *t, long long n, const double *u, int mode, float beta, long long *idx_out,
                   float *weight_out, float *mass_out, void *stream)
{
    return pb_tree_sample_batches(t, n > 0 ? 1 : 0, n, u, mode, beta, idx_out, weight_out, mass_out, stream);
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
    if (n_global == 0) return PB_OK;
    const int R = n_global <= 16384 ? 1 : 8;
    PB_LAUNCH(global_sample_kernel, persistent_grid(n_global, 8 * 4 * R, 6), 256, 0, stream, v, all_state, n_ranks, rank,
              n_global, u, beta, idx_out, weight_out, stratum_out, R);
    return PB_OK;
}

int pb_tree_export(const pb_tree *t, float *sum_heap, float *min_heap, void *stream)
{
    TreeView v;
    int rc = make_view(t, &v);
    if (rc) return rc;
    if (!sum_heap && !min_heap) return PB_E_ARG;
    for (int d = v.L; d >= 0; --d)
        PB_LAUNCH(tree_export_kernel, (unsigned)(((1LL << d) + 255) / 256), 256, 0, stream, v, d, sum_heap, min_heap);
    return PB_OK;
}

}  // extern "C"
