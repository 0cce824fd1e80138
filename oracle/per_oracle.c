/*
 * oracle/per_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, single thread) of the prioritized-replay arithmetic
 * that AechPro/Prism delegates to torchrl's PrioritizedReplayBuffer:
 *   constructed at  prism/factory/exp_buffer_factory.py:22-28
 *   used at         prism/experience/timestep_buffer.py:33,37,44,54
 *   driven by       prism/learner.py:100,106-107,120
 *
 * PARITY UNPINNED for the tree/sampler arithmetic: torchrl (and tensordict) are
 * a third-party dependency that is absent from /root/reference, not vendored,
 * not installed in the build image, and the reference pins no version (it has
 * no requirements/pyproject at all).  The reference holds no golden vector or
 * test for sampled indices, IS weights or priority updates.  What follows
 * restates torchrl's *published* algorithm (torchrl/csrc/segment_tree.h,
 * torchrl/data/replay_buffers/samplers.py::PrioritizedSampler, 0.3-0.6 era)
 * from its documented behaviour:
 *
 *  (1) tree: values[2*cap]; leaves at [cap, 2cap); identity 0 (sum) / +inf (min);
 *      update(i,v): values[i|cap]=v; while (i>1) values[i>>1] = op(values[i], values[i^1])
 *      -> every internal node is fl32(left (+) right) of its two children.
 *  (2) query(l,r): root when the range covers [0,size), else the bottom-up
 *      interval walk (different fp32 association than the root!).
 *  (3) scan_lower_bound(m): m > root -> size; else descend: go right and subtract
 *      left iff m > left.  Mass is converted to fp32 first.
 *  (4) sample: p_sum/p_min = query(0,len); mass = uniform(0,p_sum) (numpy, fp64);
 *      idx = clamp(scan(mass), max=len-1); w = (leaf[idx]/p_min)^(-beta) (fp32).
 *  (5) extend: slot priority = (max_priority + eps)^alpha, max_priority starts 1.
 *  (6) update_priority: max_priority = max(max_priority, max p); leaf=(p+eps)^alpha.
 *  (7) writer: round-robin cursor.
 *
 * Every version-dependent choice is a flag so that a later check against a real
 * torchrl only flips flags (see po_config).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product path (prism_b200/) never
 * does, and fails loudly when its CUDA library is missing.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int64_t size;        /* number of addressable slots (buffer capacity N) */
    int64_t capacity;    /* power-of-two number of leaves */
    float  *sum;         /* 2*capacity */
    float  *min;         /* 2*capacity */
    double  max_priority;/* python float in torchrl */
    int64_t cursor;      /* round-robin writer */
    int64_t len;         /* number of filled slots */
    /* version flags */
    int     strict_pow2; /* 1: capacity = first pow2 STRICTLY greater than size (torchrl loop) */
    int     weight_eps_in_denominator; /* 1: w = (p/(p_min+eps))^-beta */
    int     default_priority_fp64;     /* 1: (max_p+eps)^alpha evaluated in double then cast */
    float   alpha;
    double  eps;         /* python float in torchrl; cast to fp32 where torch does fp32 tensor math */
} po_tree;

static int64_t po_pow2_capacity(int64_t size, int strict)
{
    int64_t c = 1;
    if (strict) { while (c <= size) c <<= 1; }
    else        { while (c <  size) c <<= 1; }
    return c;
}

po_tree *po_create(int64_t size, float alpha, double eps, int strict_pow2,
                   int weight_eps_in_denominator, int default_priority_fp64)
{
    po_tree *t = (po_tree *)calloc(1, sizeof(po_tree));
    if (!t) return NULL;
    t->size = size;
    t->capacity = po_pow2_capacity(size, strict_pow2);
    t->sum = (float *)malloc(sizeof(float) * 2 * (size_t)t->capacity);
    t->min = (float *)malloc(sizeof(float) * 2 * (size_t)t->capacity);
    if (!t->sum || !t->min) { free(t->sum); free(t->min); free(t); return NULL; }
    for (int64_t i = 0; i < 2 * t->capacity; ++i) { t->sum[i] = 0.0f; t->min[i] = INFINITY; }
    t->max_priority = 1.0;
    t->alpha = alpha; t->eps = eps;
    t->strict_pow2 = strict_pow2;
    t->weight_eps_in_denominator = weight_eps_in_denominator;
    t->default_priority_fp64 = default_priority_fp64;
    return t;
}

void po_destroy(po_tree *t) { if (t) { free(t->sum); free(t->min); free(t); } }

int64_t po_capacity(const po_tree *t) { return t->capacity; }
int64_t po_len(const po_tree *t) { return t->len; }
int64_t po_cursor(const po_tree *t) { return t->cursor; }
double  po_max_priority(const po_tree *t) { return t->max_priority; }
float  *po_sum_ptr(po_tree *t) { return t->sum; }
float  *po_min_ptr(po_tree *t) { return t->min; }
void    po_set_len(po_tree *t, int64_t len, int64_t cursor) { t->len = len; t->cursor = cursor; }
void    po_set_max_priority(po_tree *t, double p) { t->max_priority = p; }

/* (1) point update, both trees, fp32 */
static void po_set_leaf(po_tree *t, int64_t i, float v)
{
    int64_t k = i | t->capacity;
    t->sum[k] = v; t->min[k] = v;
    while (k > 1) {
        float a = t->sum[k], b = t->sum[k ^ 1];
        /* node = fl32(left + right); fp32 addition is commutative so operand order is moot */
        volatile float s = a + b;
        t->sum[k >> 1] = s;
        float ma = t->min[k], mb = t->min[k ^ 1];
        t->min[k >> 1] = ma < mb ? ma : mb;
        k >>= 1;
    }
}

/* batched update = the point update applied sequentially (last duplicate wins) */
void po_update_leaves(po_tree *t, int64_t n, const int64_t *idx, const float *vals)
{
    for (int64_t j = 0; j < n; ++j) po_set_leaf(t, idx[j], vals[j]);
}

/* (2) interval query [l, r) */
float po_query_sum(const po_tree *t, int64_t l, int64_t r)
{
    if (l <= 0 && r >= t->size) return t->sum[1];
    volatile float ret = 0.0f;
    l |= t->capacity; r |= t->capacity;   /* r == capacity (full) cannot reach here */
    while (l < r) {
        if (l & 1) ret = ret + t->sum[l++];
        if (r & 1) ret = ret + t->sum[--r];
        l >>= 1; r >>= 1;
    }
    return ret;
}

float po_query_min(const po_tree *t, int64_t l, int64_t r)
{
    if (l <= 0 && r >= t->size) return t->min[1];
    float ret = INFINITY;
    l |= t->capacity; r |= t->capacity;
    while (l < r) {
        if (l & 1) { float v = t->min[l++]; ret = v < ret ? v : ret; }
        if (r & 1) { float v = t->min[--r]; ret = v < ret ? v : ret; }
        l >>= 1; r >>= 1;
    }
    return ret;
}

/* (3) prefix-sum descent on fp32 mass */
static int64_t po_scan_one(const po_tree *t, float mass)
{
    if (mass > t->sum[1]) return t->size;
    int64_t i = 1;
    volatile float m = mass;
    while (i < t->capacity) {
        i <<= 1;
        float left = t->sum[i];
        if (m > left) { m = m - left; i |= 1; }
    }
    return i ^ t->capacity;
}

void po_scan_lower_bound(const po_tree *t, int64_t n, const float *mass, int64_t *out)
{
    for (int64_t j = 0; j < n; ++j) out[j] = po_scan_one(t, mass[j]);
}

static float po_pow_leaf(float p, float eps, float alpha)
{
    /* torch.pow(p + eps, alpha) on fp32 CPU tensors; exponent 0.5 takes torch's
     * sqrt fast path, which is correctly rounded -> sqrtf */
    volatile float x = p + eps;
    if (alpha == 0.5f) return sqrtf(x);
    if (alpha == 1.0f) return x;
    return powf(x, alpha);
}

/* (4) sample from injected uniforms u in [0,1) (fp64, like numpy's random_sample).
 * mode 0: iid (torchrl):      mass = 0 + (p_sum - 0) * u_k
 * mode 1: stratified (north star): mass = ((k + u_k) / n) * p_sum
 * returns 0, or -1 empty, -2 p_sum<=0, -3 p_min<=0 */
int po_sample(const po_tree *t, int64_t n, const double *u, int mode, float beta,
              int64_t *idx_out, float *w_out, float *mass_out, float *psum_out, float *pmin_out)
{
    if (t->len <= 0) return -1;
    float p_sum = po_query_sum(t, 0, t->len);
    float p_min = po_query_min(t, 0, t->len);
    if (psum_out) *psum_out = p_sum;
    if (pmin_out) *pmin_out = p_min;
    if (!(p_sum > 0.0f)) return -2;
    if (!(p_min > 0.0f)) return -3;
    float denom = t->weight_eps_in_denominator ? (p_min + (float)t->eps) : p_min;
    for (int64_t k = 0; k < n; ++k) {
        double m64 = (mode == 0) ? (0.0 + ((double)p_sum - 0.0) * u[k])
                                 : (((double)k + u[k]) / (double)n) * (double)p_sum;
        float m = (float)m64;
        int64_t i = po_scan_one(t, m);
        if (i > t->len - 1) i = t->len - 1;
        float leaf = t->sum[i | t->capacity];
        volatile float ratio = leaf / denom;
        idx_out[k] = i;
        w_out[k] = powf(ratio, -beta);
        if (mass_out) mass_out[k] = m;
    }
    return 0;
}

/* (5) ring write of n new slots at the cursor with the default priority */
float po_default_priority(const po_tree *t)
{
    if (t->default_priority_fp64)
        return (float)pow(t->max_priority + t->eps, (double)t->alpha);
    return po_pow_leaf((float)t->max_priority, (float)t->eps, t->alpha);
}

void po_extend(po_tree *t, int64_t n, int64_t *idx_out)
{
    for (int64_t j = 0; j < n; ++j) {
        int64_t i = t->cursor;
        po_set_leaf(t, i, po_default_priority(t));
        if (idx_out) idx_out[j] = i;
        t->cursor = (t->cursor + 1) % t->size;
        if (t->len < t->size) t->len++;
    }
}

/* (6) priority write-back */
void po_update_priority(po_tree *t, int64_t n, const int64_t *idx, const float *prio)
{
    float mx = -INFINITY;
    for (int64_t j = 0; j < n; ++j) if (prio[j] > mx) mx = prio[j];
    if (n > 0 && (double)mx > t->max_priority) t->max_priority = (double)mx;
    for (int64_t j = 0; j < n; ++j) po_set_leaf(t, idx[j], po_pow_leaf(prio[j], (float)t->eps, t->alpha));
}

/* bulk load of leaves (already post-pow fp32) + pairwise rebuild; same node values
 * as n point updates because every node is fl32(l+r) whatever the update order. */
void po_build(po_tree *t, int64_t n, const float *leaves)
{
    int64_t cap = t->capacity;
    for (int64_t i = 0; i < cap; ++i) {
        t->sum[cap + i] = i < n ? leaves[i] : 0.0f;
        t->min[cap + i] = i < n ? leaves[i] : INFINITY;
    }
    for (int64_t k = cap - 1; k >= 1; --k) {
        volatile float s = t->sum[2 * k] + t->sum[2 * k + 1];
        t->sum[k] = s;
        float a = t->min[2 * k], b = t->min[2 * k + 1];
        t->min[k] = a < b ? a : b;
    }
    if (n > t->len) { t->len = n < t->size ? n : t->size; t->cursor = n % t->size; }
}

/* ------------------------------------------------------------------------- *
 * n-step return + successor resolution over the SoA transition ring.
 * Restates prism/experience/timestep_buffer.py:198-238 (_compute_n_step) and the
 * link rules of multiprocessing_experience_collection/collector_process_interface.py:146-173
 * on arrays: every stored step has a global sequence number `seq` (its extend
 * ordinal); slot = seq % size; a link is alive iff slot_seq[link % size] == link
 * (that is the weakref-died-on-overwrite rule).
 *   next_link[s] >= 0 : seq of the stored successor
 *   next_link[s] == -1: none (terminal)
 *   next_link[s] <= -2: aux observation -(link+2) (in-flight successor or the
 *                       truncated final observation; never a ring slot)
 * Outputs per start slot: R (fp64 accumulate -> fp32), gamma^k, done,
 * last slot visited, successor code (same encoding as next_link).
 * ------------------------------------------------------------------------- */
void po_nstep(int64_t size, int n_step, double gamma,
              const int64_t *slot_seq, const int64_t *next_link,
              const float *reward, const uint8_t *done, const uint8_t *trunc,
              int64_t n, const int64_t *start, float *ret_out, float *gamma_out,
              uint8_t *done_out, int64_t *last_out, int64_t *succ_out)
{
    double gammas[65];
    /* python: [gamma ** i for i in range(n_step + 1)] */
    for (int i = 0; i <= n_step && i < 65; ++i) gammas[i] = pow(gamma, (double)i);
    for (int64_t j = 0; j < n; ++j) {
        int64_t s = start[j];
        double ret = 0.0, g = 1.0;
        for (int i = 0; i < n_step; ++i) {
            ret += (double)reward[s] * gammas[i];
            g = gammas[i + 1];
            int incomplete = (i != n_step - 1);
            int64_t nl = next_link[s];
            if (nl >= 0 && !trunc[s] && incomplete) {
                int64_t ns = nl % size;
                if (slot_seq[ns] == nl) s = ns; else break;
            } else break;
        }
        ret_out[j] = (float)ret;
        gamma_out[j] = (float)g;
        done_out[j] = done[s];
        last_out[j] = s;
        int64_t nl = next_link[s];
        if (nl >= 0 && slot_seq[nl % size] != nl) nl = -1; /* dead weakref */
        succ_out[j] = nl;
    }
}
