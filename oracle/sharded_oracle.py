"""CPU restatement of sharded global stratified sampling -- TEST INFRASTRUCTURE ONLY.

New design (the reference is single-process): every rank holds one shard (an OracleTree); the ranks
all-gather {p_sum, p_min, len}; the shard roots form the virtual top of one tree of G*cap leaves, summed
pairwise in fp32; stratum k of B_glob has mass (k + u_k)/B_glob * p_total, is routed through the virtual
top to its owner and descends the owner's tree with the residual mass.  With full power-of-two shards the
result is bit-identical to sampling one tree of G*cap leaves (tests/test_sharded_gloo.py).
"""
import numpy as np


def virtual_top(shard_psums):
    """Heap array (node i -> children 2i, 2i+1) over the G shard sums, fp32 pairwise."""
    G = len(shard_psums)
    assert G & (G - 1) == 0
    top = np.zeros(2 * G, np.float32)
    top[G:] = np.asarray(shard_psums, np.float32)
    for i in range(G - 1, 0, -1):
        top[i] = np.float32(top[2 * i] + top[2 * i + 1])
    return top


def route(top, k, n_global, u):
    """Owner shard and residual fp32 mass of stratum k."""
    G = len(top) // 2
    total = top[1]
    m = np.float32(((float(k) + float(u)) / float(n_global)) * float(total))
    if m > total:
        return G - 1, m
    node = 1
    while node < G:
        node <<= 1
        left = top[node]
        if m > left:
            m = np.float32(m - left)
            node |= 1
    return node - G, m


def sample_rank(tree, rank, all_psum, all_pmin, n_global, u, beta=0.5):
    """What rank `rank` computes: (strata owned, local indices, weights)."""
    top = virtual_top(all_psum)
    pmin = np.float32(min(all_pmin))
    ks, masses = [], []
    for k in range(n_global):
        owner, m = route(top, k, n_global, u[k])
        if owner == rank:
            ks.append(k)
            masses.append(m)
    idx = tree.scan(np.asarray(masses, np.float32)) if ks else np.zeros(0, np.int64)
    idx = np.minimum(idx, len(tree) - 1)
    leaf = tree.sum[tree.capacity + idx]
    w = np.power((leaf / pmin).astype(np.float32), np.float32(-beta)).astype(np.float32)
    return np.asarray(ks, np.int64), idx, w
