"""C restatement of the sum/min tree + sampler (oracle/per_oracle.c).  PARITY UNPINNED against
torchrl (absent); these tests pin the restatement to its own frozen vectors (tests/golden/per_tree.npz)
and to the algebraic properties the published algorithm guarantees."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from helpers import load_golden
from oracle.per_oracle import OracleTree


def test_frozen_vectors():
    fx = load_golden("per_tree")
    N = int(fx["N"])
    t = OracleTree(N)
    t.extend(N)
    t.update_priority(np.arange(N), fx["p0"])
    for r in range(4):
        idx, w, mass, ps, pm = t.sample(fx["r%d.u" % r], 0.5, mode=r % 2)
        assert np.array_equal(idx, fx["r%d.idx" % r])
        assert np.array_equal(mass, fx["r%d.mass" % r])
        assert np.allclose(w, fx["r%d.w" % r], rtol=1e-6)
        assert np.float32(ps) == fx["r%d.psum" % r] and np.float32(pm) == fx["r%d.pmin" % r]
        t.update_priority(idx, fx["r%d.newp" % r])
        assert t.sum[1] == fx["r%d.root" % r]
        assert t.max_priority == float(fx["r%d.maxp" % r])
    assert np.array_equal(t.sum, fx["final.sum"]) and np.array_equal(t.min, fx["final.min"])


def test_nodes_are_pairwise_sums_and_build_equals_updates():
    rng = np.random.default_rng(0)
    N = 777
    leaves = rng.exponential(1.0, N).astype(np.float32)
    a, b = OracleTree(N), OracleTree(N)
    a.build(leaves)
    perm = rng.permutation(N)
    b.update_leaves(perm, leaves[perm])          # arbitrary order -> same tree
    b.set_len(N)
    assert np.array_equal(a.sum[1:], b.sum[1:]) and np.array_equal(a.min[1:], b.min[1:])
    cap = a.capacity
    s = a.sum
    assert np.array_equal(s[1:cap], (s[2:2 * cap:2] + s[3:2 * cap:2]).astype(np.float32))
    assert np.array_equal(a.min[1:cap], np.minimum(a.min[2:2 * cap:2], a.min[3:2 * cap:2]))


def test_capacity_choice_does_not_change_indices():
    """torchrl sizes the tree to the first pow2 STRICTLY above N; any pow2 >= N gives the same answers."""
    rng = np.random.default_rng(3)
    for N in (512, 1000, 1024):
        leaves = rng.exponential(1.0, N).astype(np.float32)
        a, b = OracleTree(N, strict_pow2=False), OracleTree(N, strict_pow2=True)
        a.build(leaves); b.build(leaves)
        u = rng.random(500)
        ia, wa, ma, _, _ = a.sample(u)
        ib, wb, mb, _, _ = b.sample(u)
        assert np.array_equal(ia, ib) and np.array_equal(wa, wb) and np.array_equal(ma, mb)


def test_last_duplicate_wins_and_max_priority():
    t = OracleTree(100)
    t.extend(100)
    t.update_priority(np.array([5, 7, 5]), np.array([1.0, 2.0, 9.0], np.float32))
    assert t.sum[t.capacity + 5] == np.sqrt(np.float32(9.0) + np.float32(1e-8))
    assert t.max_priority == 9.0
    assert t.default_priority() == np.float32(np.sqrt(9.0 + 1e-8))


def test_partial_fill_uses_interval_walk():
    t = OracleTree(1000)
    t.extend(37)
    assert len(t) == 37 and t.cursor == 37
    assert t.query_sum(0, 37) == np.float32(37.0) and t.query_min(0, 37) == np.float32(1.0)
    idx, w, mass, ps, pm = t.sample(np.array([0.0, 0.5, 0.999999]))
    assert idx.max() <= 36 and np.all(w == 1.0)


def test_empty_raises():
    with pytest.raises(RuntimeError):
        OracleTree(10).sample(np.array([0.5]))


def test_zero_priority_leaves_are_never_sampled():
    N = 64
    leaves = np.ones(N, np.float32)
    leaves[::2] = 0.0
    t = OracleTree(N)
    t.build(leaves)
    idx = t.scan(np.linspace(1e-3, 31.99, 300).astype(np.float32))
    assert np.all(idx % 2 == 1)


@settings(max_examples=40, deadline=None)
@given(st.integers(2, 300), st.integers(0, 2 ** 31 - 1))
def test_scan_is_monotone_and_matches_cumsum_when_exact(n, seed):
    rng = np.random.default_rng(seed)
    leaves = rng.integers(0, 5, n).astype(np.float32)   # small integers: every fp32 sum is exact
    if leaves.sum() == 0:
        leaves[0] = 1
    t = OracleTree(n)
    t.build(leaves)
    mass = np.sort(rng.random(64) * leaves.sum()).astype(np.float32)
    idx = t.scan(mass)
    assert np.all(np.diff(idx) >= 0)
    # exact arithmetic: scan_lower_bound == first i with cumsum[i] >= mass
    cs = np.cumsum(leaves.astype(np.float64))
    expect = np.searchsorted(cs, mass.astype(np.float64), side="left")
    assert np.array_equal(idx, np.minimum(expect, n))
