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


@settings(max_examples=40, deadline=None, derandomize=True)
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


class _PyTree(object):
    """Independent numpy restatement of the rules in SURVEY 8(c) (1)-(7), scalar by scalar in np.float32 -- a
    second witness for oracle/per_oracle.c (alpha = 0.5 only)."""

    def __init__(self, size, eps=1e-8):
        self.size, self.cap = size, 1
        while self.cap < size:
            self.cap *= 2
        self.sum = np.zeros(2 * self.cap, np.float32)
        self.min = np.full(2 * self.cap, np.inf, np.float32)
        self.len, self.cursor, self.max_priority, self.eps = 0, 0, 1.0, eps

    def _set(self, i, v):
        pos = i | self.cap
        self.sum[pos] = self.min[pos] = np.float32(v)
        while pos > 1:
            left, right = pos & ~1, pos | 1
            self.sum[pos >> 1] = np.float32(self.sum[left] + self.sum[right])
            self.min[pos >> 1] = min(self.min[left], self.min[right])
            pos >>= 1

    def _query(self, arr, op, ident):
        if self.len >= self.size:
            return arr[1]
        l, r, ret = self.cap, self.len + self.cap, ident
        while l < r:
            if l & 1:
                ret = op(ret, arr[l]); l += 1
            if r & 1:
                r -= 1; ret = op(ret, arr[r])
            l >>= 1; r >>= 1
        return ret

    def extend(self, n):
        for _ in range(n):
            self._set(self.cursor, np.float32((self.max_priority + self.eps) ** 0.5))
            self.cursor = (self.cursor + 1) % self.size
            self.len = min(self.len + 1, self.size)

    def update_priority(self, idx, prio):
        prio = np.asarray(prio, np.float32)
        if len(prio):
            self.max_priority = max(self.max_priority, float(prio.max()))
        for i, p in zip(idx, prio):
            self._set(int(i), np.sqrt(np.float32(p + np.float32(self.eps))))

    def sample(self, u, beta, stratified):
        p_sum = self._query(self.sum, lambda a, b: np.float32(a + b), np.float32(0))
        p_min = self._query(self.min, min, np.float32(np.inf))
        idx, w, n = [], [], len(u)
        for k, uk in enumerate(u):
            m = np.float32(((k + uk) / n) * float(p_sum)) if stratified else np.float32(float(p_sum) * uk)
            if m > self.sum[1]:
                i = self.size
            else:
                pos = 1
                while pos < self.cap:
                    pos <<= 1
                    if m > self.sum[pos]:
                        m = np.float32(m - self.sum[pos])
                        pos |= 1
                i = pos ^ self.cap
            i = min(i, self.len - 1)
            idx.append(i)
            w.append(np.float32(self.sum[i | self.cap] / p_min) ** np.float32(-beta))
        return np.asarray(idx), np.asarray(w, np.float32), p_sum, p_min


def test_c_oracle_against_independent_numpy_model():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None, derandomize=True)
    @given(st.integers(0, 2 ** 31 - 1), st.integers(1, 70), st.integers(1, 12))
    def check(seed, size, n_ops):
        rng = np.random.default_rng(seed)
        c, p = OracleTree(size), _PyTree(size)
        for _ in range(n_ops):
            op = rng.integers(0, 3)
            if op == 0 or len(c) == 0:
                n = int(rng.integers(1, size + 3))
                c.extend(n); p.extend(n)
            elif op == 1:
                k = int(rng.integers(1, 2 * size))
                idx = rng.integers(0, len(c), k)                          # duplicates: the last one wins
                prio = rng.exponential(1.0, k).astype(np.float32) * (rng.random(k) > 0.1)     # some exact zeros
                c.update_priority(idx, prio); p.update_priority(idx, prio)
            else:
                u = rng.random(int(rng.integers(1, 20)))
                mode = int(rng.integers(0, 2))
                if p._query(p.min, min, np.float32(np.inf)) <= 0:
                    with pytest.raises(RuntimeError):
                        c.sample(u, 0.5, mode)
                    continue
                ci, cw, _, cs, cm_ = c.sample(u, 0.5, mode)
                pi, pw, ps, pm = p.sample(u, 0.5, bool(mode))
                assert np.array_equal(ci, pi) and cs == ps and cm_ == pm
                assert np.allclose(cw, pw, rtol=1e-6, atol=0)
            assert np.array_equal(c.sum, p.sum) and np.array_equal(c.min, p.min)
            assert len(c) == p.len and c.cursor == p.cursor and c.max_priority == p.max_priority

    check()
