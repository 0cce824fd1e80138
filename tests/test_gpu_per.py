"""GPU parity of the sum/min-tree kernels (csrc/per_tree.cu) against the CPU oracle, through the
C ABI.  Bar: bit-exact indices / masses / tree arrays (integer and fp32-add-order work); IS weights
within 1e-6 relative (pow)."""
import numpy as np
import pytest
import torch

from helpers import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _trees(N, **kw):
    from oracle.per_oracle import OracleTree
    from prism_b200 import PrioritizedTree
    return PrioritizedTree(N, device=DEV, **kw), OracleTree(N)


def _assert_same_tree(g, o):
    torch.cuda.synchronize()
    gs, gm = g.sum.cpu().numpy(), g.min.cpu().numpy()
    cap = min(g.capacity, o.capacity)
    assert g.capacity == o.capacity or o.capacity < g.capacity
    if g.capacity == o.capacity:
        assert np.array_equal(gs[1:], o.sum[1:]), "sum tree differs"
        assert np.array_equal(gm[1:], o.min[1:]), "min tree differs"
    else:  # tiny sizes are padded to 32 leaves on the device: compare leaves and root
        assert np.array_equal(gs[g.capacity:g.capacity + o.size], o.sum[o.capacity:o.capacity + o.size])
        assert gs[1] == o.sum[1] and gm[1] == o.min[1]


@pytest.mark.parametrize("N", [5, 1000, 16384, 100_000, (1 << 20) + 3, 1 << 22])
def test_bulk_build_is_bit_exact(N):
    rng = np.random.default_rng(N)
    leaves = np.sqrt(rng.exponential(1.0, N).astype(np.float32) + np.float32(1e-8))
    g, o = _trees(N)
    g.build(torch.from_numpy(leaves).to(DEV))
    o.build(leaves)
    _assert_same_tree(g, o)
    st = g.state_host()
    assert st["len"] == N and st["p_sum"] == o.sum[1] and st["p_min"] == o.min[1]


@pytest.mark.parametrize("N", [1000, 100_000, 1 << 21])
@pytest.mark.parametrize("kind", ["exp", "equal", "onehot", "loguniform", "zeros"])
def test_scan_lower_bound_is_bit_exact(N, kind):
    rng = np.random.default_rng(7)
    if kind == "exp":
        leaves = np.sqrt(rng.exponential(1.0, N).astype(np.float32))
    elif kind == "equal":
        leaves = np.full(N, 0.37, np.float32)
    elif kind == "onehot":
        leaves = np.zeros(N, np.float32); leaves[N // 3] = 2.5
    elif kind == "loguniform":
        leaves = (10.0 ** rng.uniform(-6, 6, N)).astype(np.float32)
    else:
        leaves = rng.exponential(1.0, N).astype(np.float32); leaves[rng.random(N) < 0.1] = 0.0
    g, o = _trees(N)
    g.build(torch.from_numpy(leaves).to(DEV)); o.build(leaves)
    root = o.sum[1]
    mass = (rng.random(4096) * root).astype(np.float32)
    mass[:4] = [0.0, root, np.nextafter(root, np.float32(np.inf)), np.float32(root) * 0.5]
    got = g.scan(torch.from_numpy(mass).to(DEV)).cpu().numpy()
    assert np.array_equal(got, o.scan(mass))
    assert got[2] == N       # mass > root -> size, as the reference


@pytest.mark.parametrize("N,B", [(1000, 64), (100_000, 256), (1 << 20, 4096), (1 << 20, 65536), (300_000, 20_000)])
@pytest.mark.parametrize("mode", ["iid", "stratified"])
def test_sample_matches_oracle(N, B, mode):
    rng = np.random.default_rng(11)
    prio = rng.exponential(1.0, N).astype(np.float32)
    g, o = _trees(N, mode=mode)
    g.extend(N); o.extend(N)
    g.update_priority(torch.arange(N, device=DEV), torch.from_numpy(prio).to(DEV))
    o.update_priority(np.arange(N), prio)
    _assert_same_tree(g, o)
    u = rng.random(B)
    mass_out = torch.empty(B, dtype=torch.float32, device=DEV)
    idx, w = g.sample(B, u=torch.from_numpy(u).to(DEV), mass_out=mass_out)
    oi, ow, om, ps, pm = o.sample(u, 0.5, mode=1 if mode == "stratified" else 0)
    assert np.array_equal(idx.cpu().numpy(), oi)
    assert np.array_equal(mass_out.cpu().numpy(), om)
    assert np.allclose(w.cpu().numpy(), ow, rtol=1e-6, atol=0)
    if mode == "stratified":
        assert np.all(np.diff(oi) >= 0)


@pytest.mark.parametrize("N", [1000, 40_000, 1 << 20])
def test_update_priority_sequences_are_bit_exact(N):
    """Random update batches with duplicates (last wins) + the sorted fast path, whole tree compared."""
    rng = np.random.default_rng(5)
    g, o = _trees(N, mode="stratified")
    g.extend(N); o.extend(N)
    for it in range(6):
        B = [1, 7, 256, 4096, 300, 33][it]
        if it % 2 == 0:
            idx = rng.integers(0, N, B)
            idx[B // 2:] = idx[:B - B // 2]            # force duplicates
            sorted_hint = False
        else:
            idx = np.sort(rng.integers(0, N, B))       # what a stratified sample looks like
            sorted_hint = True
        prio = rng.exponential(2.0, B).astype(np.float32)
        g.update_priority(torch.from_numpy(idx).to(DEV), torch.from_numpy(prio).to(DEV), sorted=sorted_hint)
        o.update_priority(idx, prio)
        _assert_same_tree(g, o)
        st = g.state_host()
        assert st["max_priority"] == np.float32(o.max_priority)
        assert st["p_sum"] == o.query_sum(0, len(o)) and st["p_min"] == o.query_min(0, len(o))
    assert bool(torch.isfinite(g.leaves()).all())    # no dedup tag (a NaN payload) left on a leaf slot
    assert int(g.counters.abs().sum()) == 0          # arrival counters of the one-launch update back at zero


@pytest.mark.parametrize("N", [100, 1000, 70_000])
def test_partial_fill_extend_and_interval_walk(N):
    rng = np.random.default_rng(2)
    g, o = _trees(N)
    total = 0
    for n in [1, 3, 17, N // 3, N // 2, N // 2, 5]:     # wraps around the ring
        n = min(n, N)
        idx_out = torch.empty(n, dtype=torch.int64, device=DEV)
        g.extend(n, idx_out=idx_out)
        oi = o.extend(n)
        assert np.array_equal(idx_out.cpu().numpy(), oi)
        total += n
        st = g.state_host()
        assert st["len"] == len(o) and st["seq"] == total
        assert st["p_sum"] == o.query_sum(0, len(o)) and st["p_min"] == o.query_min(0, len(o))
        B = min(32, len(o))
        idx = rng.integers(0, len(o), B)
        prio = rng.exponential(3.0, B).astype(np.float32)
        g.update_priority(torch.from_numpy(idx).to(DEV), torch.from_numpy(prio).to(DEV))
        o.update_priority(idx, prio)
        _assert_same_tree(g, o)
        u = rng.random(16)
        gi, gw = g.sample(16, u=torch.from_numpy(u).to(DEV))
        oi2, ow, _, _, _ = o.sample(u)
        assert np.array_equal(gi.cpu().numpy(), oi2)
        assert np.allclose(gw.cpu().numpy(), ow, rtol=1e-6)


@pytest.mark.parametrize("N", [70_000, 1 << 20])
def test_large_batches_take_the_dense_path_bit_exact(N):
    """n >= cap/128: leaves are scattered and every level rebuilt by the streaming kernels; the
    thread-per-sample descent serves the large sample.  Same trees, same indices as the oracle."""
    rng = np.random.default_rng(17)
    g, o = _trees(N, mode="stratified")
    g.extend(N); o.extend(N)
    for it, B in enumerate([N // 16, N // 3, 40_000]):
        u = rng.random(B)
        gi, gw = g.sample(B, u=torch.from_numpy(u).to(DEV))
        oi, ow, _, _, _ = o.sample(u, 0.5, mode=1)
        assert np.array_equal(gi.cpu().numpy(), oi)
        assert np.allclose(gw.cpu().numpy(), ow, rtol=1e-6)
        prio = rng.exponential(1.5, B).astype(np.float32)
        if it == 1:                                   # unsorted with duplicates -> owner-scratch dedup + dense rebuild
            perm = rng.permutation(B)
            idx, prio = oi[perm], prio[perm]
            g.update_priority(torch.from_numpy(idx).to(DEV), torch.from_numpy(prio).to(DEV), sorted=False)
            o.update_priority(idx, prio)
        else:
            g.update_priority(gi, torch.from_numpy(prio).to(DEV), sorted=True)
            o.update_priority(oi, prio)
        _assert_same_tree(g, o)
        st = g.state_host()
        assert st["max_priority"] == np.float32(o.max_priority) and st["p_sum"] == o.sum[1]
    mass = (rng.random(50_000) * o.sum[1]).astype(np.float32)
    assert np.array_equal(g.scan(torch.from_numpy(mass).to(DEV)).cpu().numpy(), o.scan(mass))


def test_frozen_vectors_on_gpu():
    from prism_b200 import PrioritizedTree
    fx = load_golden("per_tree")
    N = int(fx["N"])
    g = PrioritizedTree(N, device=DEV)
    g.extend(N)
    g.update_priority(torch.arange(N, device=DEV), torch.from_numpy(fx["p0"]).to(DEV))
    for r in range(4):
        mass = torch.empty(64, dtype=torch.float32, device=DEV)
        idx, w = g.sample(64, u=torch.from_numpy(fx["r%d.u" % r]).to(DEV), mode=r % 2, mass_out=mass)
        assert np.array_equal(idx.cpu().numpy(), fx["r%d.idx" % r])
        assert np.array_equal(mass.cpu().numpy(), fx["r%d.mass" % r])
        assert np.allclose(w.cpu().numpy(), fx["r%d.w" % r], rtol=1e-6)
        g.update_priority(idx, torch.from_numpy(fx["r%d.newp" % r]).to(DEV), sorted=(r % 2 == 1))
        assert g.state_host()["max_priority"] == np.float32(fx["r%d.maxp" % r])
    cap = g.capacity
    assert np.array_equal(g.sum.cpu().numpy()[1:2 * cap], fx["final.sum"][1:2 * cap])
    assert np.array_equal(g.min.cpu().numpy()[1:2 * cap], fx["final.min"][1:2 * cap])


def test_empty_shard_sets_status():
    from prism_b200 import PrioritizedTree
    g = PrioritizedTree(256, device=DEV)
    idx, w = g.sample(8, u=torch.rand(8, dtype=torch.float64, device=DEV))
    assert g.state_host()["status"] & 1
    assert float(w.abs().sum()) == 0.0


def test_full_size_properties_2_24():
    """BASELINE config 3 size (16M-leaf tree, batch 4096): size-independent invariants."""
    from prism_b200 import PrioritizedTree
    N = 1 << 24
    g = PrioritizedTree(N, device=DEV, mode="stratified")
    gen = torch.Generator(device=DEV); gen.manual_seed(1)
    leaves = torch.rand(N, device=DEV, generator=gen).add_(1e-3).sqrt_()
    g.build(leaves)
    for it in range(3):
        idx, w = g.sample(4096, u=torch.rand(4096, dtype=torch.float64, device=DEV, generator=gen))
        assert bool((idx[1:] >= idx[:-1]).all()) and int(idx.min()) >= 0 and int(idx.max()) < N
        prio = torch.rand(4096, device=DEV, generator=gen) * 4
        g.update_priority(idx, prio, sorted=True)
        s, m = g.sum, g.min
        cap = g.capacity
        # every internal node is exactly fl32(left + right) / min(left, right)
        assert torch.equal(s[1:cap], s[2:2 * cap:2] + s[3:2 * cap:2])
        assert torch.equal(m[1:cap], torch.minimum(m[2:2 * cap:2], m[3:2 * cap:2]))
        # last occurrence of each index carries sqrt(p + eps)
        last = {}
        for j, i in enumerate(idx.cpu().tolist()):
            last[i] = j
        ii = torch.tensor(list(last.keys()), device=DEV)
        jj = torch.tensor(list(last.values()), device=DEV)
        assert torch.equal(s[cap + ii], (prio[jj] + 1e-8).sqrt())
    # the descent is the inverse of the prefix sum: scanning the exact prefix of leaf i lands on i
    probe = torch.randint(0, N, (512,), device=DEV, generator=gen)
    assert bool((g.scan(torch.zeros(512, device=DEV)) == 0).all())


def test_sharded_global_sampling_equals_one_big_tree():
    """G shards of capacity C concatenated == one tree of G*C leaves (fp32 pairwise top)."""
    from oracle.per_oracle import OracleTree
    from prism_b200 import PrioritizedTree
    G, C, B = 4, 1 << 12, 512
    rng = np.random.default_rng(9)
    leaves = np.sqrt(rng.exponential(1.0, G * C).astype(np.float32))
    big = OracleTree(G * C)
    big.build(leaves)
    shards = []
    for r in range(G):
        t = PrioritizedTree(C, device=DEV)
        t.build(torch.from_numpy(leaves[r * C:(r + 1) * C]).to(DEV))
        shards.append(t)
    torch.cuda.synchronize()
    all_state = torch.stack([t.state for t in shards]).contiguous()      # what the NCCL all-gather produces
    u = rng.random(B)
    oi, ow, _, _, _ = big.sample(u, 0.5, mode=1)
    ud = torch.from_numpy(u).to(DEV)
    got_idx = np.full(B, -1, np.int64)
    got_w = np.zeros(B, np.float32)
    for r in range(G):
        strat = torch.empty(B, dtype=torch.int64, device=DEV)
        idx, w = shards[r].sample_global(G, r, all_state, B, ud, stratum_out=strat)
        st = shards[r].state_host()
        n = st["owned_n"]
        k = strat[:n].cpu().numpy()
        assert np.array_equal(k, np.arange(st["owned_lo"], st["owned_lo"] + n))
        got_idx[k] = idx[:n].cpu().numpy() + r * C
        got_w[k] = w[:n].cpu().numpy()
        assert float(w[n:].abs().sum()) == 0.0
    assert np.array_equal(got_idx, oi)
    assert np.allclose(got_w, ow, rtol=1e-6)


def test_device_philox_uniforms():
    """u = None: uniforms are drawn inside the sampling kernel (Philox keyed by the shard seed, counter =
    (stratum, call number)); the call number advances per launch, equal seeds give equal draws."""
    from prism_b200 import PrioritizedTree
    N, B = 1 << 16, 2048
    leaves = torch.rand(N, device=DEV) + 0.1
    a, b = PrioritizedTree(N, device=DEV, mode="stratified"), PrioritizedTree(N, device=DEV, mode="stratified")
    a.build(leaves); b.build(leaves)
    a.seed(77); b.seed(77)
    m1 = torch.empty(B, device=DEV); m2 = torch.empty(B, device=DEV); m3 = torch.empty(B, device=DEV)
    i1, _ = a.sample(B, mass_out=m1)
    i1 = i1.clone()
    i2, _ = a.sample(B, mass_out=m2)
    i3, _ = b.sample(B, mass_out=m3)
    assert torch.equal(i1, i3) and torch.equal(m1, m3)        # same seed, same call number
    assert not torch.equal(m1, m2)                            # the call number advanced
    total = float(a.sum[1])
    k = torch.arange(B, device=DEV, dtype=torch.float64)
    u1 = m1.double() / total * B - k                           # recover u_k from the stratified mass
    assert float(u1.min()) > -1e-3 and float(u1.max()) < 1 + 1e-3
    assert 0.4 < float(u1.mean()) < 0.6 and 0.2 < float(u1.std()) < 0.4
    big, _ = a.sample(1 << 16)                                 # thread-per-sample kernel, same generator
    assert int(big.min()) >= 0 and int(big.max()) < N and bool((big[1:] >= big[:-1]).all())
    assert a.state_host()["status"] == 0


@pytest.mark.parametrize("N", [1 << 15, 1 << 20, (1 << 21) + 77])
@pytest.mark.parametrize("kind", ["onehot", "cluster", "one_line", "spread", "padded"])
def test_one_launch_sorted_update_adversarial_runs(N, kind):
    """The single-launch sorted update (leaders, arrival counters, last-CTA top) on index sets that stress it: every
    entry on ONE leaf (a run of 4096 duplicates), all entries inside one 1024-leaf line of the level above, a few dense
    clusters, an even spread, and a batch whose tail is -1 padding rows (sharded sampling)."""
    rng = np.random.default_rng(N % 1000 + len(kind))
    g, o = _trees(N, mode="stratified")
    g.extend(N); o.extend(N)
    B = 4096
    if kind == "onehot":
        idx = np.full(B, N // 3, np.int64)
    elif kind == "cluster":
        centers = rng.integers(0, N - 64, 7)
        idx = np.sort((centers[rng.integers(0, 7, B)] + rng.integers(0, 64, B)).astype(np.int64))
    elif kind == "one_line":
        base = (N // 2048) * 1024
        idx = np.sort(base + rng.integers(0, 1024, B)).astype(np.int64)
    else:
        idx = np.sort(rng.integers(0, N, B)).astype(np.int64)
    prio = rng.exponential(2.0, B).astype(np.float32)
    gi = idx.copy()
    if kind == "padded":
        gi[-300:] = -1
        idx, oprio = idx[:-300], prio[:-300]
    else:
        oprio = prio
    for _ in range(2):                                  # twice: the counters must come back to zero in between
        g.update_priority(torch.from_numpy(gi).to(DEV), torch.from_numpy(prio).to(DEV), sorted=True)
        o.update_priority(idx, oprio)
        _assert_same_tree(g, o)
        assert int(g.counters.abs().sum()) == 0
        st = g.state_host()
        assert st["max_priority"] == np.float32(o.max_priority)
        assert st["p_sum"] == o.query_sum(0, len(o)) and st["p_min"] == o.query_min(0, len(o))
        prio = prio[::-1].copy()
        oprio = prio[:len(idx)] if kind == "padded" else prio


@pytest.mark.parametrize("N,B", [(1 << 20, 4096), (1 << 22, 1024), (50_000, 256)])
@pytest.mark.parametrize("K", [1, 4, 16])
def test_batches_in_flight_equal_sequential_reference_calls(N, B, K):
    """K batches sampled in ONE launch against the same tree == K reference sample() calls with no update in
    between; writing all K*B priorities back in ONE call == K successive update_priority calls (later batches win)."""
    rng = np.random.default_rng(K * 7 + B)
    prio0 = rng.exponential(1.0, N).astype(np.float32)
    g, o = _trees(N, mode="stratified")
    g.extend(N); o.extend(N)
    g.update_priority(torch.arange(N, device=DEV), torch.from_numpy(prio0).to(DEV), sorted=True)
    o.update_priority(np.arange(N), prio0)
    for rnd in range(2):
        u = rng.random((K, B))
        mass = torch.empty(K * B, dtype=torch.float32, device=DEV)
        idx, w = g.sample(B, u=torch.from_numpy(u.reshape(-1)).to(DEV), n_batches=K, mass_out=mass)
        newp = rng.exponential(2.0, (K, B)).astype(np.float32)
        ois = []
        for b in range(K):
            oi, ow, om, _, _ = o.sample(u[b], 0.5, mode=1)
            assert np.array_equal(idx[b * B:(b + 1) * B].cpu().numpy(), oi), (rnd, b)
            assert np.array_equal(mass[b * B:(b + 1) * B].cpu().numpy(), om)
            assert np.allclose(w[b * B:(b + 1) * B].cpu().numpy(), ow, rtol=1e-6, atol=0)
            ois.append(oi)
        g.update_priority(idx, torch.from_numpy(newp.reshape(-1)).to(DEV), sorted=(K == 1))
        for b in range(K):
            o.update_priority(ois[b], newp[b])
        _assert_same_tree(g, o)
        st = g.state_host()
        assert st["max_priority"] == np.float32(o.max_priority) and st["p_sum"] == o.sum[1] and st["p_min"] == o.min[1]


def test_checkpoint_round_trip_and_legacy_heap_format():
    from prism_b200 import PrioritizedTree
    N = 70_000
    rng = np.random.default_rng(3)
    g, o = _trees(N)
    g.extend(50_000); o.extend(50_000)
    idx = rng.integers(0, 50_000, 4096)
    prio = rng.exponential(2.0, 4096).astype(np.float32)
    g.update_priority(torch.from_numpy(idx).to(DEV), torch.from_numpy(prio).to(DEV))
    o.update_priority(idx, prio)
    sd = g.state_dict()
    h = PrioritizedTree(N, device=DEV)
    h.load_state_dict(sd)
    _assert_same_tree(h, o)
    assert h.state_host() == g.state_host()
    legacy = {"sum": torch.from_numpy(o.sum.copy()), "min": torch.from_numpy(o.min.copy()), "state": sd["state"],
              "size": N, "alpha": 0.5, "beta": 0.5, "eps": 1e-8}
    h2 = PrioritizedTree(N, device=DEV)
    h2.load_state_dict(legacy)
    _assert_same_tree(h2, o)
