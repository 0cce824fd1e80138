"""Device-resident prioritized-replay priority store (sum-tree + min-tree).

Stands in for the sampler half of torchrl's ``PrioritizedReplayBuffer`` that the reference
builds at prism/factory/exp_buffer_factory.py:22-28 and drives from
prism/experience/timestep_buffer.py:33,37,54 and prism/learner.py:100-107,119-120.
All arithmetic runs in libprism_b200 (csrc/per_tree.cu); this class only owns the
tensors and forwards raw pointers.  There is no CPU path.
"""
import ctypes as C

import torch

from .. import _lib


def _pow2_ceil(n):
    c = 1
    while c < n:
        c <<= 1
    return c


class PrioritizedTree:
    """Sum/min segment trees over ``size`` slots in the compact 32-ary layout of csrc/per_tree.cu (every 5th level
    stored, the levels in between rebuilt in registers; the min tree shares the leaf array).  ``sum`` / ``min`` export
    the full level-ordered arrays a pointer-walking tree would hold (tests, checkpoints); ``leaves()`` is a view.

    Attribute names follow torchrl's PrioritizedSampler where the reference pokes at
    them (``_beta`` is written by prism/learner.py:107; ``_alpha``, ``_eps``).
    """

    MODE_IID = 0          # torchrl: mass ~ U(0, p_sum) iid
    MODE_STRATIFIED = 1   # north star: mass_k = (k + u_k)/B * p_sum

    def __init__(self, size, alpha=0.5, beta=0.5, eps=1e-8, device="cuda:0", mode="iid",
                 weight_eps_in_denominator=False, default_priority_fp64=True):
        self._lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.PbError("PrioritizedTree needs a CUDA device (no CPU fallback); got %s" % device)
        self.size = int(size)
        if self.size <= 0:
            raise ValueError("size must be positive")
        self.capacity = max(32, _pow2_ceil(self.size))
        self._alpha = float(alpha)
        self._beta = float(beta)
        self._eps = float(eps)
        self.mode = self.MODE_STRATIFIED if mode in ("stratified", 1) else self.MODE_IID
        n_sum, n_min, n_cnt, leaf_off, top = (C.c_longlong(), C.c_longlong(), C.c_longlong(), C.c_longlong(), C.c_int())
        _lib.check(self._lib.pb_tree_layout(self.capacity, C.byref(n_sum), C.byref(n_min), C.byref(n_cnt),
                                            C.byref(leaf_off), C.byref(top)), "pb_tree_layout")
        self.top_level, self._leaf_offset = top.value, leaf_off.value
        self.sum_store = torch.empty(n_sum.value, dtype=torch.float32, device=self.device)
        self.min_store = torch.empty(n_min.value, dtype=torch.float32, device=self.device)
        self.counters = torch.zeros(n_cnt.value, dtype=torch.int32, device=self.device)
        self.state = torch.zeros(64, dtype=torch.uint8, device=self.device)
        self._c = _lib.pb_tree(
            sum=self.sum_store.data_ptr(), min=self.min_store.data_ptr(), owner=None,
            counters=self.counters.data_ptr(), state=self.state.data_ptr(), capacity=self.capacity, size=self.size,
            alpha=self._alpha, eps_f32=self._eps, eps_f64=self._eps,
            weight_eps_in_denominator=int(bool(weight_eps_in_denominator)),
            default_priority_fp64=int(bool(default_priority_fp64)))
        self._ref = C.byref(self._c)
        self.reset()
        self.seed(0x5EED)

    # -- helpers -----------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _i64(self, t):
        t = torch.as_tensor(t)
        if t.device != self.device or t.dtype != torch.int64 or not t.is_contiguous():
            t = t.to(device=self.device, dtype=torch.int64, non_blocking=True).contiguous()
        return t.view(-1)

    def _f32(self, t):
        t = torch.as_tensor(t)
        if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
            t = t.to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()
        return t.view(-1)

    # -- mutation ----------------------------------------------------------------
    def reset(self):
        _lib.check(self._lib.pb_tree_init(self._ref, self._stream()), "pb_tree_init")
        if getattr(self, "_seed", None) is not None:
            self.seed(self._seed)

    def build(self, leaves):
        """Bulk-load post-pow fp32 leaves for slots [0, n) and rebuild every node."""
        leaves = self._f32(leaves)
        _lib.check(self._lib.pb_tree_build(self._ref, leaves.data_ptr(), leaves.numel(), self._stream()),
                   "pb_tree_build")

    def extend(self, n, idx_out=None):
        """n new slots at the ring cursor get the default priority (max_priority+eps)^alpha."""
        if idx_out is not None:
            assert idx_out.dtype == torch.int64 and idx_out.is_cuda and idx_out.numel() >= n
        _lib.check(self._lib.pb_tree_extend(self._ref, int(n), _lib.ptr(idx_out), self._stream()),
                   "pb_tree_extend")

    def update_priority(self, index, priority, sorted=False):
        index = self._i64(index)
        priority = self._f32(priority)
        assert index.numel() == priority.numel()
        _lib.check(self._lib.pb_tree_update_priority(self._ref, index.numel(), index.data_ptr(),
                                                     priority.data_ptr(), int(bool(sorted)), self._stream()),
                   "pb_tree_update_priority")

    def set_leaves(self, index, leaves, sorted=False):
        index = self._i64(index)
        leaves = self._f32(leaves)
        assert index.numel() == leaves.numel()
        _lib.check(self._lib.pb_tree_set_leaves(self._ref, index.numel(), index.data_ptr(), leaves.data_ptr(),
                                                int(bool(sorted)), self._stream()), "pb_tree_set_leaves")

    def refresh_stats(self):
        _lib.check(self._lib.pb_tree_stats(self._ref, self._stream()), "pb_tree_stats")

    # -- sampling ----------------------------------------------------------------
    def scan(self, mass, out=None):
        mass = self._f32(mass)
        if out is None:
            out = torch.empty(mass.numel(), dtype=torch.int64, device=self.device)
        _lib.check(self._lib.pb_tree_scan(self._ref, mass.numel(), mass.data_ptr(), out.data_ptr(),
                                          self._stream()), "pb_tree_scan")
        return out

    def sample(self, n, u=None, mode=None, beta=None, idx_out=None, weight_out=None, mass_out=None, n_batches=1):
        """Draw ``n_batches`` batches of n indices each, all against the current tree (outputs (n_batches * n,),
        batch-major).  ``u``: fp64 uniforms in [0,1) (device); drawn inside the kernel (Philox4x32-10 keyed by
        ``seed()``) when omitted.  ``n_batches`` > 1 = batches in flight: write their priorities back with
        ``update_priority(idx, prio, sorted=False)`` on the concatenation (later batches win)."""
        batch, n_batches = int(n), int(n_batches)
        n = batch * n_batches
        if u is not None:
            u = torch.as_tensor(u)
            if u.device != self.device or u.dtype != torch.float64 or not u.is_contiguous():
                u = u.to(device=self.device, dtype=torch.float64, non_blocking=True).contiguous()
        if idx_out is None:
            idx_out = torch.empty(n, dtype=torch.int64, device=self.device)
        if weight_out is None:
            weight_out = torch.empty(n, dtype=torch.float32, device=self.device)
        mode = self.mode if mode is None else (self.MODE_STRATIFIED if mode in ("stratified", 1) else self.MODE_IID)
        beta = self._beta if beta is None else float(beta)
        _lib.check(self._lib.pb_tree_sample_batches(self._ref, n_batches, batch, _lib.ptr(u), mode, beta,
                                                    idx_out.data_ptr(), weight_out.data_ptr(), _lib.ptr(mass_out),
                                                    self._stream()), "pb_tree_sample_batches")
        return idx_out, weight_out

    def sample_global(self, n_ranks, rank, all_state, n_global, u, beta=None,
                      idx_out=None, weight_out=None, stratum_out=None):
        """Sharded global stratified sampling: every rank evaluates all strata and keeps its
        run.  ``all_state``: (n_ranks, 64) uint8, the all-gathered shard state blocks."""
        n_global = int(n_global)
        if idx_out is None:
            idx_out = torch.empty(n_global, dtype=torch.int64, device=self.device)
        if weight_out is None:
            weight_out = torch.empty(n_global, dtype=torch.float32, device=self.device)
        beta = self._beta if beta is None else float(beta)
        assert all_state.dtype == torch.uint8 and all_state.numel() == 64 * n_ranks and all_state.is_contiguous()
        assert u is None or (u.dtype == torch.float64 and u.is_cuda)
        _lib.check(self._lib.pb_tree_sample_global(self._ref, int(n_ranks), int(rank), all_state.data_ptr(),
                                                   n_global, _lib.ptr(u),
                                                   beta, idx_out.data_ptr(), weight_out.data_ptr(),
                                                   _lib.ptr(stratum_out), self._stream()), "pb_tree_sample_global")
        return idx_out, weight_out

    def sample_global_peer(self, peer, n_global, u, beta=None, idx_out=None, weight_out=None, stratum_out=None):
        """``sample_global`` on the shard states of the latest ``PeerGroup.state_put`` exchange: the kernel itself waits
        (bounded) until every rank's put has landed in this rank's peer-mapped slot -- no separate barrier launch."""
        import ctypes as C
        n_global = int(n_global)
        if idx_out is None:
            idx_out = torch.empty(n_global, dtype=torch.int64, device=self.device)
        if weight_out is None:
            weight_out = torch.empty(n_global, dtype=torch.float32, device=self.device)
        beta = self._beta if beta is None else float(beta)
        assert u is None or (u.dtype == torch.float64 and u.is_cuda)
        _lib.check(self._lib.pb_tree_sample_global_peer(self._ref, C.byref(peer.c), n_global, _lib.ptr(u), beta,
                                                        idx_out.data_ptr(), weight_out.data_ptr(),
                                                        _lib.ptr(stratum_out), self._stream()),
                   "pb_tree_sample_global_peer")
        return idx_out, weight_out

    def seed(self, seed, call=0):
        """Key of the in-kernel Philox generator (all ranks of a sharded buffer must share it)."""
        self._seed = int(seed) & 0x7FFFFFFF
        words = self.state.view(torch.int32)
        words[13] = int(call) & 0x7FFFFFFF
        words[14] = self._seed

    # -- host-visible state (synchronises; not on the hot path) -------------------
    def state_host(self):
        raw = self.state.cpu().numpy().tobytes()
        s = _lib.pb_per_state.from_buffer_copy(raw)
        return {"len": s.len, "seq": s.seq, "max_priority": s.max_priority, "p_sum": s.p_sum, "p_min": s.p_min,
                "status": s.status, "owned_lo": s.owned_lo, "owned_n": s.owned_n}

    def stats_tensor(self):
        """{p_sum, p_min} as a 2-float device view of the state block (feeds the all-gather)."""
        return self.state.view(torch.float32)[5:7]

    def leaves(self):
        """The post-pow fp32 leaf values of slots [0, size): a device view, no copy."""
        return self.sum_store[self._leaf_offset:self._leaf_offset + self.size]

    def export(self):
        """(sum, min): the full level-ordered arrays (2 * capacity floats each; node i has children 2i, 2i+1; leaves at
        [capacity, 2 * capacity); never-written min leaves +inf) -- what the reference's segment trees hold."""
        s = torch.empty(2 * self.capacity, dtype=torch.float32, device=self.device)
        m = torch.empty(2 * self.capacity, dtype=torch.float32, device=self.device)
        _lib.check(self._lib.pb_tree_export(self._ref, s.data_ptr(), m.data_ptr(), self._stream()), "pb_tree_export")
        return s, m

    @property
    def sum(self):
        return self.export()[0]

    @property
    def min(self):
        return self.export()[1]

    def snapshot(self):
        """Device copies of everything a mutation touches (graph warm-ups that must not disturb the priorities)."""
        return (self.sum_store.clone(), self.min_store.clone(), self.state.clone())

    def restore(self, snap):
        self.sum_store.copy_(snap[0]); self.min_store.copy_(snap[1]); self.state.copy_(snap[2])

    def len_tensor(self):
        return self.state.view(torch.int64)[0:1]

    @property
    def _max_priority(self):
        return self.state_host()["max_priority"]

    def state_dict(self):
        # every node is a deterministic function of the leaves: the leaves + the state block are the whole store
        return {"leaves": self.leaves().cpu(), "state": self.state.cpu(), "size": self.size,
                "alpha": self._alpha, "beta": self._beta, "eps": self._eps}

    def load_state_dict(self, sd):
        assert sd["size"] == self.size
        if "leaves" in sd:
            leaves = sd["leaves"]
        else:                                   # checkpoints of the full-heap layout: leaves at [capacity, capacity + size)
            cap = sd["sum"].numel() // 2
            leaves = sd["sum"][cap:cap + self.size]
        # slots at or past `len` were never written (they read as +inf on the min side): build from the filled prefix
        n_filled = int(sd["state"].view(torch.int64)[0])
        self.build(leaves[:n_filled].to(self.device))
        self.state.copy_(sd["state"])           # len / seq / max_priority / p_sum / p_min / rng as saved
        self.counters.zero_()
        self._beta = sd["beta"]
