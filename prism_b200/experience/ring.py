"""Device-resident transition ring (struct-of-arrays) with pinned-host ingest staging.

Replaces the reference's ``ListStorage`` of linked ``Timestep`` objects
(prism/experience/timestep.py:12-28; links built at
multiprocessing_experience_collection/collector_process_interface.py:146-173) and the
Python batch loop of prism/experience/timestep_buffer.py:79-257.  Kernels: csrc/store.cu.
"""
import ctypes as C
import os

import numpy as np
import torch

from .. import _lib

STEP_DONE, STEP_TRUNC, STEP_NO_NEXT = 1, 2, 4
_I32, _I64, _F32, _U8, _BOOL = (np.dtype(t) for t in (np.int32, np.int64, np.float32, np.uint8, np.bool_))
# IngestSlot.fill stages a block with ONE host call (pb_store_stage_block) instead of a dozen numpy assignments when the
# arrays already have the staged layout; since the step dropped to ~72 us the host side of an iteration shows in the
# end-to-end rate (3.19 -> 3.31 M transitions/s on one box).  PB_NATIVE_STAGE=0: the numpy path always.  The call itself
# is covered on CPU (tests/test_oracle_buffer.py), both paths on the device (tests/test_gpu_buffer.py).
NATIVE_STAGE = os.environ.get("PB_NATIVE_STAGE", "1") != "0"


class TransitionRing:
    def __init__(self, size, obs_shape, frame_stack=1, n_step=3, gamma=0.99, storage_dtype=torch.float32,
                 obs_scale=False, max_streams=256, trunc_pool=None, staging_rows=256, device="cuda:0"):
        self._lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.PbError("TransitionRing needs a CUDA device (no CPU fallback); got %s" % device)
        self.size = int(size)
        self.obs_shape = tuple(int(s) for s in obs_shape)
        self.obs_elems = int(np.prod(self.obs_shape)) if len(self.obs_shape) else 1
        self.frame_stack = int(frame_stack)
        self.n_step = int(n_step)
        self.gamma = float(gamma)
        if storage_dtype not in (torch.float32, torch.uint8):
            raise ValueError("storage_dtype must be float32 or uint8")
        self.storage_dtype = storage_dtype
        self.obs_scale = bool(obs_scale)
        self.max_streams = int(max_streams)
        pool = int(trunc_pool) if trunc_pool is not None else max(64, self.size // 64)
        self.aux_size = self.max_streams + pool
        self.staging_rows = int(min(staging_rows, self.size))
        d = self.device
        self.obs = torch.zeros((self.size, self.obs_elems), dtype=storage_dtype, device=d)
        self.aux_obs = torch.zeros((self.aux_size, self.obs_elems), dtype=storage_dtype, device=d)
        self.action = torch.zeros(self.size, dtype=torch.int32, device=d)
        self.reward = torch.zeros(self.size, dtype=torch.float32, device=d)
        self.done = torch.zeros(self.size, dtype=torch.uint8, device=d)
        self.trunc = torch.zeros(self.size, dtype=torch.uint8, device=d)
        self.slot_seq = torch.full((self.size,), -1, dtype=torch.int64, device=d)
        self.next_link = torch.full((self.size,), -1, dtype=torch.int64, device=d)
        self.prev_link = torch.full((self.size,), -1, dtype=torch.int64, device=d)
        # bumped whenever a device array is re-allocated (the descriptor changes): captured CUDA graphs hold the
        # descriptor BY VALUE, so LearnerStep / ingest graphs compare generations and re-capture
        self.generation = 0
        self._make_desc()
        # host-side planner state (mirrors what the collector knows about its streams)
        self.seq = 0
        self.stream_last = np.full(self.max_streams, -1, dtype=np.int64)
        self.trunc_cursor = np.zeros(1, dtype=np.int64)
        self.trunc_owner = np.full(pool, -1, dtype=np.int64)
        self._alloc_staging()

    # ---------------------------------------------------------------------------
    def _make_desc(self):
        self._c = _lib.pb_store(
            obs=self.obs.data_ptr(), aux_obs=self.aux_obs.data_ptr(), action=self.action.data_ptr(),
            reward=self.reward.data_ptr(), done=self.done.data_ptr(), trunc=self.trunc.data_ptr(),
            slot_seq=self.slot_seq.data_ptr(), next_link=self.next_link.data_ptr(),
            prev_link=self.prev_link.data_ptr(), size=self.size, aux_size=self.aux_size,
            obs_elems=self.obs_elems, obs_dtype=0 if self.storage_dtype == torch.float32 else 1,
            obs_scale=int(self.obs_scale), frame_stack=self.frame_stack, n_step=self.n_step, pad=0,
            gamma=self.gamma)
        self._ref = C.byref(self._c)

    def _alloc_staging(self):
        S = self.staging_rows
        # rows: [2][S][obs_elems] pinned + device mirror; meta: S packed 64-byte pb_step_meta records
        self.h_rows = torch.zeros((2, S, self.obs_elems), dtype=self.storage_dtype).pin_memory()
        self.d_rows = torch.zeros((2, S, self.obs_elems), dtype=self.storage_dtype, device=self.device)
        self.h_meta = torch.zeros(S * 64, dtype=torch.uint8).pin_memory()
        self.d_meta = torch.zeros(S * 64, dtype=torch.uint8, device=self.device)
        self.meta = self.h_meta.numpy().view(np.dtype(_lib.STEP_META_DTYPE))
        assert self.meta.dtype.itemsize == 64 and self.meta.shape == (S,)
        self.h_rows_np = self.h_rows.numpy()
        self.m_stream = np.zeros(S, dtype=np.int32)
        self.m_flags = np.zeros(S, dtype=np.uint8)
        self.n_staged = 0

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def __len__(self):
        return min(self.seq, self.size)

    # ---------------------------------------------------------------------------
    def stage(self, stream_id, obs, action, reward, done, trunc, next_obs):
        """Stage ONE step on the pinned host block (no device work).  Returns True when
        the staging block is full and must be flushed."""
        j = self.n_staged
        if j == 0:
            self.wait_staging()
        row = self.h_rows_np[0, j]
        row[...] = np.asarray(obs, dtype=row.dtype).reshape(-1)
        flags = (STEP_DONE if done else 0) | (STEP_TRUNC if trunc else 0)
        if next_obs is None:
            if not done:
                flags |= STEP_NO_NEXT
        else:
            nrow = self.h_rows_np[1, j]
            nrow[...] = np.asarray(next_obs, dtype=nrow.dtype).reshape(-1)
        self.m_stream[j] = stream_id
        self.m_flags[j] = flags
        rec = self.meta[j]
        rec["action"] = int(action)
        rec["reward"] = float(reward)
        rec["done"] = 1 if done else 0
        rec["trunc"] = 1 if trunc else 0
        self.n_staged = j + 1
        return self.n_staged >= self.staging_rows

    def stage_batch(self, stream_ids, obs, action, reward, done, trunc, next_obs):
        """Stage up to staging_rows steps at once from host arrays; returns how many were taken."""
        j = self.n_staged
        n = min(len(stream_ids), self.staging_rows - j)
        if n <= 0:
            return 0
        if j == 0:
            self.wait_staging()
        sl = slice(j, j + n)
        self.h_rows_np[0, sl] = np.asarray(obs[:n]).reshape(n, -1)
        self.h_rows_np[1, sl] = np.asarray(next_obs[:n]).reshape(n, -1)
        d = np.asarray(done[:n]).astype(np.uint8)
        t = np.asarray(trunc[:n]).astype(np.uint8)
        self.m_stream[sl] = stream_ids[:n]
        self.m_flags[sl] = d * STEP_DONE + t * STEP_TRUNC
        self.meta["action"][sl] = action[:n]
        self.meta["reward"][sl] = reward[:n]
        self.meta["done"][sl] = d
        self.meta["trunc"][sl] = t
        self.n_staged = j + n
        return n

    def flush(self):
        """Plan links on the host, copy the staged block host->device (3 async copies from
        pinned memory) and scatter it into the ring.  Returns the number of steps written."""
        n = self.n_staged
        if n == 0:
            return 0
        while True:
            rc = self._lib.pb_store_extend_plan(
                self.size, self.aux_size, self.max_streams, n, self.seq, self.m_stream.ctypes.data,
                self.m_flags.ctypes.data, self.stream_last.ctypes.data, self.trunc_cursor.ctypes.data,
                self.trunc_owner.ctypes.data, self.h_meta.data_ptr())
            if rc == _lib.PB_E_POOL:
                self._grow_trunc_pool()
                continue
            _lib.check(rc, "pb_store_extend_plan")
            break
        self.d_rows[0, :n].copy_(self.h_rows[0, :n], non_blocking=True)
        self.d_rows[1, :n].copy_(self.h_rows[1, :n], non_blocking=True)
        self.d_meta[:n * 64].copy_(self.h_meta[:n * 64], non_blocking=True)
        # the pinned block is reused by the next stage(): remember when the copies are done
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._copy_event = ev
        _lib.check(self._lib.pb_store_scatter(self._ref, n, self.d_rows[0].data_ptr(), self.d_rows[1].data_ptr(),
                                              self.d_meta.data_ptr(), self._stream()), "pb_store_scatter")
        self.last_flush_h2d_bytes = n * (2 * self.obs_elems * self.h_rows.element_size() + 64)
        self.seq += n
        self.n_staged = 0
        return n

    def make_ingest_slot(self, n):
        return IngestSlot(self, n)

    def wait_staging(self):
        ev = getattr(self, "_copy_event", None)
        if ev is not None:
            ev.synchronize()
            self._copy_event = None

    def bytes_staged_per_row(self):
        return 2 * self.obs_elems * self.h_rows.element_size()

    def _grow_trunc_pool(self):
        old_pool = self.aux_size - self.max_streams
        new_pool = old_pool * 2
        aux = torch.zeros((self.max_streams + new_pool, self.obs_elems), dtype=self.storage_dtype, device=self.device)
        aux[:self.aux_size].copy_(self.aux_obs)
        # unroll the ring so that live rows keep their index and the cursor points at fresh rows
        owner = np.full(new_pool, -1, dtype=np.int64)
        owner[:old_pool] = self.trunc_owner
        self.trunc_owner = owner
        self.trunc_cursor[0] = old_pool
        self.aux_obs = aux
        self.aux_size = self.max_streams + new_pool
        self.generation += 1
        self._make_desc()

    # ---------------------------------------------------------------------------
    def gather(self, idx, obs_out, next_obs_out, ret_out, gamma_out, nonterm_out, action_out):
        n = idx.numel()
        row = self.frame_stack * self.obs_elems
        if (obs_out.numel() < n * row or next_obs_out.numel() < n * row or ret_out.numel() < n or gamma_out.numel() < n
                or nonterm_out.numel() < n or action_out.numel() < n):
            raise ValueError("gather of %d rows into a batch with %d rows" % (n, obs_out.numel() // max(row, 1)))
        _lib.check(self._lib.pb_store_gather(
            self._ref, n, idx.data_ptr(), obs_out.data_ptr(), next_obs_out.data_ptr(), ret_out.data_ptr(),
            gamma_out.data_ptr(), nonterm_out.data_ptr(), action_out.data_ptr(), self._stream()),
            "pb_store_gather")

    def nstep(self, idx):
        n = idx.numel()
        d = self.device
        ret = torch.empty(n, dtype=torch.float32, device=d)
        gam = torch.empty(n, dtype=torch.float32, device=d)
        done = torch.empty(n, dtype=torch.uint8, device=d)
        last = torch.empty(n, dtype=torch.int64, device=d)
        succ = torch.empty(n, dtype=torch.int64, device=d)
        _lib.check(self._lib.pb_store_nstep(self._ref, n, idx.data_ptr(), ret.data_ptr(), gam.data_ptr(),
                                            done.data_ptr(), last.data_ptr(), succ.data_ptr(), self._stream()),
                   "pb_store_nstep")
        return ret, gam, done, last, succ

    def clear(self):
        self.slot_seq.fill_(-1)
        self.next_link.fill_(-1)
        self.prev_link.fill_(-1)
        self.seq = 0
        self.stream_last[:] = -1
        self.trunc_cursor[0] = 0
        self.trunc_owner[:] = -1
        self.n_staged = 0

    def state_dict(self):
        n = len(self)
        return {"obs": self.obs[:n].cpu(), "aux_obs": self.aux_obs.cpu(), "action": self.action[:n].cpu(),
                "reward": self.reward[:n].cpu(), "done": self.done[:n].cpu(), "trunc": self.trunc[:n].cpu(),
                "slot_seq": self.slot_seq[:n].cpu(), "next_link": self.next_link[:n].cpu(),
                "prev_link": self.prev_link[:n].cpu(), "seq": self.seq, "stream_last": self.stream_last.copy(),
                "trunc_cursor": int(self.trunc_cursor[0]), "trunc_owner": self.trunc_owner.copy(),
                "size": self.size, "obs_shape": self.obs_shape}

    def load_state_dict(self, sd):
        assert sd["size"] == self.size and tuple(sd["obs_shape"]) == self.obs_shape
        n = sd["obs"].shape[0]
        while self.aux_size < sd["aux_obs"].shape[0]:
            self._grow_trunc_pool()
        self.clear()
        self.obs[:n].copy_(sd["obs"]); self.aux_obs[:sd["aux_obs"].shape[0]].copy_(sd["aux_obs"])
        self.action[:n].copy_(sd["action"]); self.reward[:n].copy_(sd["reward"])
        self.done[:n].copy_(sd["done"]); self.trunc[:n].copy_(sd["trunc"])
        self.slot_seq[:n].copy_(sd["slot_seq"]); self.next_link[:n].copy_(sd["next_link"])
        self.prev_link[:n].copy_(sd["prev_link"])
        self.seq = sd["seq"]
        # collector state cannot be recovered after a reload: every stream restarts
        # (the reference truncates in-flight trajectories on save for the same reason,
        # prism/experience/timestep_buffer.py:274-297)
        self.stream_last[:] = -1
        self.trunc_cursor[0] = sd["trunc_cursor"]
        self.trunc_owner[:len(sd["trunc_owner"])] = sd["trunc_owner"]

    def inflight_stream_rows(self):
        """Stream ids whose aux row still backs the successor observation of a stored tail step (next_link = -(sid+2)
        with sid < max_streams): after a reload these ids must not be handed to a new stream, or the first step of
        that stream would overwrite the row (the reference truncates such trajectories at save,
        prism/experience/timestep_buffer.py:274-297).  Synchronises; not on the hot path."""
        n = len(self)
        if n == 0:
            return []
        nl = self.next_link[:n]
        rows = (-nl[nl <= -2] - 2)
        rows = rows[rows < self.max_streams]
        return sorted(set(int(r) for r in rows.cpu().tolist()))


class IngestSlot:
    """Fixed-size ingest block (n steps per push) whose device half -- three pinned->device copies and the
    scatter kernel -- has static addresses, so it can be captured in a CUDA graph and replayed once per
    learner iteration (the reference collects `timesteps_per_iteration` = 4 steps per iteration).
    Host half: fill() writes the pinned block and runs the link planner."""

    def __init__(self, ring, n, n_uniforms=0):
        self.ring, self.n = ring, int(n)
        E, dt, dev = ring.obs_elems, ring.storage_dtype, ring.device
        esz = torch.empty(0, dtype=dt).element_size()
        rows_bytes = 2 * self.n * E * esz
        rows_bytes_al = (rows_bytes + 63) // 64 * 64
        # ONE pinned block [obs rows | next_obs rows | pb_step_meta records | fp64 uniforms] and its device mirror: one
        # H2D copy per iteration
        self.n_uniforms = int(n_uniforms)
        u_off = rows_bytes_al + self.n * 64
        self.h_block = torch.zeros(u_off + 8 * self.n_uniforms, dtype=torch.uint8).pin_memory()
        self.d_block = torch.zeros_like(self.h_block, device=dev)
        self.h_rows = self.h_block[:rows_bytes].view(dt).view(2, self.n, E)
        self.d_rows = self.d_block[:rows_bytes].view(dt).view(2, self.n, E)
        self.h_meta = self.h_block[rows_bytes_al:u_off]
        self.d_meta = self.d_block[rows_bytes_al:u_off]
        self.h_u = self.h_block[u_off:].view(torch.float64) if self.n_uniforms else None
        self.d_u = self.d_block[u_off:].view(torch.float64) if self.n_uniforms else None
        self.h_u_np = self.h_u.numpy() if self.n_uniforms else None
        self.meta = self.h_meta.numpy().view(np.dtype(_lib.STEP_META_DTYPE))
        self.rows_np = self.h_rows.numpy()
        self.stream = np.zeros(self.n, dtype=np.int32)
        self.flags = np.zeros(self.n, dtype=np.uint8)
        self.h2d_bytes = self.h_block.numel()
        # constants of the per-iteration native staging call
        self._row_dtype, self._rows_size = self.rows_np.dtype, self.rows_np[0].size
        self._row_nbytes, self._rows_ptr, self._meta_ptr = self.rows_np[0, 0].nbytes, self.rows_np.ctypes.data, self.h_meta.data_ptr()

    def _fill_native(self, stream_ids, obs, action, reward, done, trunc, next_obs):
        """One host call (pb_store_stage_block) when every array already has the staged layout and dtype; returns
        False (nothing touched) otherwise.  This runs once per learner iteration on the end-to-end path: the checks are
        plain attribute reads against constants prepared in __init__."""
        ring, n = self.ring, self.n
        nd = np.ndarray
        try:
            if not (type(obs) is nd and type(next_obs) is nd and type(stream_ids) is nd and type(action) is nd
                    and type(reward) is nd and type(done) is nd and type(trunc) is nd):
                return False
            row_dt, row_sz = self._row_dtype, self._rows_size
            if not (obs.dtype == row_dt and obs.size == row_sz and obs.flags.c_contiguous
                    and next_obs.dtype == row_dt and next_obs.size == row_sz and next_obs.flags.c_contiguous
                    and stream_ids.dtype == _I32 and stream_ids.size == n and stream_ids.flags.c_contiguous
                    and (action.dtype == _I64 or action.dtype == _I32) and action.size == n and action.flags.c_contiguous
                    and reward.dtype == _F32 and reward.size == n and reward.flags.c_contiguous
                    and (done.dtype == _BOOL or done.dtype == _U8) and done.size == n and done.flags.c_contiguous
                    and (trunc.dtype == _BOOL or trunc.dtype == _U8) and trunc.size == n and trunc.flags.c_contiguous):
                return False
        except AttributeError:
            return False
        rc = ring._lib.pb_store_stage_block(
            ring.size, ring.aux_size, ring.max_streams, n, ring.seq, self._row_nbytes, obs.ctypes.data,
            next_obs.ctypes.data, stream_ids.ctypes.data, action.ctypes.data, action.itemsize, reward.ctypes.data, done.ctypes.data,
            trunc.ctypes.data, self._rows_ptr, ring.stream_last.ctypes.data, ring.trunc_cursor.ctypes.data,
            ring.trunc_owner.ctypes.data, self._meta_ptr)
        if rc == _lib.PB_E_POOL:
            raise _lib.PbError("truncated-observation pool exhausted under a captured ingest graph: "
                               "construct the ring with a larger trunc_pool")
        if rc:
            _lib.check(rc, "pb_store_stage_block")
        ring.seq += n
        return True

    def fill(self, stream_ids, obs, action, reward, done, trunc, next_obs):
        if NATIVE_STAGE and self._fill_native(stream_ids, obs, action, reward, done, trunc, next_obs):
            return
        ring, n = self.ring, self.n
        self.rows_np[0] = np.asarray(obs).reshape(n, -1)
        self.rows_np[1] = np.asarray(next_obs).reshape(n, -1)
        d = np.asarray(done).astype(np.uint8)
        t = np.asarray(trunc).astype(np.uint8)
        self.stream[:] = stream_ids
        self.flags[:] = d * STEP_DONE + t * STEP_TRUNC
        m = self.meta
        m["action"] = action
        m["reward"] = reward
        m["done"] = d
        m["trunc"] = t
        rc = ring._lib.pb_store_extend_plan(
            ring.size, ring.aux_size, ring.max_streams, n, ring.seq, self.stream.ctypes.data, self.flags.ctypes.data,
            ring.stream_last.ctypes.data, ring.trunc_cursor.ctypes.data, ring.trunc_owner.ctypes.data,
            self.h_meta.data_ptr())
        if rc == _lib.PB_E_POOL:
            raise _lib.PbError("truncated-observation pool exhausted under a captured ingest graph: "
                               "construct the ring with a larger trunc_pool")
        _lib.check(rc, "pb_store_extend_plan")
        ring.seq += n

    def enqueue(self):
        """Device half (capturable): H2D of the block, then the scatter kernel."""
        ring = self.ring
        self.d_block.copy_(self.h_block, non_blocking=True)
        _lib.check(ring._lib.pb_store_scatter(ring._ref, self.n, self.d_rows[0].data_ptr(), self.d_rows[1].data_ptr(),
                                              self.d_meta.data_ptr(), ring._stream()), "pb_store_scatter")


class FusedIngest:
    """A fixed-size ingest (n steps per learner iteration) whose device half runs INSIDE the learner's step graph,
    on the branch that follows the priority write-back, i.e. concurrently with backward / Adam -- the reference
    interleaves `timesteps_per_iteration` new steps with every update (learner.py:95-125, collector drain) and pays
    for them on the same thread.

    Two staging blocks (pinned + device).  Host side, per iteration (``stage``): wait until the pinned block of this
    parity is free, write it, run the link planner, and enqueue its H2D copy on a copy stream (ordered after the
    graph replay that last read the device block).  Device side (``consume``, captured): the scatter kernel picks the
    block by the parity of a device-resident replay counter, then the trees get the default priorities.
    Steps staged before iteration k's launch become sampleable from iteration k+1 on."""

    def __init__(self, ring, tree, n, n_uniforms=0):
        import ctypes as C
        self.ring, self.tree, self.n = ring, tree, int(n)
        self.n_uniforms = int(n_uniforms)
        self.slots = [IngestSlot(ring, n, n_uniforms), IngestSlot(ring, n, n_uniforms)]
        dev = ring.device
        self.counter = torch.zeros(1, dtype=torch.int64, device=dev)
        self.copy_stream = torch.cuda.Stream(device=dev)
        self._copy_handle = self.copy_stream.cuda_stream
        lib = self._lib = ring._lib
        # raw events (one ctypes call per operation on the per-iteration path)
        self.copied, self.consumed = [], []     # H2D of block p finished / the replay that read device block p finished
        for _ in range(2):
            for lst in (self.copied, self.consumed):
                ev = C.c_void_p()
                _lib.check(lib.pb_event_create(C.byref(ev)), "pb_event_create")
                lst.append(ev)
        self._copied_valid, self._consumed_valid = [False, False], [False, False]
        self._ptrs = [(s.d_block.data_ptr(), s.h_block.data_ptr(), s.h_block.numel()) for s in self.slots]
        self.calls = 0
        self.h2d_bytes = self.slots[0].h2d_bytes

    def select_uniforms(self, dst, after_counter_inc):
        """Capturable: copy this replay's uniforms out of its staging block into ``dst`` (device fp64).
        ``after_counter_inc``: whether the scatter (which bumps the replay counter) already ran in stream order."""
        a, b = self.slots
        _lib.check(self._lib.pb_select_copy_f64(dst.data_ptr(), a.d_u.data_ptr(), b.d_u.data_ptr(), self.counter.data_ptr(),
                                                -1 if after_counter_inc else 0, self.n_uniforms, self.ring._stream()),
                   "pb_select_copy_f64")

    def stage(self, stream_ids, obs, action, reward, done, trunc, next_obs, main_stream, u=None):
        """Host half: fill the pinned block of this parity, copy it on the copy stream (after the replay that last
        read the device block), make ``main_stream`` (raw handle) wait for the copy.  Returns the parity."""
        lib = self._lib
        p = self.calls & 1
        self.calls += 1
        if self._copied_valid[p]:
            lib.pb_event_synchronize(self.copied[p])          # pinned block free again (two iterations old)
        self.slots[p].fill(stream_ids, obs, action, reward, done, trunc, next_obs)
        if u is not None:
            self.slots[p].h_u_np[:] = u
        dst, src, nbytes = self._ptrs[p]
        rc = lib.pb_staged_copy_submit(self._copy_handle, self.consumed[p] if self._consumed_valid[p] else None, dst, src,
                                       nbytes, self.copied[p], main_stream)
        if rc:
            _lib.check(rc, "pb_staged_copy_submit")
        self._copied_valid[p] = True
        return p

    def plan_block(self, stream_ids, obs, action, reward, done, trunc, next_obs, u=None):
        """Host half WITHOUT the copy: run the link planner for the next n steps and return the staged block's bytes
        (a pinned uint8 tensor: rows | step records | uniforms).  Blocks planned in order can be parked anywhere
        (e.g. in HBM) and fed back in the same order with ``stage_device``."""
        p = self._planned & 1 if hasattr(self, "_planned") else 0
        self._planned = getattr(self, "_planned", 0) + 1
        if self._copied_valid[p]:
            self._lib.pb_event_synchronize(self.copied[p])
        self.slots[p].fill(stream_ids, obs, action, reward, done, trunc, next_obs)
        if u is not None:
            self.slots[p].h_u_np[:] = u
        return self.slots[p].h_block.clone()

    def stage_device(self, block, main_stream):
        """Device-resident variant of ``stage``: ``block`` (uint8 CUDA tensor from ``plan_block``) is copied device to
        device into the staging block of this parity on the step's own stream (ordered after the replay that last
        read it).  Returns the parity."""
        p = self.calls & 1
        self.calls += 1
        dst, _, nbytes = self._ptrs[p]
        if block.numel() != nbytes or not block.is_cuda:
            raise _lib.PbError("stage_device needs the %d-byte CUDA block plan_block produced" % nbytes)
        _lib.check(self._lib.pb_copy_d2d_async(dst, block.data_ptr(), nbytes, main_stream), "pb_copy_d2d_async")
        return p

    def mark_consumed(self, p, main_stream):
        self._lib.pb_event_record(self.consumed[p], main_stream)
        self._consumed_valid[p] = True

    def consume(self):
        """Device half (capturable): scatter from the block of this replay's parity, then default priorities."""
        self.consume_ring()
        self.consume_tree()

    def consume_tree(self):
        if self.tree is not None:
            self.tree.extend(self.n)

    def consume_ring(self):
        """The ring half only (touches no tree array: may run beside the priority write-back)."""
        ring, a, b = self.ring, self.slots[0], self.slots[1]
        _lib.check(ring._lib.pb_store_scatter_dbuf(ring._ref, self.n, a.d_rows[0].data_ptr(), a.d_rows[1].data_ptr(),
                                                   a.d_meta.data_ptr(), b.d_rows[0].data_ptr(), b.d_rows[1].data_ptr(),
                                                   b.d_meta.data_ptr(), self.counter.data_ptr(), ring._stream()),
                   "pb_store_scatter_dbuf")
