"""Drop-in ``TimestepBuffer`` backed by the device ring + device sum/min trees.

Same surface as the reference class (prism/experience/timestep_buffer.py:10-77):
``extend / sample / update_priority / set_static_batch / get_static_batch / empty / save /
load`` and the attribute paths the learner pokes (``buffer.buffer._sampler._beta`` at
prism/learner.py:107, ``buffer.buffer._batch_size`` at timestep_buffer.py:44).

What changed underneath (B200-first):
  * transitions live in HBM (SoA ring), not in Python objects; ``extend`` only stages the
    step on pinned host memory, the staged block is scattered by one kernel at the next
    sample()/update_priority() (same program order as the reference);
  * sample = one tree-descent kernel + one fused n-step/gather kernel writing straight
    into the agent's static batch; no host loop, no unpinned H2D copies;
  * update_priority takes device tensors and never synchronises (the reference's torchrl
    call forces a D2H copy of the TD errors every iteration).
"""
import os
import pickle

import numpy as np
import torch

from .. import _lib
from .batch import Batch
from .per import PrioritizedTree
from .ring import TransitionRing


class _Writer:
    """Round-robin writer state (torchrl RoundRobinWriter stand-in)."""

    def __init__(self, owner):
        self._owner = owner

    @property
    def _cursor(self):
        ring = self._owner._storage
        return 0 if ring is None else ring.seq % ring.size


class DevicePrioritizedReplayBuffer:
    """Stand-in for ``torchrl.data.PrioritizedReplayBuffer(storage=ListStorage(N), alpha, beta,
    batch_size)`` as built at prism/factory/exp_buffer_factory.py:22-28.  ``prioritized=False``
    gives the uniform ``ReplayBuffer`` of :29-33."""

    def __init__(self, capacity, alpha=0.5, beta=0.5, batch_size=None, eps=1e-8, device="cuda:0",
                 prioritized=True, sampling="iid", storage_dtype=torch.float32, obs_scale=False,
                 max_streams=256, staging_rows=256, weight_eps_in_denominator=False,
                 default_priority_fp64=True):
        _lib.load()  # fail loudly if the CUDA library is missing
        self.device = torch.device(device)
        self.capacity = int(capacity)
        self._batch_size = batch_size
        self.prioritized = bool(prioritized)
        self.sampling = sampling
        self._storage = None  # TransitionRing, created at the first extend (needs the obs shape)
        self._storage_opts = dict(storage_dtype=storage_dtype, obs_scale=obs_scale, max_streams=max_streams,
                                  staging_rows=staging_rows)
        self._sampler = PrioritizedTree(self.capacity, alpha=alpha, beta=beta, eps=eps, device=device,
                                        mode=sampling, weight_eps_in_denominator=weight_eps_in_denominator,
                                        default_priority_fp64=default_priority_fp64) if prioritized else None
        self._writer = _Writer(self)

    def __len__(self):
        return 0 if self._storage is None else len(self._storage)


class TimestepBuffer(object):
    def __init__(self, torchrl_buffer, frame_stack=1, device="cuda:0", n_step=3, gamma=0.99):
        if not isinstance(torchrl_buffer, DevicePrioritizedReplayBuffer):
            raise TypeError("prism_b200.TimestepBuffer wraps a DevicePrioritizedReplayBuffer "
                            "(see prism_b200.factory.build_exp_buffer); got %r" % type(torchrl_buffer))
        self.buffer = torchrl_buffer
        self.device = device
        if torch.device(device).type != "cuda":
            raise _lib.PbError("prism_b200.TimestepBuffer is device-resident: device must be CUDA, got %s" % device)
        self.frame_stack = frame_stack
        self.n_step = n_step
        self.gamma = gamma
        self.gammas = [gamma ** i for i in range(n_step + 1)]
        self._batch = None
        self._obs = self._next_obs = self._reward = self._nonterminal = self._gamma = self._action = None
        self._idx = None
        self._weight = None
        self._free_streams = None
        self._stream_tail = {}
        self._last_sorted = False
        self._injected_u = None

    # ------------------------------------------------------------------ ingest
    def _ring(self, obs_shape=None):
        ring = self.buffer._storage
        if ring is None:
            if obs_shape is None:
                return None
            ring = TransitionRing(self.buffer.capacity, obs_shape, frame_stack=self.frame_stack, n_step=self.n_step,
                                  gamma=self.gamma, device=self.device, **self.buffer._storage_opts)
            self.buffer._storage = ring
            self._free_streams = list(range(ring.max_streams - 1, -1, -1))
        return ring

    def _stream_for(self, timestep):
        # the stream id rides on the newest Timestep object of each chain (dataclass objects are
        # unhashable, so an attribute rather than a dict key)
        prev = timestep.prev() if timestep.prev is not None else None
        sid = None
        if prev is not None:
            sid = getattr(prev, "_pb_stream", None)
            if sid is not None:
                prev._pb_stream = None
        if sid is None:
            if not self._free_streams:
                raise _lib.PbError("more concurrent collector streams than max_streams=%d" %
                                   self.buffer._storage.max_streams)
            sid = self._free_streams.pop()
        return sid

    def extend(self, timestep):
        """Stage one completed step (reference: timestep_buffer.py:32-33).  The step is linked to
        its stream through ``timestep.prev``; its successor observation is read from
        ``timestep.next`` (the in-flight step, or the truncated final observation)."""
        obs = timestep.obs
        ring = self._ring(tuple(obs.shape))
        sid = self._stream_for(timestep)
        nxt = timestep.next
        if nxt is not None and not hasattr(nxt, "obs"):
            nxt = nxt()  # weakref to the in-flight step
        next_obs = None if (nxt is None or timestep.done) else nxt.obs
        done, trunc = bool(timestep.done), bool(timestep.truncated)
        full = ring.stage(sid, _to_numpy(obs), timestep.action, timestep.reward, done, trunc,
                          None if next_obs is None else _to_numpy(next_obs))
        if done or trunc:
            self._free_streams.append(sid)
            self._stream_tail.pop(sid, None)
        else:
            timestep._pb_stream = sid
            # the reference's ListStorage keeps every stored Timestep alive, so the collector's weak
            # `prev` link of the NEXT step resolves; here only the newest step of a stream must stay alive
            self._stream_tail[sid] = timestep
        if full:
            self._flush()
        return ring.seq + ring.n_staged - 1

    def extend_batch(self, stream_ids, obs, action, reward, done, trunc, next_obs):
        """Batched ingest: n steps in arrival order, each tagged with its collector
        stream.  Host arrays; staged through pinned memory in blocks of ``staging_rows``."""
        obs = np.asarray(obs)
        ring = self._ring(tuple(obs.shape[1:]))
        stream_ids = np.asarray(stream_ids, dtype=np.int32)
        n, off = len(stream_ids), 0
        while off < n:
            took = ring.stage_batch(stream_ids[off:], obs[off:], np.asarray(action)[off:], np.asarray(reward)[off:],
                                    np.asarray(done)[off:], np.asarray(trunc)[off:], np.asarray(next_obs)[off:])
            off += took
            if ring.n_staged >= ring.staging_rows:
                self._flush()
        return n

    def ingest_graph(self, n):
        """A replayable ingest of exactly ``n`` steps per call (H2D + scatter + default priorities as ONE
        CUDA graph): ``push = buffer.ingest_graph(4); push(stream_ids, obs, action, reward, done, trunc,
        next_obs)``.  The host only writes pinned memory, plans links and replays."""
        return _IngestGraph(self, n)

    def _flush(self):
        ring = self.buffer._storage
        if ring is None or ring.n_staged == 0:
            return 0
        n = ring.flush()
        if self.buffer._sampler is not None:
            self.buffer._sampler.extend(n)
        self._mutations = getattr(self, "_mutations", 0) + 1      # LearnerStep(prefetch=True) re-samples after outside changes
        return n

    # ------------------------------------------------------------------ sampling
    def inject_uniforms(self, u):
        """Use these fp64 uniforms (device or pinned-host tensor) for the next sample() instead of
        the device generator -- how parity tests and the e2e bench feed identical randomness."""
        self._injected_u = u

    @torch.no_grad()
    def sample(self, batch_size=None, return_info=False):
        self._flush()
        ring = self.buffer._storage
        if ring is None or len(ring) == 0:
            raise RuntimeError("Cannot sample from an empty storage.")  # torchrl's _EMPTY_STORAGE_ERROR
        if batch_size is None:
            batch_size = self.buffer._batch_size
        B = int(batch_size)
        if self._batch is None:
            self._alloc_static_batch(B, ring)
            self._own_batch = True
        elif B > self._obs.shape[0]:
            # more rows than the static batch holds: grow a batch this buffer allocated itself, refuse to write past one
            # handed over by the agent (set_static_batch) -- the reference raises there as well (copy_ shape mismatch)
            if not getattr(self, "_own_batch", False):
                raise ValueError("sample(batch_size=%d) exceeds the static batch of %d rows" % (B, self._obs.shape[0]))
            self._alloc_static_batch(B, ring)
            self._own_batch = True
        if self._idx is None or self._idx.numel() != B:
            self._idx = torch.empty(B, dtype=torch.int64, device=self.device)
            self._weight = torch.ones(B, dtype=torch.float32, device=self.device)
        u, self._injected_u = self._injected_u, None
        tree = self.buffer._sampler
        if tree is not None:
            tree.sample(B, u=u, idx_out=self._idx, weight_out=self._weight)
            self._last_sorted = tree.mode == PrioritizedTree.MODE_STRATIFIED
        else:
            if u is None:
                u = torch.rand(B, dtype=torch.float64, device=self.device)
            else:
                u = torch.as_tensor(u).to(self.device, non_blocking=True)
            torch.mul(u, float(len(ring)), out=u)
            self._idx.copy_(u.long().clamp_(max=len(ring) - 1))
            self._last_sorted = False
        ring.gather(self._idx, self._obs, self._next_obs, self._reward, self._gamma, self._nonterminal, self._action)
        if return_info:
            info = {"index": self._idx}
            if tree is not None:
                info["_weight"] = self._weight
            return self._batch, info
        return self._batch

    def update_priority(self, indices, priorities):
        """Reference: timestep_buffer.py:53-54 (learner.py:119-120).  Device tensors, no sync."""
        if self.buffer._sampler is None:
            return
        self._flush()
        self._mutations = getattr(self, "_mutations", 0) + 1
        sorted_hint = self._last_sorted and isinstance(indices, torch.Tensor) and self._idx is not None \
            and indices.data_ptr() == self._idx.data_ptr()
        self.buffer._sampler.update_priority(indices, priorities, sorted=sorted_hint)

    # ------------------------------------------------------------------ static batch
    def _alloc_static_batch(self, batch_size, ring):
        d = self.device
        fs = self.frame_stack
        batch = Batch({
            "observation": torch.zeros(batch_size, fs, *ring.obs_shape, dtype=torch.float32, device=d),
            "next": Batch({
                "observation": torch.zeros(batch_size, fs, *ring.obs_shape, dtype=torch.float32, device=d),
                "reward": torch.zeros(batch_size, 1, dtype=torch.float32, device=d)},
                batch_size=batch_size, device=d),
            "nonterminal": torch.zeros(batch_size, 1, dtype=torch.bool, device=d),
            "gamma": torch.ones(batch_size, 1, dtype=torch.float32, device=d),
            "action": torch.zeros(batch_size, 1, dtype=torch.long, device=d),
        }, batch_size=batch_size, device=d)
        self.set_static_batch(batch)

    def set_static_batch(self, batch):
        self._own_batch = False
        self._batch = batch
        self._obs = batch["observation"]
        self._next_obs = batch["next"]["observation"]
        self._reward = batch["next"]["reward"]
        self._nonterminal = batch["nonterminal"]
        self._gamma = batch["gamma"]
        self._action = batch["action"]
        for name, t, dt in (("observation", self._obs, torch.float32), ("next.observation", self._next_obs, torch.float32),
                            ("next.reward", self._reward, torch.float32), ("nonterminal", self._nonterminal, torch.bool),
                            ("gamma", self._gamma, torch.float32), ("action", self._action, torch.int64)):
            if not (t.is_cuda and t.is_contiguous() and t.dtype == dt):
                raise _lib.PbError("static batch tensor %r must be a contiguous CUDA %s tensor" % (name, dt))

    def get_static_batch(self):
        return self._batch

    # ------------------------------------------------------------------ misc
    def empty(self):
        ring = self.buffer._storage
        if ring is not None:
            ring.clear()
            self._free_streams = list(range(ring.max_streams - 1, -1, -1))
            self._stream_tail = {}
        if self.buffer._sampler is not None:
            self.buffer._sampler.reset()

    def __len__(self):
        ring = self.buffer._storage
        return 0 if ring is None else min(ring.seq + ring.n_staged, ring.size)

    def save(self, path):
        """Checkpoint the device buffer (tensors + cursors + trees).  Native format; the
        reference's pickle layout (timestep_buffer.py:259-302) is dead code upstream
        (checkpointer.py:23-24 returns early)."""
        self._flush()
        buffer_path = os.path.join(path, "experience_buffer")
        os.makedirs(buffer_path, exist_ok=True)
        ring = self.buffer._storage
        sd = {"ring": None if ring is None else ring.state_dict(),
              "tree": None if self.buffer._sampler is None else self.buffer._sampler.state_dict()}
        torch.save(sd, os.path.join(buffer_path, "device_buffer.pt"))

    def load_reference(self, path):
        """Read a checkpoint written by the REFERENCE's ``TimestepBuffer.save`` (timestep_buffer.py:259-302):
        ``experience_buffer/timesteps.pkl``, the stored steps as one flat ``Timestep.serialize`` list in ring-slot order,
        unfinished trajectories artificially truncated by the writer.  The steps are re-ingested in id (creation)
        order through the batched path.  torchrl's sampler / writer dumps next to it are not read: priorities restart
        at the default (max) priority, as for freshly collected steps.  Returns the number of steps loaded."""
        from ..async_components import wire
        with open(os.path.join(path, "experience_buffer", "timesteps.pkl"), "rb") as f:
            serialized = _NumbersOnlyUnpickler(f).load()
        flat = np.asarray([wire.NULL_VALUE if v is None else v for v in serialized], dtype=np.float64)
        decoder = wire.TimestepWireDecoder(max_streams=self.buffer._storage_opts.get("max_streams", 256))
        rows = decoder.feed(flat, by_id=True)
        if rows is None:
            return 0
        self.empty()
        n = self.extend_batch(*rows)
        self._flush()
        # the decoder assigned the stream ids: hand back the ones of finished episodes, keep the live ones reserved
        ring = self.buffer._storage
        live = set(decoder._tail.values())
        self._free_streams = [s for s in range(ring.max_streams - 1, -1, -1) if s not in live]
        return n

    def load(self, path):
        buffer_path = os.path.join(path, "experience_buffer")
        native = os.path.join(buffer_path, "device_buffer.pt")
        if not os.path.exists(native) and os.path.exists(os.path.join(buffer_path, "timesteps.pkl")):
            # a checkpoint directory written by the reference's own TimestepBuffer.save
            return self.load_reference(path)
        sd = torch.load(native, weights_only=False)
        if sd["ring"] is not None:
            ring = self._ring(tuple(sd["ring"]["obs_shape"]))
            ring.load_state_dict(sd["ring"])
            # tails of episodes that were in flight at save time keep their aux row: those stream ids stay reserved
            reserved = set(ring.inflight_stream_rows())
            self._free_streams = [s for s in range(ring.max_streams - 1, -1, -1) if s not in reserved]
            self._stream_tail = {}
        if sd["tree"] is not None and self.buffer._sampler is not None:
            self.buffer._sampler.load_state_dict(sd["tree"])


class _NumbersOnlyUnpickler(pickle.Unpickler):
    """timesteps.pkl is a flat list of Python numbers: refuse anything that would import a class."""

    def find_class(self, module, name):
        raise pickle.UnpicklingError("reference checkpoint may only hold numbers, found %s.%s" % (module, name))


def _to_numpy(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


class _IngestGraph:
    def __init__(self, buffer, n):
        ring = buffer.buffer._storage
        if ring is None:
            raise _lib.PbError("ingest_graph needs the ring to exist (extend at least one step first)")
        self.buffer, self.ring, self.n = buffer, ring, int(n)
        self.slot = ring.make_ingest_slot(n)
        self.tree = buffer.buffer._sampler
        self.done = None
        dev = ring.device
        buffer._flush()
        # the warm-up below really ingests: feed it a harmless terminal step per row on stream 0 ... no:
        # warm up the copies only (no state change), then capture copies + scatter + tree extend
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self.slot.d_block.copy_(self.slot.h_block, non_blocking=True)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._capture()
        self.h2d_bytes = self.slot.h2d_bytes

    def _capture(self):
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.slot.enqueue()
            if self.tree is not None:
                self.tree.extend(self.n)
        self._generation = self.ring.generation

    def __call__(self, stream_ids, obs, action, reward, done, trunc, next_obs):
        self.buffer._flush()
        if self.done is not None:
            self.done.synchronize()          # the previous replay has consumed the pinned block
        if self.ring.generation != self._generation:
            # the flush above grew the aux pool: the captured scatter holds the old descriptor by value
            torch.cuda.synchronize(self.ring.device)
            self._capture()
        self.slot.fill(stream_ids, obs, action, reward, done, trunc, next_obs)
        self.buffer._mutations = getattr(self.buffer, "_mutations", 0) + 1
        self.graph.replay()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.ring.device))
        self.done = ev
