"""Flat-list wire formats of the Redis transport and their columnar decoders.

Two payloads cross the wire, both as ONE flat list of numbers:

* a block of collector steps, each laid out by ``Timestep.serialize`` (prism/experience/timestep.py:30-101):
  ``id | n_obs, obs..., n_shape, shape... (or NULL) | truncated successor: id, n_obs, obs..., n_shape, shape... (or
  NULL) | reward, done, truncated, action, n_step_return, n_step_gamma, n_step_done, needs_n_step, episodic_reward |
  n_step_next id, prev id, next id`` with ``None -> NULL_VALUE``;
* a training batch: six tensors, each ``len(shape), *shape, n_elements, *values``
  (prism/async_components/async_experience_buffer.py:76-96).

The reference rebuilds a linked list of Python ``Timestep`` objects from the first (timestep.py:103-262) and feeds them
to the buffer one by one.  ``TimestepWireDecoder`` turns the same block into the arrays of the device buffer's batched
ingest (``TimestepBuffer.extend_batch``) instead: records are indexed by the native codec (csrc/wire.cu), observations
are sliced as array views, and the prev / next ids are resolved into collector streams -- a step is released once its
successor's observation is known (the successor's own record, or the truncated observation it carries), which is the
same rule the reference's link resolution applies through ``next``.
"""
import ctypes as C

import numpy as np

from .. import _lib

NULL_VALUE = -1313                       # prism/experience/timestep.py:9
_REC_W = 12
# offsets inside a record's 12 scalar fields (timestep.py:58-96)
_REWARD, _DONE, _TRUNC, _ACTION, _NSR, _NSG, _NSD, _NEEDS, _EPI, _NSN_ID, _PREV_ID, _NEXT_ID = range(12)


def _deref(link):
    """Timestep links are weak references, except ``next`` of a truncated step (a plain object)."""
    if link is None:
        return None
    return link if hasattr(link, "id") else link()


def _obs_fields(obs):
    a = obs.detach().cpu().numpy() if hasattr(obs, "detach") else np.asarray(obs)
    return [int(a.size)] + a.reshape(-1).tolist() + [a.ndim] + [int(s) for s in a.shape]


def serialize_timestep(ts):
    """The reference's ``Timestep.serialize`` (timestep.py:30-101): a flat Python list."""
    out = [ts.id]
    out += _obs_fields(ts.obs) if ts.obs is not None else [NULL_VALUE]
    if ts.truncated:
        succ = _deref(ts.next)
        out.append(succ.id)
        out += _obs_fields(succ.obs)
    else:
        out.append(NULL_VALUE)
    out += [ts.reward, ts.done, ts.truncated]
    action = ts.action
    if action is not None and hasattr(action, "item"):
        action = action.item()
    out.append(NULL_VALUE if action is None else action)
    out += [ts.n_step_return, ts.n_step_gamma, ts.n_step_done, ts.needs_n_step, ts.episodic_reward]
    for link in (ts.n_step_next, ts.prev, ts.next):
        target = _deref(link)
        out.append(NULL_VALUE if target is None else target.id)
    return [NULL_VALUE if v is None else v for v in out]


def _scalar(v):
    """One wire number with the type msgpack would see: None -> NULL, bool / int / float kept apart."""
    if v is None:
        return np.asarray(NULL_VALUE, dtype=np.int64)
    if isinstance(v, (bool, np.bool_)):
        return np.asarray(v, dtype=np.bool_)
    if isinstance(v, (int, np.integer)):
        return np.asarray(v, dtype=np.int64)
    return np.asarray(v, dtype=np.float64)


def _obs_segments(obs):
    a = obs.detach().cpu().numpy() if hasattr(obs, "detach") else np.asarray(obs)
    return [np.asarray(a.size, dtype=np.int64), a, np.asarray([a.ndim, *a.shape], dtype=np.int64)]


def timestep_segments(timesteps):
    """The segments (for ``pack_numbers``) of a block of steps: the same bytes as packing the concatenated
    ``serialize_timestep`` lists, without expanding observations into Python floats."""
    null = np.asarray(NULL_VALUE, dtype=np.int64)
    segs = []
    for ts in timesteps:
        segs.append(np.asarray(ts.id, dtype=np.int64))
        segs += _obs_segments(ts.obs) if ts.obs is not None else [null]
        if ts.truncated:
            succ = _deref(ts.next)
            segs.append(np.asarray(succ.id, dtype=np.int64))
            segs += _obs_segments(succ.obs)
        else:
            segs.append(null)
        action = ts.action
        if action is not None and hasattr(action, "item"):
            action = action.item()
        for v in (ts.reward, ts.done, ts.truncated, action, ts.n_step_return, ts.n_step_gamma, ts.n_step_done,
                  ts.needs_n_step, ts.episodic_reward):
            segs.append(_scalar(v))
        for link in (ts.n_step_next, ts.prev, ts.next):
            target = _deref(link)
            segs.append(null if target is None else np.asarray(target.id, dtype=np.int64))
    return segs


def index_timesteps(flat):
    """(records, 12) int64 table of a decoded block -- columns documented at pb_wire_index_timesteps."""
    lib = _lib.load()
    flat = np.ascontiguousarray(flat, dtype=np.float64)
    n_rec = C.c_longlong(0)
    _lib.check(lib.pb_wire_index_timesteps(flat.ctypes.data, flat.size, 0, None, C.byref(n_rec)),
               "pb_wire_index_timesteps")
    rec = np.empty((max(n_rec.value, 1), _REC_W), dtype=np.int64)
    _lib.check(lib.pb_wire_index_timesteps(flat.ctypes.data, flat.size, rec.shape[0], rec.ctypes.data, C.byref(n_rec)),
               "pb_wire_index_timesteps")
    return flat, rec[:n_rec.value]


def _obs_view(flat, off, n, shape_off, shape_n):
    shape = tuple(int(s) for s in flat[shape_off:shape_off + shape_n])
    return flat[off:off + n].reshape(shape)


class TimestepWireDecoder(object):
    """Stateful: feed it the decoded number blocks in arrival order; it returns the steps that became complete as the
    argument tuple of ``TimestepBuffer.extend_batch`` (or None).  ``max_streams``: the device ring's stream budget
    (concurrent unfinished episodes on the wire)."""

    def __init__(self, max_streams=256):
        self.max_streams = int(max_streams)
        self._free = list(range(self.max_streams - 1, -1, -1))
        self._tail = {}            # id of the newest step of a live stream -> stream id
        self._waiting = {}         # id of the awaited successor -> the held step (stream, obs, action, reward)
        self.n_decoded = 0

    @property
    def n_waiting(self):
        return len(self._waiting)

    def feed(self, flat, by_id=False):
        """``by_id``: take the records in increasing id order instead of block order (a checkpoint lists them in ring
        slot order; ids are the collectors' creation counter, so within a stream they follow arrival order)."""
        flat, rec = index_timesteps(flat)
        if len(rec) == 0:
            return None
        tails = flat[rec[:, 10, None] + np.arange(12)[None, :]]                 # (records, 12) scalar fields
        rows = []                                                               # (stream, obs, action, reward, done, trunc, next_obs)
        for r in (np.argsort(rec[:, 0], kind="stable") if by_id else range(len(rec))):
            c, t = rec[r], tails[r]
            ts_id = int(c[0])
            if c[1] < 0:
                raise _lib.PbError("step %d arrived without an observation" % ts_id)
            obs = _obs_view(flat, c[1], c[2], c[3], c[4])
            held = self._waiting.pop(ts_id, None)
            if held is not None:                                                # this record completes its predecessor
                rows.append(held + (False, False, obs))
            prev_id = int(t[_PREV_ID])
            stream = self._tail.pop(prev_id, None) if prev_id != NULL_VALUE else None
            if stream is None:
                if not self._free:
                    raise _lib.PbError("more concurrent collector streams on the wire than max_streams=%d" % self.max_streams)
                stream = self._free.pop()
            done, trunc = bool(t[_DONE] != 0), bool(t[_TRUNC] != 0)
            action = 0 if t[_ACTION] == NULL_VALUE else int(t[_ACTION])
            step = (stream, obs, action, float(t[_REWARD]))
            if trunc and c[6] >= 0:
                rows.append(step + (done, True, _obs_view(flat, c[6], c[7], c[8], c[9])))
                self._free.append(stream)
            elif done:
                rows.append(step + (True, trunc, np.zeros_like(obs)))
                self._free.append(stream)
            else:
                next_id = int(t[_NEXT_ID])
                if next_id == NULL_VALUE:
                    raise _lib.PbError("step %d is neither terminal nor linked to a successor" % ts_id)
                self._waiting[next_id] = step
                self._tail[ts_id] = stream
        self.n_decoded += len(rec)
        if not rows:
            return None
        return (np.asarray([r[0] for r in rows], dtype=np.int32),
                np.stack([r[1] for r in rows]).astype(np.float32),
                np.asarray([r[2] for r in rows], dtype=np.int64),
                np.asarray([r[3] for r in rows], dtype=np.float32),
                np.asarray([r[4] for r in rows], dtype=np.bool_),
                np.asarray([r[5] for r in rows], dtype=np.bool_),
                np.stack([r[6] for r in rows]).astype(np.float32))


# ---- training batch ---------------------------------------------------------------------------------------------
def batch_segments(tensors):
    """Segments (for ``pack_numbers``) of a training batch: per tensor ``len(shape), *shape, n_elements, *values``
    (async_experience_buffer.py:76-96).  ``tensors``: numpy arrays in the reference's order (obs, next_obs, reward,
    nonterminal, gamma, action)."""
    segs = []
    for a in tensors:
        a = np.ascontiguousarray(a)
        segs.append(np.asarray([a.ndim, *a.shape, a.size], dtype=np.int64))
        segs.append(a.reshape(-1))
    return segs


def split_batch(flat, n_tensors=6):
    """Inverse of ``batch_segments`` on the decoded numbers: float32 arrays of the sent shapes
    (async_experience_buffer.py:148-184)."""
    out, idx = [], 0
    for _ in range(n_tensors):
        ndim = int(flat[idx])
        shape = tuple(int(s) for s in flat[idx + 1:idx + 1 + ndim])
        n = int(flat[idx + 1 + ndim])
        idx += 2 + ndim
        if n != int(np.prod(shape)) or idx + n > flat.size:
            raise _lib.PbError("malformed batch on the wire")
        out.append(flat[idx:idx + n].astype(np.float32).reshape(shape))
        idx += n
    return out
