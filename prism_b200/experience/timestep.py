"""`Timestep`: one environment step as the collector hands it to the buffer.

Field-for-field mirror of the reference dataclass (prism/experience/timestep.py:12-28) so
that collectors written against the reference keep working; the device buffer only reads
``obs, reward, done, truncated, action, prev, next`` (duck-typed, so the reference's own
Timestep objects are accepted too).  The n-step cache fields are kept for API
compatibility but are never used: the device path recomputes n-step returns at gather
time (csrc/store.cu), which is equivalent.
"""
from dataclasses import dataclass
from typing import Any


@dataclass()
class Timestep(object):
    id: int
    obs: Any = None
    reward: float = None
    done: bool = None
    truncated: bool = None
    action: int = None
    n_step_return: float = None
    n_step_gamma: float = None
    n_step_done: bool = None
    needs_n_step: bool = True
    episodic_reward: float = 0

    n_step_next: Any = None
    prev: Any = None
    next: Any = None

    # ---- wire format of the Redis transport (prism/experience/timestep.py:30-262) ---------------------------------
    def serialize(self):
        from ..async_components.wire import serialize_timestep
        return serialize_timestep(self)

    @classmethod
    def deserialize_linked_list(cls, serialized_timesteps, timestep_id_map=None):
        """Rebuild linked ``Timestep`` objects from a flat block (object path; the batched device ingest uses
        ``async_components.wire.TimestepWireDecoder`` instead).  Same contract as the reference: returns the steps
        whose prev / n_step_next / next links all resolved, and ``{id: (timestep, [n_step_next id, prev id, next
        id])}`` of those still waiting, to be passed back in with the next block."""
        import weakref

        import numpy as np
        import torch

        from ..async_components import wire

        known = dict(timestep_id_map) if timestep_id_map else {}
        flat, rec = wire.index_timesteps(np.asarray(serialized_timesteps, dtype=np.float64))
        for c in rec:
            t = flat[c[10]:c[10] + 12]

            def opt(k, cast):
                return None if t[k] == wire.NULL_VALUE else cast(t[k])

            def obs_at(o):
                return torch.from_numpy(wire._obs_view(flat, o[0], o[1], o[2], o[3]).astype(np.float32))

            ts = cls(id=int(c[0]), obs=None if c[1] < 0 else obs_at(c[1:5]), reward=opt(0, float), done=opt(1, bool),
                     truncated=opt(2, bool), action=opt(3, int), n_step_return=opt(4, float), n_step_gamma=opt(5, float),
                     n_step_done=opt(6, bool), needs_n_step=opt(7, bool), episodic_reward=opt(8, float))
            links = [opt(9, int), opt(10, int), opt(11, int)]
            if c[6] >= 0:                                  # truncated: the successor travels inside the record
                ts.next = cls(id=int(c[5]), obs=obs_at(c[6:10]))
                ts.next.prev = weakref.ref(ts)
                links[2] = None
            known[ts.id] = (ts, links)
        complete, waiting = [], {}
        for ts_id, (ts, links) in known.items():
            links = list(links)
            for k, attr in enumerate(("n_step_next", "prev", "next")):
                if links[k] is None:
                    continue
                target = known.get(links[k])
                if target is not None:
                    setattr(ts, attr, weakref.ref(target[0]))
                    links[k] = None
            if all(l is None for l in links):
                complete.append(ts)
            else:
                waiting[ts_id] = (ts, links)
        return complete, waiting
