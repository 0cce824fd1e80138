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
