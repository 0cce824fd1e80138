"""CPU restatement of the reference's transition store -- TEST INFRASTRUCTURE ONLY.

Follows, on plain Python objects and numpy (this is how the reference itself works: a linked
list of per-step heap objects walked by a Python loop):
  * the link rules of the collector,
    multiprocessing_experience_collection/collector_process_interface.py:146-173
    (done -> no next; truncated -> strong next holding the final observation, never stored;
    otherwise weak prev/next to the neighbours; the newest step of a stream is in flight and
    has reward None);
  * TimestepBuffer._compute_n_step, prism/experience/timestep_buffer.py:198-238;
  * TimestepBuffer._timesteps_to_batch / _stack_obs_into, prism/experience/timestep_buffer.py:79-257;
  * the ring storage + priority sampler of torchrl through oracle.per_oracle.OracleTree
    (PARITY UNPINNED for that part, see per_oracle.c).

Validated against the reference's own TimestepBuffer (imported from /root/reference by
oracle/gen_golden.py; frozen outputs in tests/golden/nstep_gather_*.npz).

One deliberate difference, shared with the product: frame-stack rows the reference does not write
(short chains, timestep_buffer.py:240-257) are zero here, where the reference leaves stale pixels
from earlier batches.
"""
import weakref

import numpy as np

from .per_oracle import OracleTree


class Step:
    """One environment step (fields of prism/experience/timestep.py:12-28 that the path reads)."""
    __slots__ = ("id", "obs", "reward", "done", "truncated", "action", "n_step_return", "n_step_gamma",
                 "n_step_done", "needs_n_step", "n_step_next", "prev", "next", "__weakref__")

    def __init__(self, id):
        self.id = id
        self.obs = None
        self.reward = None
        self.done = None
        self.truncated = None
        self.action = None
        self.n_step_return = None
        self.n_step_gamma = None
        self.n_step_done = None
        self.needs_n_step = True
        self.n_step_next = None
        self.prev = None
        self.next = None


class StreamLinker:
    """Builds linked Steps for one collector stream the way CollectorProcessInterface does."""

    def __init__(self, first_obs, make_step):
        self.make_step = make_step
        self.current = make_step()
        self.current.obs = first_obs

    def step(self, action, reward, done, trunc, next_obs, final_obs=None):
        """Complete the in-flight step with (action, reward, done, trunc); `next_obs` starts the next step
        (the reset observation after done/trunc); `final_obs` is the truncated episode's last observation.
        Returns the completed step (what the collector hands to buffer.extend)."""
        cur = self.current
        cur.action = action
        nxt = self.make_step()
        nxt.obs = next_obs
        cur.reward = float(reward)
        cur.done = bool(done)
        cur.truncated = bool(trunc)
        if trunc:
            tail = self.make_step()
            tail.obs = final_obs
            tail.prev = weakref.ref(cur)
            cur.next = tail                      # strong reference: only the truncated step owns it
        elif not done:
            nxt.prev = weakref.ref(cur)
            cur.next = weakref.ref(nxt)
        self.current = nxt
        return cur


def _deref(link):
    if link is None:
        return None
    return link if isinstance(link, Step) else link()


class OracleTimestepBuffer:
    """Ring of Steps + priority trees + batch assembly, reference semantics."""

    def __init__(self, capacity, batch_size, frame_stack=1, n_step=3, gamma=0.99, alpha=0.5, beta=0.5,
                 eps=1e-8, prioritized=True, **tree_flags):
        self.capacity = capacity
        self.batch_size = batch_size
        self.frame_stack = frame_stack
        self.n_step = n_step
        self.gammas = [gamma ** i for i in range(n_step + 1)]
        self.beta = beta
        self.storage = []
        self.cursor = 0
        self.tree = OracleTree(capacity, alpha=alpha, eps=eps, **tree_flags) if prioritized else None

    def __len__(self):
        return len(self.storage)

    # -- ring write (ListStorage + RoundRobinWriter) ------------------------------------------
    def extend(self, step):
        if self.cursor < len(self.storage):
            self.storage[self.cursor] = step
        else:
            self.storage.append(step)
        idx = self.cursor
        self.cursor = (self.cursor + 1) % self.capacity
        if self.tree is not None:
            self.tree.extend(1)
        return idx

    # -- n-step (timestep_buffer.py:198-238) ---------------------------------------------------
    def compute_n_step(self, start):
        ts, ret, gamma, incomplete = start, 0, 1, False
        for i in range(self.n_step):
            ret += ts.reward * self.gammas[i]
            gamma = self.gammas[i + 1]
            incomplete = i != self.n_step - 1
            link = ts.next
            if link is None or ts.truncated or not incomplete:
                break
            nxt = _deref(link)
            if nxt is None or nxt.reward is None:
                break
            ts = nxt
        start.n_step_return = ret
        start.n_step_gamma = gamma
        start.n_step_done = ts.done
        start.needs_n_step = incomplete and not ts.done and not ts.truncated
        succ = ts.next
        start.n_step_next = weakref.ref(succ) if isinstance(succ, Step) else succ

    # -- batch assembly (timestep_buffer.py:79-196, 240-257) -----------------------------------
    def _stack(self, ts, out, nxt=None, nxt_out=None):
        i = self.frame_stack - 1
        while i >= 0 and ts is not None:
            out[i] = ts.obs
            if nxt is not None:
                nxt_out[i] = nxt.obs
                nxt = _deref(nxt.prev)
            if ts.prev is None:
                break
            ts = ts.prev()
            i -= 1

    def batch_from(self, steps):
        B, fs = len(steps), self.frame_stack
        shape = np.asarray(steps[0].obs).shape
        obs = np.zeros((B, fs) + shape, np.float32)
        nobs = np.zeros((B, fs) + shape, np.float32)
        rew = np.zeros((B, 1), np.float32)
        nonterm = np.zeros((B, 1), bool)
        gam = np.ones((B, 1), np.float32)
        act = np.zeros((B, 1), np.int64)
        for i, ts in enumerate(steps):
            if ts.needs_n_step:
                self.compute_n_step(ts)
            succ = _deref(ts.n_step_next)
            if succ is None:
                if fs == 1:
                    obs[i, 0] = ts.obs
                else:
                    self._stack(ts, obs[i])
                nobs[i] = obs[i]
            elif fs == 1:
                obs[i, 0] = ts.obs
                nobs[i, 0] = succ.obs
            else:
                self._stack(ts, obs[i], succ, nobs[i])
            rew[i] = ts.n_step_return
            nonterm[i] = 1 - ts.n_step_done
            gam[i] = ts.n_step_gamma
            act[i] = ts.action
        return {"observation": obs, "next": {"observation": nobs, "reward": rew}, "nonterminal": nonterm,
                "gamma": gam, "action": act}

    # -- sample / update (PrioritizedSampler semantics) -----------------------------------------
    def sample(self, u=None, batch_size=None, mode=0, rng=None):
        B = self.batch_size if batch_size is None else batch_size
        if u is None:
            u = (rng or np.random).random(B)
        if self.tree is not None:
            idx, w, _, _, _ = self.tree.sample(u, self.beta, mode)
        else:
            idx = np.minimum((np.asarray(u) * len(self.storage)).astype(np.int64), len(self.storage) - 1)
            w = np.ones(B, np.float32)
        batch = self.batch_from([self.storage[i] for i in idx])
        return batch, {"index": idx, "_weight": w}

    def update_priority(self, idx, prio):
        if self.tree is not None:
            self.tree.update_priority(idx, np.abs(prio))
