"""oracle/buffer_oracle.py (linked-object restatement) and the array n-step restatement in
oracle/per_oracle.c must reproduce the REFERENCE TimestepBuffer outputs frozen in
tests/golden/nstep_gather_*.npz; the host link planner of the product library (pure host code,
runs without a GPU) must produce links that lead to the same answers."""

import numpy as np
import pytest

from helpers import load_golden, replay_script_oracle, script_from_fixture, script_successor_obs


@pytest.mark.parametrize("name", ["nstep_gather_fs1", "nstep_gather_fs4"])
def test_oracle_buffer_matches_reference(name):
    fx = load_golden(name)
    buf, linkers, done = None, None, 0
    for cp in fx["checkpoints"].tolist():
        buf, linkers = replay_script_oracle(fx, upto=cp, buf=buf, linkers=linkers, start=done)
        done = cp
        idx = fx["cp%d.index" % cp]
        batch = buf.batch_from([buf.storage[i] for i in idx])
        tag = "cp%d." % cp
        assert np.array_equal(batch["observation"], fx[tag + "observation"])
        assert np.array_equal(batch["next"]["observation"], fx[tag + "next_observation"])
        assert np.array_equal(batch["action"], fx[tag + "action"])
        assert np.array_equal(batch["nonterminal"], fx[tag + "nonterminal"])
        # n-step returns within 1e-6 relative (north star); identical fp64 accumulation -> exact here
        assert np.array_equal(batch["next"]["reward"], fx[tag + "reward"])
        assert np.array_equal(batch["gamma"], fx[tag + "gamma"])


def plan_script(fx, upto, staging=7):
    """Run the product's HOST planner (pb_store_extend_plan) over the trace and apply its output to
    numpy mirrors of the ring arrays -- what pb_store_scatter does on the device."""
    from prism_b200 import _lib
    lib = _lib.load()
    S = script_from_fixture(fx)
    size = int(fx["capacity"])
    n_streams, pool = 4, 16
    aux_size = n_streams + pool
    E = S["obs"].shape[1]
    ring = {"obs": np.zeros((size, E), np.float32), "aux": np.zeros((aux_size, E), np.float32),
            "action": np.zeros(size, np.int32), "reward": np.zeros(size, np.float32),
            "done": np.zeros(size, np.uint8), "trunc": np.zeros(size, np.uint8),
            "slot_seq": np.full(size, -1, np.int64), "next_link": np.full(size, -1, np.int64),
            "prev_link": np.full(size, -1, np.int64)}
    stream_last = np.full(n_streams, -1, np.int64)
    cursor = np.zeros(1, np.int64)
    owner = np.full(pool, -1, np.int64)
    seq0 = 0
    P = lambda a: a.ctypes.data
    while seq0 < upto:
        n = min(staging, upto - seq0)
        sl = slice(seq0, seq0 + n)
        sid = np.ascontiguousarray(S["stream"][sl], np.int32)
        flags = (S["done"][sl].astype(np.uint8) * 1 + S["trunc"][sl].astype(np.uint8) * 2).astype(np.uint8)
        meta = np.zeros(n, dtype=np.dtype(_lib.STEP_META_DTYPE))
        rc = lib.pb_store_extend_plan(size, aux_size, n_streams, n, seq0, P(sid), P(flags), P(stream_last), P(cursor),
                                      P(owner), P(meta))
        assert rc == 0
        seq, prev, nxt, aux, ps, pv = (meta[k] for k in ("seq", "prev_link", "next_link", "aux_row", "patch_slot",
                                                         "patch_val"))
        for j in range(n):
            t = seq0 + j
            slot = seq[j] % size
            ring["obs"][slot] = S["obs"][t]
            ring["action"][slot], ring["reward"][slot] = S["action"][t], S["reward"][t]
            ring["done"][slot], ring["trunc"][slot] = S["done"][t], S["trunc"][t]
            ring["slot_seq"][slot], ring["prev_link"][slot], ring["next_link"][slot] = seq[j], prev[j], nxt[j]
            if ps[j] >= 0:
                ring["next_link"][ps[j]] = pv[j]
            if aux[j] >= 0:
                ring["aux"][aux[j]] = script_successor_obs(S, t)
        seq0 += n
    return ring, size


def assemble_from_arrays(ring, size, idx, fs, n_step, gamma, obs_shape):
    """numpy model of pb_store_gather on top of the C n-step restatement."""
    from oracle.per_oracle import nstep_arrays
    ret, gam, done, last, succ = nstep_arrays(size, n_step, gamma, ring["slot_seq"], ring["next_link"],
                                              ring["reward"], ring["done"], ring["trunc"], idx)
    B, E = len(idx), ring["obs"].shape[1]
    obs = np.zeros((B, fs, E), np.float32)
    nobs = np.zeros((B, fs, E), np.float32)

    def alive(link):
        return link >= 0 and ring["slot_seq"][link % size] == link

    def walk(slot, hops):
        for _ in range(hops):
            pl = ring["prev_link"][slot]
            if not alive(pl):
                return None
            slot = pl % size
        return slot

    for b, start in enumerate(idx):
        for c in range(fs):
            cur = walk(start, c)
            if cur is None:
                continue
            obs[b, fs - 1 - c] = ring["obs"][cur]
            if succ[b] == -1:
                nobs[b, fs - 1 - c] = ring["obs"][cur]
            elif c == 0:
                nobs[b, fs - 1] = ring["obs"][succ[b] % size] if succ[b] >= 0 else ring["aux"][-succ[b] - 2]
            else:
                c2 = walk(last[b], c - 1)
                if c2 is not None:
                    nobs[b, fs - 1 - c] = ring["obs"][c2]
    shape = (B, fs) + tuple(obs_shape)
    return obs.reshape(shape), nobs.reshape(shape), ret, gam, done, ring["action"][idx]


@pytest.mark.parametrize("name", ["nstep_gather_fs1", "nstep_gather_fs4"])
def test_host_planner_and_array_model_match_reference(name):
    fx = load_golden(name)
    fs, obs_shape = int(fx["frame_stack"]), tuple(fx["obs_shape"].tolist())
    for cp in fx["checkpoints"].tolist():
        ring, size = plan_script(fx, cp)
        idx = fx["cp%d.index" % cp]
        obs, nobs, ret, gam, done, act = assemble_from_arrays(ring, size, idx, fs, int(fx["n_step"]),
                                                              float(fx["gamma"]), obs_shape)
        tag = "cp%d." % cp
        assert np.array_equal(obs, fx[tag + "observation"])
        assert np.array_equal(nobs, fx[tag + "next_observation"])
        assert np.array_equal(ret, fx[tag + "reward"].reshape(-1))
        assert np.array_equal(gam, fx[tag + "gamma"].reshape(-1))
        assert np.array_equal(done == 0, fx[tag + "nonterminal"].reshape(-1))
        assert np.array_equal(act, fx[tag + "action"].reshape(-1))


def test_planner_pool_exhaustion_is_side_effect_free():
    from prism_b200 import _lib
    lib = _lib.load()
    size, n_streams, pool = 64, 2, 2
    stream_last = np.full(n_streams, -1, np.int64)
    cursor = np.zeros(1, np.int64)
    owner = np.full(pool, -1, np.int64)
    n = 3
    sid = np.zeros(n, np.int32)
    flags = np.full(n, 2, np.uint8)            # three truncations, two pool rows
    meta = np.zeros(n, dtype=np.dtype(_lib.STEP_META_DTYPE))
    P = lambda a: a.ctypes.data
    rc = lib.pb_store_extend_plan(size, n_streams + pool, n_streams, n, 0, P(sid), P(flags), P(stream_last), P(cursor),
                                  P(owner), P(meta))
    assert rc == _lib.PB_E_POOL
    assert (stream_last == -1).all() and cursor[0] == 0 and (owner == -1).all()


@pytest.mark.parametrize("dtype,action_dtype", [(np.float32, np.int64), (np.uint8, np.int32)])
def test_stage_block_equals_the_numpy_staging(dtype, action_dtype):
    """pb_store_stage_block (one host call) == what IngestSlot.fill does with numpy assignments + the planner:
    staged rows, the whole 64-byte metadata records and every piece of planner state, block after block."""
    from prism_b200 import _lib
    lib = _lib.load()
    fx = load_golden("nstep_gather_fs4")
    S = script_from_fixture(fx)
    size, n_streams, pool, n = int(fx["capacity"]), 4, 16, 5
    E = S["obs"].shape[1]
    P = lambda a: a.ctypes.data
    cast = (lambda a: np.ascontiguousarray(a, dtype)) if dtype is np.float32 else \
        (lambda a: np.ascontiguousarray((np.abs(a) * 40).astype(np.uint8)))

    def fresh():
        return {"last": np.full(n_streams, -1, np.int64), "cursor": np.zeros(1, np.int64), "owner": np.full(pool, -1, np.int64),
                "rows": np.zeros((2, n, E), dtype), "meta": np.zeros(n, dtype=np.dtype(_lib.STEP_META_DTYPE))}

    a, b = fresh(), fresh()
    for seq0 in range(0, (len(S["stream"]) // n) * n, n):
        sl = slice(seq0, seq0 + n)
        sid = np.ascontiguousarray(S["stream"][sl], np.int32)
        obs = cast(S["obs"][sl])
        nxt = cast(np.stack([script_successor_obs(S, t) for t in range(seq0, seq0 + n)]))
        action = np.ascontiguousarray(S["action"][sl], action_dtype)
        reward = np.ascontiguousarray(S["reward"][sl], np.float32)
        done, trunc = np.ascontiguousarray(S["done"][sl]), np.ascontiguousarray(S["trunc"][sl])
        assert done.dtype == np.bool_
        # numpy staging, as in IngestSlot.fill
        a["rows"][0], a["rows"][1] = obs, nxt
        d8, t8 = done.astype(np.uint8), trunc.astype(np.uint8)
        flags = (d8 * 1 + t8 * 2).astype(np.uint8)
        a["meta"]["action"], a["meta"]["reward"], a["meta"]["done"], a["meta"]["trunc"] = action, reward, d8, t8
        assert lib.pb_store_extend_plan(size, n_streams + pool, n_streams, n, seq0, P(sid), P(flags), P(a["last"]),
                                        P(a["cursor"]), P(a["owner"]), P(a["meta"])) == 0
        # one call
        assert lib.pb_store_stage_block(size, n_streams + pool, n_streams, n, seq0, obs[0].nbytes, P(obs), P(nxt), P(sid),
                                        P(action), action.itemsize, P(reward), P(done), P(trunc), P(b["rows"]), P(b["last"]),
                                        P(b["cursor"]), P(b["owner"]), P(b["meta"])) == 0
        assert a["rows"].tobytes() == b["rows"].tobytes() and a["meta"].tobytes() == b["meta"].tobytes()
        for k in ("last", "cursor", "owner"):
            assert np.array_equal(a[k], b[k])
    assert S["trunc"].any() and S["done"].any() and (a["owner"] >= 0).any()


def test_stage_block_failure_leaves_the_block_untouched():
    from prism_b200 import _lib
    lib = _lib.load()
    size, n_streams, pool, n, E = 64, 2, 2, 3, 4
    P = lambda a: a.ctypes.data
    last, cursor, owner = np.full(n_streams, -1, np.int64), np.zeros(1, np.int64), np.full(pool, -1, np.int64)
    rows = np.full((2, n, E), 7.0, np.float32)
    meta = np.zeros(n, dtype=np.dtype(_lib.STEP_META_DTYPE))
    obs = np.ones((n, E), np.float32)
    keep = [np.zeros(n, np.int32), np.zeros(n, np.int64), np.zeros(n, np.float32), np.zeros(n, np.uint8),
            np.ones(n, np.uint8)]                                        # three truncations, two pool rows
    args = (P(obs), P(obs), P(keep[0]), P(keep[1]), 8) + tuple(P(a) for a in keep[2:])
    rc = lib.pb_store_stage_block(size, n_streams + pool, n_streams, n, 0, 4 * E, *args, P(rows), P(last), P(cursor),
                                  P(owner), P(meta))
    assert rc == _lib.PB_E_POOL
    assert (rows == 7.0).all() and (last == -1).all() and cursor[0] == 0 and (owner == -1).all()
    assert lib.pb_store_stage_block(size, n_streams + pool, n_streams, n, 0, 4 * E, None, *args[1:], P(rows), P(last),
                                    P(cursor), P(owner), P(meta)) == -1      # PB_E_ARG


def test_planner_property_random_traces_match_the_linked_object_oracle():
    """Random 4-stream traces, ring sizes (wrapping several times), frame stacks and staging block sizes: the host
    planner + the array model of the gather kernel give the same batches as the linked-object oracle buffer."""
    from hypothesis import given, settings, strategies as st
    from oracle.gen_golden import make_script

    @settings(max_examples=80, deadline=None, derandomize=True)
    @given(st.integers(0, 2 ** 31 - 1), st.integers(8, 64), st.integers(1, 150), st.integers(1, 4), st.integers(1, 11),
           st.floats(0.0, 0.3), st.floats(0.0, 0.08))
    def check(seed, capacity, n_steps, fs, staging, p_done, p_trunc):
        obs_shape = (2, 2)
        S = make_script(seed, n_streams=4, n_steps=n_steps, obs_shape=obs_shape, p_done=p_done, p_trunc=p_trunc)
        fx = {"capacity": np.array(capacity), "frame_stack": np.array(fs), "n_step": np.array(3),
              "gamma": np.array(0.99), "obs_shape": np.array(obs_shape)}
        fx.update({"script." + k: v for k, v in S.items()})
        ring, size = plan_script(fx, n_steps, staging=min(staging, capacity))
        buf, linkers = replay_script_oracle(fx)
        idx = np.arange(min(n_steps, capacity))
        want = buf.batch_from([buf.storage[i] for i in idx])
        obs, nobs, ret, gam, done, act = assemble_from_arrays(ring, size, idx, fs, 3, 0.99, obs_shape)
        assert np.array_equal(obs, want["observation"]) and np.array_equal(nobs, want["next"]["observation"])
        assert np.array_equal(ret, want["next"]["reward"].reshape(-1)) and np.array_equal(gam, want["gamma"].reshape(-1))
        assert np.array_equal(done == 0, want["nonterminal"].reshape(-1)) and np.array_equal(act, want["action"].reshape(-1))

    check()
