"""GPU parity of the device transition ring (csrc/store.cu) + the drop-in TimestepBuffer against
(a) the REFERENCE TimestepBuffer outputs frozen in tests/golden/nstep_gather_*.npz and
(b) the CPU oracle buffer on identical traces, priorities and uniforms.
Bar: observations / actions / flags / sampled indices bit-exact; n-step returns within 1e-6 rel."""
import numpy as np
import pytest
import torch

from helpers import load_golden, script_from_fixture, script_successor_obs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _product_buffer(capacity, frame_stack, batch_size=8, sampling="iid", **kw):
    from prism_b200 import DevicePrioritizedReplayBuffer, TimestepBuffer
    rb = DevicePrioritizedReplayBuffer(capacity, alpha=0.5, beta=0.5, batch_size=batch_size, device=DEV,
                                       sampling=sampling, max_streams=8, staging_rows=5, **kw)
    return TimestepBuffer(rb, frame_stack=frame_stack, device=DEV, n_step=3, gamma=0.99)


def _gather_all(tb, idx):
    """Run the fused n-step/gather kernel on chosen slots (bypassing the sampler)."""
    ring = tb.buffer._storage
    tb._flush()
    B = len(idx)
    fs = tb.frame_stack
    obs = torch.full((B, fs) + ring.obs_shape, 7.0, device=DEV)      # poison: every row must be written
    nobs = torch.full((B, fs) + ring.obs_shape, 7.0, device=DEV)
    ret = torch.empty(B, 1, device=DEV); gam = torch.empty(B, 1, device=DEV)
    nt = torch.empty(B, 1, dtype=torch.bool, device=DEV); act = torch.empty(B, 1, dtype=torch.int64, device=DEV)
    ring.gather(torch.as_tensor(idx, dtype=torch.int64, device=DEV), obs, nobs, ret, gam, nt, act)
    torch.cuda.synchronize()
    return [t.cpu().numpy() for t in (obs, nobs, ret, gam, nt, act)]


def _check_against_fixture(fx, cp, got):
    obs, nobs, ret, gam, nt, act = got
    tag = "cp%d." % cp
    assert np.array_equal(obs, fx[tag + "observation"])
    assert np.array_equal(nobs, fx[tag + "next_observation"])
    assert np.array_equal(act, fx[tag + "action"])
    assert np.array_equal(nt, fx[tag + "nonterminal"])
    assert np.allclose(ret, fx[tag + "reward"], rtol=1e-6, atol=0)
    assert np.allclose(gam, fx[tag + "gamma"], rtol=1e-6, atol=0)


@pytest.mark.parametrize("name", ["nstep_gather_fs1", "nstep_gather_fs4"])
def test_object_api_matches_reference_golden(name):
    """extend(Timestep) with collector-style linked objects, exactly how the reference is driven."""
    from oracle.buffer_oracle import StreamLinker
    from prism_b200 import Timestep
    fx = load_golden(name)
    S = script_from_fixture(fx)
    obs_shape = tuple(fx["obs_shape"].tolist())
    tb = _product_buffer(int(fx["capacity"]), int(fx["frame_stack"]))
    ids = [0]

    def make_step():
        ids[0] += 1
        return Timestep(ids[0])

    linkers, cps = {}, set(fx["checkpoints"].tolist())
    for t in range(len(S["stream"])):
        s = int(S["stream"][t])
        if s not in linkers:
            linkers[s] = StreamLinker(torch.from_numpy(S["obs"][t].reshape(obs_shape).copy()), make_step)
        step = linkers[s].step(int(S["action"][t]), float(S["reward"][t]), bool(S["done"][t]), bool(S["trunc"][t]),
                               torch.from_numpy(S["next_obs"][t].reshape(obs_shape).copy()),
                               torch.from_numpy(S["final_obs"][t].reshape(obs_shape).copy()))
        tb.extend(step)
        if (t + 1) in cps:
            _check_against_fixture(fx, t + 1, _gather_all(tb, fx["cp%d.index" % (t + 1)]))


@pytest.mark.parametrize("name", ["nstep_gather_fs1", "nstep_gather_fs4"])
def test_batched_ingest_matches_reference_golden(name):
    """extend_batch with explicit collector stream ids (the ingest path) gives the same buffer."""
    fx = load_golden(name)
    S = script_from_fixture(fx)
    tb = _product_buffer(int(fx["capacity"]), int(fx["frame_stack"]))
    n = len(S["stream"])
    succ = np.stack([script_successor_obs(S, t) for t in range(n)])
    obs_shape = tuple(fx["obs_shape"].tolist())
    done = 0
    for cp in fx["checkpoints"].tolist():
        sl = slice(done, cp)
        tb.extend_batch(S["stream"][sl], S["obs"][sl].reshape((-1,) + obs_shape), S["action"][sl], S["reward"][sl],
                        S["done"][sl], S["trunc"][sl], succ[sl].reshape((-1,) + obs_shape))
        done = cp
        _check_against_fixture(fx, cp, _gather_all(tb, fx["cp%d.index" % cp]))


@pytest.mark.parametrize("sampling", ["iid", "stratified"])
@pytest.mark.parametrize("frame_stack", [1, 3])
def test_learner_loop_matches_oracle(sampling, frame_stack):
    """collect -> sample -> update_priority iterations (prism/learner.py:80-120) on both sides with
    identical traces, uniforms and new priorities: indices, weights and batches must agree."""
    from oracle.buffer_oracle import OracleTimestepBuffer, Step, StreamLinker
    from oracle.gen_golden import make_script
    cap, B, obs_shape = 96, 16, (4, 2)
    S = make_script(21, n_streams=5, n_steps=400, obs_shape=obs_shape, p_done=0.06, p_trunc=0.04)
    tb = _product_buffer(cap, frame_stack, batch_size=B, sampling=sampling)
    ob = OracleTimestepBuffer(cap, B, frame_stack=frame_stack, n_step=3, gamma=0.99)
    ids = [0]

    def make_step():
        ids[0] += 1
        return Step(ids[0])

    linkers, rng = {}, np.random.default_rng(3)
    succ = np.stack([script_successor_obs(S, t) for t in range(400)])
    mode = 1 if sampling == "stratified" else 0
    t = 0
    while t < 400:
        n = int(rng.integers(1, 9))
        sl = slice(t, min(400, t + n))
        tb.extend_batch(S["stream"][sl], S["obs"][sl].reshape((-1,) + obs_shape), S["action"][sl], S["reward"][sl],
                        S["done"][sl], S["trunc"][sl], succ[sl].reshape((-1,) + obs_shape))
        for k in range(sl.start, sl.stop):
            s = int(S["stream"][k])
            if s not in linkers:
                linkers[s] = StreamLinker(S["obs"][k].reshape(obs_shape).copy(), make_step)
            ob.extend(linkers[s].step(int(S["action"][k]), float(S["reward"][k]), bool(S["done"][k]),
                                      bool(S["trunc"][k]), S["next_obs"][k].reshape(obs_shape).copy(),
                                      S["final_obs"][k].reshape(obs_shape).copy()))
        t = sl.stop
        if t < 20:
            continue
        u = rng.random(B)
        tb.inject_uniforms(torch.from_numpy(u).to(DEV))
        batch, info = tb.sample(return_info=True)
        obatch, oinfo = ob.sample(u=u, mode=mode)
        assert np.array_equal(info["index"].cpu().numpy(), oinfo["index"])
        assert np.allclose(info["_weight"].cpu().numpy(), oinfo["_weight"], rtol=1e-6)
        assert np.array_equal(batch["observation"].cpu().numpy(), obatch["observation"])
        assert np.array_equal(batch["next"]["observation"].cpu().numpy(), obatch["next"]["observation"])
        assert np.array_equal(batch["action"].cpu().numpy(), obatch["action"])
        assert np.array_equal(batch["nonterminal"].cpu().numpy(), obatch["nonterminal"])
        assert np.allclose(batch["next"]["reward"].cpu().numpy(), obatch["next"]["reward"], rtol=1e-6, atol=0)
        assert np.allclose(batch["gamma"].cpu().numpy(), obatch["gamma"], rtol=1e-6, atol=0)
        newp = rng.exponential(1.0, B).astype(np.float32)
        tb.update_priority(info["index"], torch.from_numpy(newp).to(DEV))
        ob.update_priority(oinfo["index"], newp)
    torch.cuda.synchronize()
    tree = tb.buffer._sampler
    cap2 = tree.capacity
    assert np.array_equal(tree.sum.cpu().numpy()[cap2:cap2 + cap], ob.tree.sum[ob.tree.capacity:ob.tree.capacity + cap])
    assert tree.state_host()["max_priority"] == np.float32(ob.tree.max_priority)


def test_uint8_storage_widens_exactly():
    """Atari-shaped frames stored as uint8, emitted as k/255 fp32 (gymnasium scale_obs)."""
    from prism_b200 import DevicePrioritizedReplayBuffer, TimestepBuffer
    rng = np.random.default_rng(0)
    n, shape = 40, (84, 84)
    frames = rng.integers(0, 256, (n + 1,) + shape, dtype=np.uint8)
    rb = DevicePrioritizedReplayBuffer(64, batch_size=8, device=DEV, storage_dtype=torch.uint8, obs_scale=True,
                                       max_streams=2, staging_rows=16)
    tb = TimestepBuffer(rb, frame_stack=4, device=DEV, n_step=3, gamma=0.99)
    done = np.zeros(n, bool); done[17] = True
    tb.extend_batch(np.zeros(n, np.int32), frames[:n], rng.integers(0, 18, n), rng.standard_normal(n).astype(np.float32),
                    done, np.zeros(n, bool), frames[1:n + 1])
    obs, nobs, ret, gam, nt, act = _gather_all(tb, np.arange(n))
    ref = frames.astype(np.float32) / np.float32(255.0)
    for i in (0, 3, 10, 17, 18, 25, 39):
        ep_start = 0 if i <= 17 else 18
        for c in range(4):
            j = i - c
            want = ref[j] if j >= ep_start else np.zeros(shape, np.float32)
            assert np.array_equal(obs[i, 3 - c], want), (i, c)
    assert np.array_equal(nobs[17], obs[17])            # terminal: next_obs := obs
    assert np.array_equal(nobs[5, 3], ref[8])           # 3-step successor
    assert np.array_equal(nobs[38, 3], ref[40])         # in-flight tail: successor is the staged next frame


def test_save_load_roundtrip(tmp_path):
    fx = load_golden("nstep_gather_fs4")
    S = script_from_fixture(fx)
    obs_shape = tuple(fx["obs_shape"].tolist())
    n = 100
    succ = np.stack([script_successor_obs(S, t) for t in range(n)])
    tb = _product_buffer(int(fx["capacity"]), 4)
    tb.extend_batch(S["stream"][:n], S["obs"][:n].reshape((-1,) + obs_shape), S["action"][:n], S["reward"][:n],
                    S["done"][:n], S["trunc"][:n], succ.reshape((-1,) + obs_shape))
    before = _gather_all(tb, fx["cp100.index"])
    tb.save(str(tmp_path))
    tb2 = _product_buffer(int(fx["capacity"]), 4)
    tb2.load(str(tmp_path))
    after = _gather_all(tb2, fx["cp100.index"])
    for a, b in zip(before, after):
        assert np.array_equal(a, b)
    assert np.array_equal(tb.buffer._sampler.sum.cpu().numpy(), tb2.buffer._sampler.sum.cpu().numpy())


def test_empty_buffer_raises_like_torchrl():
    tb = _product_buffer(32, 1)
    with pytest.raises(RuntimeError):
        tb.sample()


@pytest.mark.parametrize("name,n", [("nstep_gather_fs1", 2), ("nstep_gather_fs4", 1)])
def test_graph_captured_ingest_matches_reference_golden(name, n):
    """ingest_graph(n): pinned staging -> H2D -> scatter -> default priorities replayed as one CUDA graph
    must build exactly the same buffer as the eager ingest."""
    fx = load_golden(name)
    S = script_from_fixture(fx)
    obs_shape = tuple(fx["obs_shape"].tolist())
    tb = _product_buffer(int(fx["capacity"]), int(fx["frame_stack"]))
    total = len(S["stream"])
    succ = np.stack([script_successor_obs(S, t) for t in range(total)])
    first = n                                                        # the ring must exist before capture
    tb.extend_batch(S["stream"][:first], S["obs"][:first].reshape((-1,) + obs_shape), S["action"][:first],
                    S["reward"][:first], S["done"][:first], S["trunc"][:first], succ[:first].reshape((-1,) + obs_shape))
    push = tb.ingest_graph(n)
    cps = set(fx["checkpoints"].tolist())
    for t in range(first, total - total % n, n):
        sl = slice(t, t + n)
        push(S["stream"][sl], S["obs"][sl], S["action"][sl], S["reward"][sl], S["done"][sl], S["trunc"][sl], succ[sl])
        for cp in range(t + 1, t + n + 1):
            if cp in cps and cp == t + n:
                _check_against_fixture(fx, cp, _gather_all(tb, fx["cp%d.index" % cp]))
    st = tb.buffer._sampler.state_host()
    assert st["seq"] == total - total % n and st["len"] == min(st["seq"], int(fx["capacity"]))


def test_wire_blocks_fill_the_device_buffer():
    """Remote-actor path: the reference's step blocks (tests/golden/wire.npz, bytes frozen from
    Timestep.serialize) -> native decoder -> extend_batch, held by the learner-side interface in local-buffer mode,
    give the same device buffer as handing the same steps to extend_batch with the collector's own stream ids."""
    from helpers import FakeRedis
    from prism_b200.async_components.async_experience_buffer import AsyncExperienceBufferInterface
    from prism_b200.async_components.redis import RedisInterface
    fx = load_golden("wire")
    S = script_from_fixture(fx)
    obs_shape = tuple(fx["obs_shape"].tolist())
    n = len(S["stream"])
    server = FakeRedis()
    for k in range(int(fx["n_blocks"])):                               # what a reference collector would have pushed
        server.lpush(RedisInterface.TIMESTEPS_KEY, fx["block%d.packed" % k].tobytes())
    wired = _product_buffer(128, 2)
    learner = AsyncExperienceBufferInterface("h", 0, DEV, redis_interface=RedisInterface(client=server), local_buffer=wired)
    n_in = learner.drain()
    assert n_in == n - learner._decoder.n_waiting and len(wired) == n_in and learner.drain() == 0
    # the order the steps were released in, recovered from the stored observations
    got = _gather_all(wired, np.arange(n_in))
    key = {S["obs"][t].tobytes(): t for t in range(n)}
    t_of = np.asarray([key[o[-1].reshape(-1).tobytes()] for o in got[0]])
    assert len(set(t_of.tolist())) == n_in
    direct = _product_buffer(128, 2)
    succ = np.stack([script_successor_obs(S, t) for t in range(n)])
    direct.extend_batch(S["stream"][t_of], S["obs"][t_of].reshape((-1,) + obs_shape), S["action"][t_of],
                        S["reward"][t_of], S["done"][t_of], S["trunc"][t_of], succ[t_of].reshape((-1,) + obs_shape))
    want = _gather_all(direct, np.arange(n_in))
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
