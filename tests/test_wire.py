"""Remote-actor wire formats against the reference's own serializers (tests/golden/wire.npz, frozen by
oracle/gen_golden.py from prism/experience/timestep.py and prism/async_components/*) -- CPU only: the codec is host
code of libprism_b200.so; the device hand-off is ``TimestepBuffer.extend_batch`` (tests/test_gpu_buffer.py)."""
import os

import msgpack
import numpy as np
import pytest
import torch

from helpers import FakeRedis
from oracle.buffer_oracle import StreamLinker
from prism_b200 import _lib
from prism_b200.async_components import compression_methods as cm
from prism_b200.async_components import wire
from prism_b200.async_components.async_experience_buffer import AsyncExperienceBuffer, AsyncExperienceBufferInterface
from prism_b200.async_components.redis import RedisInterface
from prism_b200.config import Config
from prism_b200.experience.timestep import Timestep

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "wire.npz"))
NONE = cm.MessageSerializer(compression_type="NONE")


# ---------------------------------------------------------------------------------------------- native codec
def test_pack_numbers_is_byte_identical_to_msgpack():
    ints = [0, 1, 127, 128, 255, 256, 65535, 65536, 2 ** 32 - 1, 2 ** 32, 2 ** 53, 2 ** 62, -1, -32, -33, -128, -129,
            -32768, -32769, -2 ** 31, -2 ** 31 - 1, -2 ** 53, -2 ** 63, -1313]
    assert cm.pack_numbers([np.asarray(ints, dtype=np.int64)]) == msgpack.packb(ints)
    rng = np.random.default_rng(0)
    f32 = rng.standard_normal(3000).astype(np.float32)
    f64 = rng.standard_normal(17)
    flags = rng.random(40) < 0.5
    mixed = [3, 2, 5, 10] + f32.tolist() + flags.tolist() + f64.tolist() + [7]
    packed = cm.pack_numbers([np.asarray([3, 2, 5, 10]), f32, flags, f64, 7])
    assert packed == msgpack.packb(mixed)
    assert np.array_equal(cm.unpack_numbers(packed), np.asarray([float(v) for v in mixed]))


@pytest.mark.parametrize("n", [0, 1, 15, 16, 65535, 65536])
def test_array_header_boundaries(n):
    values = [float(i) for i in range(n)]
    packed = cm.pack_numbers([np.asarray(values, dtype=np.float64)])
    assert packed == msgpack.packb(values)
    assert np.array_equal(cm.unpack_numbers(packed), np.asarray(values))


def test_unpack_numbers_reads_every_numeric_encoding():
    values = [5, -3, 200, -200, 70000, -70000, 2 ** 40, -2 ** 40, 1.5, True, False]
    assert cm.unpack_numbers(msgpack.packb(values)).tolist() == [float(v) for v in values]
    single = msgpack.packb([1.25, -2.5], use_single_float=True)
    assert cm.unpack_numbers(single).tolist() == [1.25, -2.5]


@pytest.mark.parametrize("payload", [msgpack.packb([1, None]), msgpack.packb([1, "a"]), msgpack.packb([[1]]),
                                     msgpack.packb({"a": 1}), msgpack.packb([1, 2]) + b"\x00", msgpack.packb([1.0, 2.0])[:-1],
                                     b"\xdd\xff\xff\xff\xff", b""])
def test_unpack_numbers_rejects_anything_else(payload):
    with pytest.raises(_lib.PbError):
        cm.unpack_numbers(payload)


def test_index_rejects_malformed_blocks():
    good = np.asarray(wire.serialize_timestep(Timestep(4, obs=np.zeros((2, 2), np.float32), reward=1.0, done=True,
                                                        truncated=False, action=1)), dtype=np.float64)
    flat, rec = wire.index_timesteps(good)
    assert rec.shape == (1, 12) and rec[0, 0] == 4 and rec[0, 2] == 4 and rec[0, 4] == 2 and rec[0, 11] == good.size
    for bad in (good[:-1], good[:3], np.concatenate([good, good[:5]])):
        with pytest.raises(_lib.PbError):
            wire.index_timesteps(bad)
    lying = good.copy()
    lying[1] = 1e9                                   # observation longer than the block
    with pytest.raises(_lib.PbError):
        wire.index_timesteps(lying)
    assert wire.index_timesteps(np.zeros(0))[1].shape == (0, 12)


# ---------------------------------------------------------------------------------------------- envelope
def test_envelope_tags_and_errors(monkeypatch):
    ser = cm.MessageSerializer()                                          # asks for LZ4 like the reference default
    small = ser.pack([1, 2, 3])
    assert msgpack.unpackb(small)[0] == "NONE" and ser.unpack(small) == [1, 2, 3]
    zeros = [0.0] * 2000
    big = ser.pack(zeros)
    tag, body = msgpack.unpackb(big)
    assert cm.LZ4MessageCompressor.available() and tag == "LZ4"          # lz4, or pyarrow's liblz4 (present here)
    assert body[:4] == b"\x04\x22\x4d\x18" and len(body) < len(msgpack.packb(zeros)) // 10
    assert ser.unpack(big) == zeros
    assert ser.pack(None) is None and ser.unpack(None) is None
    with pytest.raises(ValueError):
        ser.unpack(msgpack.packb(("ZSTD", b"x")))
    # no LZ4 codec at all: the sender degrades to the NONE tag, the receiver refuses loudly
    monkeypatch.setattr(cm, "_LZ4_PROVIDER", [None])
    assert msgpack.unpackb(ser.pack(zeros))[0] == "NONE"
    with pytest.raises(_lib.PbError):
        ser.unpack(big)


def test_lz4_frame_known_answer():
    """A frame written by hand from the LZ4 frame specification (stored block, no checksums): magic, FLG 0x60
    (version 01, independent blocks), BD 0x40 (64 KB), header checksum 0x82, block size with the 'uncompressed' bit,
    end mark -- and the reverse direction: what we emit starts with the same descriptor."""
    payload = msgpack.packb([1.5, 2, True])
    stored = b"\x04\x22\x4d\x18\x60\x40\x82" + (len(payload) | 0x80000000).to_bytes(4, "little") + payload + b"\x00" * 4
    assert cm.LZ4MessageCompressor.decompress(stored) == payload
    ser = cm.MessageSerializer()
    assert ser.unpack(msgpack.packb(("LZ4", stored))) == [1.5, 2, True]
    assert np.array_equal(ser.unpack_numbers(msgpack.packb(("LZ4", stored))), [1.5, 2.0, 1.0])
    ours = cm.LZ4MessageCompressor.compress(b"\x00" * 4096)
    assert ours[:4] == b"\x04\x22\x4d\x18" and (ours[4] >> 6) == 1 and len(ours) < 100
    assert cm.LZ4MessageCompressor.decompress(ours) == b"\x00" * 4096


# ---------------------------------------------------------------------------------------------- golden: sender side
def _replay_script():
    """Rebuild the collector trace of the fixture with THIS package's Timestep; yields the blocks as the sender
    would submit them."""
    S = {k[len("script."):]: GOLD[k] for k in GOLD.files if k.startswith("script.")}
    shape, block = tuple(GOLD["obs_shape"]), int(GOLD["block"])
    ids = [0]

    def make_step():
        ids[0] += 1
        return Timestep(ids[0])

    linkers, keep, pending, blocks = {}, [], [], []
    n = len(S["stream"])
    for t in range(n):
        s = int(S["stream"][t])
        if s not in linkers:
            linkers[s] = StreamLinker(torch.from_numpy(S["obs"][t].reshape(shape).copy()), make_step)
        step = linkers[s].step(int(S["action"][t]), float(S["reward"][t]), bool(S["done"][t]), bool(S["trunc"][t]),
                               torch.from_numpy(S["next_obs"][t].reshape(shape).copy()),
                               torch.from_numpy(S["final_obs"][t].reshape(shape).copy()))
        keep.append(step)
        pending.append(step)
        if len(pending) == block or t == n - 1:
            blocks.append(pending)
            pending = []
    return S, blocks, keep + [lk.current for lk in linkers.values()]      # in-flight steps stay alive too


def test_serialize_matches_reference_bytes():
    S, blocks, keep = _replay_script()
    assert len(blocks) == int(GOLD["n_blocks"])
    assert [ts.id for b in blocks for ts in b] == GOLD["step_ids"].tolist()
    for k, steps in enumerate(blocks):
        serialized = []
        for ts in steps:
            serialized += ts.serialize()
        assert NONE.pack(serialized) == GOLD["block%d.packed" % k].tobytes(), "block %d" % k
        # the array-segment sender (what RedisInterface.submit_timesteps uses) emits the same bytes
        assert NONE.pack_numbers(wire.timestep_segments(steps)) == GOLD["block%d.packed" % k].tobytes()


def test_linked_list_matches_reference_release_order():
    waiting = {}
    for k in range(int(GOLD["n_blocks"])):
        flat = NONE.unpack_numbers(GOLD["block%d.packed" % k].tobytes())
        complete, waiting = Timestep.deserialize_linked_list(flat, waiting)
        assert [ts.id for ts in complete] == GOLD["block%d.released" % k].tolist()
        assert sorted(waiting) == GOLD["block%d.waiting" % k].tolist()
        for ts in complete:                                   # released = every link it names is resolved
            if not ts.done and not ts.truncated:
                assert ts.next() is not None and ts.next().obs is not None


# ---------------------------------------------------------------------------------------------- golden: columnar decoder
def _check_decoder_output(S, parts, dec):
    """The decoder's rows against the collector trace ``S`` they were serialized from."""
    n = len(S["stream"])
    if not parts:
        assert dec.n_waiting == n == dec.n_decoded
        return
    sid, obs, action, reward, done, trunc, next_obs = (np.concatenate([p[i] for p in parts]) for i in range(7))
    assert obs.dtype == np.float32 and sid.dtype == np.int32 and action.dtype == np.int64
    assert len(sid) == n - dec.n_waiting and dec.n_decoded == n
    key = {S["obs"][t].tobytes(): t for t in range(n)}
    assert len(key) == n
    t_of = np.asarray([key[o.reshape(-1).tobytes()] for o in obs])
    assert len(set(t_of.tolist())) == len(t_of)
    assert np.array_equal(action, S["action"][t_of]) and np.array_equal(reward, S["reward"][t_of])
    assert np.array_equal(done, S["done"][t_of]) and np.array_equal(trunc, S["trunc"][t_of])
    want_next = np.where(S["trunc"][t_of, None], S["final_obs"][t_of], S["next_obs"][t_of])
    want_next[S["done"][t_of]] = 0.0
    assert np.array_equal(next_obs.reshape(len(t_of), -1), want_next)
    # every collector stream comes out in its own order, and one episode keeps one stream id
    live = {}
    for j, t in enumerate(t_of):
        s = int(S["stream"][t])
        if s in live:
            last_t, last_sid = live[s]
            assert t > last_t
            if last_sid is not None:
                assert sid[j] == last_sid
        live[s] = (t, None if (S["done"][t] or S["trunc"][t]) else sid[j])
    open_sids = [v[1] for v in live.values() if v[1] is not None]
    assert len(open_sids) == len(set(open_sids))
    # only the newest unfinished step of a stream may still be waiting for its successor
    missing = sorted(set(range(n)) - set(t_of.tolist()))
    assert len(missing) == dec.n_waiting
    for t in missing:
        assert not S["done"][t] and not S["trunc"][t]
        assert not np.any(S["stream"][t + 1:] == S["stream"][t])


def test_decoder_emits_the_script_in_stream_order():
    S = {k[len("script."):]: GOLD[k] for k in GOLD.files if k.startswith("script.")}
    dec = wire.TimestepWireDecoder(max_streams=8)
    parts = []
    for k in range(int(GOLD["n_blocks"])):
        rows = dec.feed(NONE.unpack_numbers(GOLD["block%d.packed" % k].tobytes()))
        if rows is not None:
            parts.append(rows)
    _check_decoder_output(S, parts, dec)


def test_decoder_property_random_traces_and_block_boundaries():
    """Random collector traces (1-5 streams, ends and truncations at random rates) cut into blocks at random places:
    same invariants as on the reference's fixture; the stream budget is exactly the number of collectors."""
    from hypothesis import given, settings, strategies as st
    from oracle.gen_golden import make_script

    @settings(max_examples=40, deadline=None, derandomize=True)
    @given(st.integers(0, 2 ** 31 - 1), st.integers(1, 5), st.integers(1, 60), st.floats(0.0, 0.5), st.floats(0.0, 0.5),
           st.integers(1, 9))
    def check(seed, n_streams, n_steps, p_done, p_trunc, max_block):
        S = make_script(seed, n_streams=n_streams, n_steps=n_steps, obs_shape=(2, 2), p_done=p_done, p_trunc=p_trunc)
        rng = np.random.default_rng(seed)
        ids = [0]

        def make_step():
            ids[0] += 1
            return Timestep(ids[0])

        linkers, steps = {}, []
        for t in range(n_steps):
            s = int(S["stream"][t])
            if s not in linkers:
                linkers[s] = StreamLinker(S["obs"][t].reshape(2, 2).copy(), make_step)
            steps.append(linkers[s].step(int(S["action"][t]), float(S["reward"][t]), bool(S["done"][t]), bool(S["trunc"][t]),
                                         S["next_obs"][t].reshape(2, 2).copy(), S["final_obs"][t].reshape(2, 2).copy()))
        dec = wire.TimestepWireDecoder(max_streams=n_streams)
        parts, at = [], 0
        while at < n_steps:
            size = int(rng.integers(1, max_block + 1))
            block = steps[at:at + size]
            at += size
            rows = dec.feed(cm.unpack_numbers(cm.pack_numbers(wire.timestep_segments(block))))
            if rows is not None:
                parts.append(rows)
        _check_decoder_output(S, parts, dec)

    check()


def test_decoded_streams_rebuild_the_collectors_links():
    """The CPU twin of tests/test_gpu_buffer.py::test_wire_blocks_fill_the_device_buffer: feeding the decoder's rows
    (with the stream ids IT resolved from the prev / next ids) into the oracle buffer yields the same n-step / frame-
    stacked batches as feeding the same steps with the collector's own stream ids."""
    from oracle.buffer_oracle import OracleTimestepBuffer, Step
    S = {k[len("script."):]: GOLD[k] for k in GOLD.files if k.startswith("script.")}
    shape = tuple(int(d) for d in GOLD["obs_shape"])
    dec = wire.TimestepWireDecoder(max_streams=8)
    parts = [dec.feed(NONE.unpack_numbers(GOLD["block%d.packed" % k].tobytes())) for k in range(int(GOLD["n_blocks"]))]
    sid, obs, action, reward, done, trunc, next_obs = (np.concatenate([p[i] for p in parts if p is not None]) for i in range(7))
    key = {S["obs"][t].tobytes(): t for t in range(len(S["stream"]))}
    t_of = [key[o.reshape(-1).tobytes()] for o in obs]

    def fill(streams, rows):
        buf = OracleTimestepBuffer(256, batch_size=4, frame_stack=2, n_step=3, gamma=0.99, prioritized=False)
        ids, linkers, steps = [0], {}, []

        def make_step():
            ids[0] += 1
            return Step(ids[0])

        for s, (o, a, r, d, tr, nxt) in zip(streams, rows):
            if s not in linkers:
                linkers[s] = StreamLinker(o, make_step)
            step = linkers[s].step(int(a), float(r), bool(d), bool(tr), nxt, nxt)
            if d or tr:
                del linkers[s]                       # the next step on this stream id opens a new episode
            buf.extend(step)
            steps.append(step)
        return buf.batch_from(steps), list(linkers.values())

    wired, keep_a = fill(sid.tolist(), zip(obs, action, reward, done, trunc, next_obs))
    succ = [S["final_obs"][t] if S["trunc"][t] else S["next_obs"][t] for t in t_of]
    direct, keep_b = fill([int(S["stream"][t]) for t in t_of],
                          [(S["obs"][t].reshape(shape), S["action"][t], S["reward"][t], S["done"][t], S["trunc"][t],
                            succ[j].reshape(shape)) for j, t in enumerate(t_of)])
    for a, b in ((wired["observation"], direct["observation"]), (wired["next"]["observation"], direct["next"]["observation"]),
                 (wired["next"]["reward"], direct["next"]["reward"]), (wired["nonterminal"], direct["nonterminal"]),
                 (wired["gamma"], direct["gamma"]), (wired["action"], direct["action"])):
        assert np.array_equal(a, b)
    assert wired["gamma"].min() < 0.99 ** 2 and not wired["nonterminal"].all()       # the trace exercises 3-step windows and ends


def test_decoder_stream_budget():
    dec = wire.TimestepWireDecoder(max_streams=1)
    a = Timestep(1, obs=np.zeros(3, np.float32), reward=0.0, done=False, truncated=False, action=0)
    a.next = Timestep(2, obs=np.ones(3, np.float32))
    b = Timestep(10, obs=np.zeros(3, np.float32), reward=0.0, done=False, truncated=False, action=0)
    b.next = Timestep(11, obs=np.ones(3, np.float32))
    assert dec.feed(np.asarray(a.serialize(), dtype=np.float64)) is None and dec.n_waiting == 1
    with pytest.raises(_lib.PbError):
        dec.feed(np.asarray(b.serialize(), dtype=np.float64))


# ---------------------------------------------------------------------------------------------- golden: batch format
def test_batch_format_matches_reference():
    tensors = [GOLD["batch.in%d" % k] for k in range(6)]
    assert NONE.pack_numbers(wire.batch_segments(tensors)) == GOLD["batch.packed"].tobytes()
    back = wire.split_batch(NONE.unpack_numbers(GOLD["batch.packed"].tobytes()))
    for k in range(6):
        assert back[k].dtype == np.float32 and np.array_equal(back[k], GOLD["batch.out%d" % k])
    with pytest.raises(_lib.PbError):
        wire.split_batch(NONE.unpack_numbers(GOLD["batch.packed"].tobytes())[:-3])


# ---------------------------------------------------------------------------------------------- both ends over a fake Redis
class RecordingBuffer(object):
    """Stands where the device TimestepBuffer stands: records extend_batch, serves a fixed static batch."""

    def __init__(self, B, shape):
        self.rows, self.B, self.shape, self.n_sampled = [], B, shape, 0

    def extend_batch(self, *rows):
        self.rows.append(rows)
        return len(rows[0])

    def sample(self, batch_size=None, return_info=False):
        g = torch.Generator().manual_seed(self.n_sampled)
        self.n_sampled += 1
        self._obs = torch.randn((self.B, 1) + self.shape, generator=g)
        self._next_obs = torch.randn((self.B, 1) + self.shape, generator=g)
        self._reward = torch.randn(self.B, 1, generator=g)
        self._nonterminal = torch.rand(self.B, 1, generator=g) < 0.8
        self._gamma = torch.full((self.B, 1), 0.99 ** 3)
        self._action = torch.randint(0, 5, (self.B, 1), generator=g)


def test_both_ends_over_fake_redis():
    S, blocks, keep = _replay_script()
    shape = tuple(int(s) for s in GOLD["obs_shape"])
    server = FakeRedis()
    cfg = Config(batch_size=4, device="cpu")
    sender = AsyncExperienceBufferInterface("h", 0, "cpu", redis_interface=RedisInterface(client=server), block_size=7)
    sender._redis_interface.serializer = cm.MessageSerializer(min_size_to_compress=64)     # small blocks: force the LZ4 leg
    sender._redis_interface.set_config(cfg)
    assert RedisInterface(client=server).get_config().batch_size == 4
    steps = [ts for b in blocks for ts in b]
    for ts in steps:
        sender.extend(ts)
    sender.flush()
    assert int(server.get(RedisInterface.TOTAL_TIMESTEPS_COLLECTED_KEY)) == len(steps)
    # the blocks on the wire are the reference's, byte for byte, inside the (LZ4-framed) envelope
    queued = list(reversed(server.lists[RedisInterface.TIMESTEPS_KEY]))
    assert [msgpack.unpackb(m)[0] for m in queued] == ["LZ4"] * len(queued)
    assert [NONE._unwrap(m) for m in queued] == [NONE._unwrap(GOLD["block%d.packed" % k].tobytes())
                                                 for k in range(int(GOLD["n_blocks"]))]

    recorder = RecordingBuffer(4, shape)
    middle = AsyncExperienceBuffer("h", 0, redis_interface=RedisInterface(client=server), experience_buffer=recorder)
    middle._idle_sleep = 0.0
    middle._time_between_collect_calls = 0.0
    middle._time_between_command_pings = 0.0
    middle._redis_interface.set_current_command(RedisInterface.START_COLLECTING_COMMAND)
    middle.run(max_iterations=3)
    n_rows = sum(len(r[0]) for r in recorder.rows)
    assert n_rows == len(steps) - middle._decoder.n_waiting == middle._n_collected
    assert recorder.n_sampled == 3 and len(server.lists[RedisInterface.TRAINING_BATCH_KEY]) == 3
    sent = [recorder._obs.clone(), recorder._next_obs.clone(), recorder._reward.clone(),
            recorder._nonterminal.float(), recorder._gamma.clone(), recorder._action.float()]

    learner = AsyncExperienceBufferInterface("h", 0, "cpu", redis_interface=RedisInterface(client=server))
    batch, info = learner.sample(return_info=True)
    assert info == 1 and batch["action"].dtype == torch.int64
    got = [batch["observation"], batch["next"]["observation"], batch["next"]["reward"], batch["nonterminal"],
           batch["gamma"], batch["action"].float()]
    for a, b in zip(got, sent):                       # lpush + lrange from the head: the newest batch comes first
        assert torch.equal(a.float(), b)
    assert learner.sample() is batch and learner.sample() is batch
    with pytest.raises(_lib.PbError):
        learner.update_priority(None, None)
    middle._redis_interface.set_current_command(RedisInterface.SHUTDOWN_COMMAND)
    middle.run(max_iterations=50)                     # returns on the command, not on the iteration cap
    assert recorder.n_sampled == 4


def test_local_buffer_mode_drains_into_the_learners_buffer():
    S, blocks, keep = _replay_script()
    server = FakeRedis()
    sender = AsyncExperienceBufferInterface("h", 0, "cpu", redis_interface=RedisInterface(client=server), block_size=7)
    steps = [ts for b in blocks for ts in b]
    for ts in steps:
        sender.extend(ts)
    sender.flush()

    class Local(RecordingBuffer):
        buffer = type("TD", (), {"_storage_opts": {"max_streams": 16}})()

        def sample(self, return_info=False):
            return ("batch", {"index": 0}) if return_info else "batch"

        def update_priority(self, idx, p):
            self.updated = (idx, p)

    local = Local(4, (3, 2))
    learner = AsyncExperienceBufferInterface("h", 0, "cpu", redis_interface=RedisInterface(client=server), local_buffer=local)
    assert learner.sample(return_info=True) == ("batch", {"index": 0})
    assert sum(len(r[0]) for r in local.rows) == len(steps) - learner._decoder.n_waiting
    assert learner._decoder.max_streams == 16
    learner.update_priority(1, 2)
    assert local.updated == (1, 2)


def test_config_wire_form_keeps_foreign_fields():
    cfg = Config(batch_size=64, per_alpha=0.7)
    back = Config.deserialize(cfg.serialize().encode())
    assert back == cfg
    import json
    foreign = json.loads(cfg.serialize())
    foreign["env_name"] = "MinAtar/Breakout-v1"              # a field only the reference's 91-field Config has
    back = Config.deserialize(json.dumps(foreign).encode())
    assert back.env_name == "MinAtar/Breakout-v1" and back.batch_size == 64


def test_factory_routes_redis_mode_like_the_reference():
    """exp_buffer_factory.py:11-18.  No Redis client library in this image: constructing either end without an
    injected client must fail loudly at the import, not fall back to anything."""
    import prism_b200
    for side in ("server", "client"):
        cfg = Config(run_through_redis=True, redis_side=side, device="cpu")
        try:
            import redis  # noqa: F401
        except ImportError:
            with pytest.raises(ImportError):
                prism_b200.build_exp_buffer(cfg)
    with pytest.raises(ValueError):
        prism_b200.build_exp_buffer(Config(run_through_redis=True, redis_side="sideways", device="cpu"))


def test_codec_property_random_payloads():
    """Random typed segments: the packer is byte-identical to msgpack on the concatenated Python list and the unpacker
    inverts it (NaN and infinities included)."""
    from hypothesis import given, settings, strategies as st

    seg = st.one_of(
        st.lists(st.integers(min_value=-2 ** 63, max_value=2 ** 63 - 1), max_size=40).map(lambda v: np.asarray(v, dtype=np.int64)),
        st.lists(st.floats(width=32, allow_nan=True, allow_infinity=True), max_size=40).map(lambda v: np.asarray(v, dtype=np.float32)),
        st.lists(st.floats(allow_nan=True, allow_infinity=True), max_size=40).map(lambda v: np.asarray(v, dtype=np.float64)),
        st.lists(st.booleans(), max_size=40).map(lambda v: np.asarray(v, dtype=np.bool_)))

    @settings(max_examples=150, deadline=None, derandomize=True)
    @given(st.lists(seg, max_size=6))
    def check(segments):
        as_list = [x for s in segments for x in s.tolist()]
        packed = cm.pack_numbers(segments)
        assert packed == msgpack.packb(as_list)
        back = cm.unpack_numbers(packed)
        want = np.asarray([float(x) for x in as_list], dtype=np.float64)
        assert back.shape == want.shape and np.array_equal(back, want, equal_nan=True)

    check()


def test_reference_checkpoint_reader():
    """TimestepBuffer.load_reference on a checkpoint written by the REFERENCE's TimestepBuffer.save (frozen by
    oracle/gen_golden.py: wrapped 48-slot ring, three streams): every stored step comes back, in creation order per
    stream, unfinished trajectories truncated onto their in-flight observation."""
    from prism_b200.experience.timestep_buffer import TimestepBuffer
    gold_dir = os.path.join(os.path.dirname(__file__), "golden")
    fx = np.load(os.path.join(gold_dir, "ref_checkpoint.npz"))
    S = {k[len("script."):]: fx[k] for k in fx.files if k.startswith("script.")}

    class Fake(object):
        """The attributes load_reference touches, without a device."""

        def __init__(self):
            self.rows, self.emptied, self.flushed = None, 0, 0
            ring = type("Ring", (), {"max_streams": 8})()
            self.buffer = type("TD", (), {"_storage_opts": {"max_streams": 8}, "_storage": ring})()

        def empty(self):
            self.emptied += 1

        def extend_batch(self, *rows):
            self.rows = rows
            return len(rows[0])

        def _flush(self):
            self.flushed += 1

    fake = Fake()
    n = TimestepBuffer.load_reference(fake, os.path.join(gold_dir, "ref_checkpoint"))
    stored = fx["stored_ids"]
    assert n == len(stored) == int(fx["capacity"]) and fake.emptied == 1 and fake.flushed == 1
    assert fake._free_streams == list(range(7, -1, -1))            # every trajectory in the file is closed
    sid, obs, action, reward, done, trunc, next_obs = fake.rows
    t_of_id = {int(i): t for t, i in enumerate(fx["step_ids"])}
    key = {S["obs"][t].tobytes(): t for t in range(len(S["stream"]))}
    t_of = np.asarray([key[o.reshape(-1).tobytes()] for o in obs])
    assert sorted(t_of.tolist()) == sorted(t_of_id[int(i)] for i in stored)
    assert np.array_equal(action, S["action"][t_of]) and np.array_equal(reward, S["reward"][t_of])
    assert np.array_equal(done, S["done"][t_of])
    # per stream: creation order, one stream id per episode; the newest unfinished step of a stream was closed by the writer
    last = {}
    for j, t in enumerate(t_of):
        s = int(S["stream"][t])
        assert s not in last or t > last[s]
        last[s] = t
    tails = {t for t in last.values() if not S["done"][t] and not S["trunc"][t]}
    assert tails, "the fixture must hold at least one unfinished trajectory"
    for j, t in enumerate(t_of):
        assert trunc[j] == (bool(S["trunc"][t]) or t in tails)
        if S["done"][t]:
            continue
        want = S["final_obs"][t] if S["trunc"][t] else S["next_obs"][t]
        assert np.array_equal(next_obs[j].reshape(-1), want)
    # anything but numbers in the pickle is refused
    import pickle
    import tempfile
    with tempfile.TemporaryDirectory(dir=gold_dir) as d:
        os.makedirs(os.path.join(d, "experience_buffer"))
        with open(os.path.join(d, "experience_buffer", "timesteps.pkl"), "wb") as f:
            pickle.dump([1, 2.0, Config()], f)
        with pytest.raises(pickle.UnpicklingError):
            TimestepBuffer.load_reference(Fake(), d)


def test_reference_self_check_chain_through_the_object_path():
    """The chain of the reference's own ``complex_save_load_test`` (timestep_buffer.py:397-475: 26 steps, an episode end
    every 5 below 20, a truncation every 12 from 10 on, observations changing shape after the first step) serialized and
    rebuilt: ids, observations, flags, prev / next identities and needs_n_step survive; the newest step waits for its
    in-flight successor."""
    import weakref
    steps, t = [], Timestep(1, obs=torch.zeros(3, 4))
    keep = [t]
    for i in range(1, 27):
        t.reward, t.action = i, torch.ones(1)
        t.done = 20 > i > 0 and i % 5 == 0
        t.truncated = i >= 10 and i % 12 == 0
        nt = Timestep(i + 1, obs=torch.ones(2, 42, 42) * i)
        if t.truncated:
            t.next = Timestep(113 * (i + 1), obs=torch.ones(2, 42, 42) * -i)
            t.next.prev = weakref.ref(t)
        elif not t.done:
            nt.prev, t.next = weakref.ref(t), weakref.ref(nt)
        steps.append(t)
        keep.append(nt)
        t = nt
    flat = []
    for s in steps:
        flat += s.serialize()
    back, waiting = Timestep.deserialize_linked_list(NONE.unpack(NONE.pack(flat)))
    assert sorted(waiting) == [26] and waiting[26][1] == [None, None, 27]
    by_id = {s.id: s for s in back}
    by_id[26] = waiting[26][0]
    assert sorted(by_id) == list(range(1, 27))
    for bef in steps:
        aft = by_id[bef.id]
        assert torch.equal(torch.as_tensor(bef.obs), aft.obs) and aft.reward == float(bef.reward) and aft.action == 1
        assert (bef.done, bef.truncated, bef.needs_n_step) == (aft.done, aft.truncated, aft.needs_n_step)
        if bef.truncated:
            assert isinstance(aft.next, Timestep) and aft.next.id == bef.next.id
            assert torch.equal(aft.next.obs, bef.next.obs) and aft.next.prev() is aft
        elif bef.next is not None and bef.id != 26:
            assert aft.next() is by_id[bef.next().id]
        else:
            assert aft.next is None
        if bef.prev is not None:
            assert aft.prev() is by_id[bef.prev().id]
        else:
            assert aft.prev is None
