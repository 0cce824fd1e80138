"""Redis side of the remote-actor transport: which key holds what, and in which form.

Wire contract shared with the reference's ``RedisInterface`` (prism/async_components/redis/redis_interface.py:6-137;
same key strings, same list / scalar usage, same method names), so that either end may be a reference process:

    key                         type    written by                     payload
    "timesteps"                 list    collectors (lpush, newest at   envelope(flat Timestep block)
                                        the head)
    "training_batch"            list    the buffer process (lpush)     envelope(flat batch)
    "model" / "current_epoch"   scalar  learner                        envelope(parameter list) / int
    "config"                    scalar  learner                        Config JSON
    "env_info"                  scalar  collector                      envelope([len(shape), *shape, n_acts, n_agents])
    "current_command", "training_reward", "total_timesteps_collected"  plain scalars

Numeric payloads are encoded / decoded by the native codec (``get_timestep_arrays``, ``get_waiting_batch_arrays``,
``add_batch_segments``; csrc/wire.cu); the list-returning reference methods remain for callers that want Python lists.
``client``: any object with the redis-py calls used here (get, set, lpush, lrange, ltrim, delete, incrby, pipeline,
flushall).  Default: ``redis.Redis(host, port)`` -- the ``redis`` module is then required; there is no fallback.
"""
import time

import numpy as np

from .. import compression_methods, wire


class RedisInterface(object):
    TIMESTEPS_KEY = "timesteps"
    MODEL_PARAMS_KEY = "model"
    CURRENT_EPOCH_KEY = "current_epoch"
    CONFIG_KEY = "config"
    CURRENT_COMMAND_KEY = "current_command"
    TOTAL_TIMESTEPS_COLLECTED_KEY = "total_timesteps_collected"
    ENV_INFO_KEY = "env_info"
    TRAINING_REWARD_KEY = "training_reward"
    TRAINING_BATCH_KEY = "training_batch"

    START_COLLECTING_COMMAND = "start_collecting"
    SHUTDOWN_COMMAND = "shutdown"

    def __init__(self, host='localhost', port=6379, client=None):
        if client is None:
            import redis as redis_py
            client = redis_py.Redis(host=host, port=port)
        self.redis = client
        self.serializer = compression_methods.MessageSerializer()
        self.max_queue_size = 100_000_000
        self.last_known_epoch = None
        self.waiting_timestep_id_map = {}

    # ---- helpers -----------------------------------------------------------------------------------------------------
    def _atomically(self, *ops):
        """Run ``(method, *args)`` tuples as one pipeline; returns the list of results."""
        pipe = self.redis.pipeline()
        for name, *args in ops:
            getattr(pipe, name)(*args)
        return pipe.execute()

    def _take_all(self, key):
        """Every queued message of a list key (head first, i.e. newest first), removing them."""
        return self._atomically(("lrange", key, 0, -1), ("delete", key))[0] or []

    # ---- plain scalars -------------------------------------------------------------------------------------------------
    def get_training_reward(self):
        return self.redis.get(self.TRAINING_REWARD_KEY)

    def set_training_reward(self, reward):
        self.redis.set(self.TRAINING_REWARD_KEY, reward)

    def get_current_command(self):
        raw = self.redis.get(self.CURRENT_COMMAND_KEY)            # redis-py hands back bytes
        return raw.decode() if isinstance(raw, (bytes, bytearray)) else raw

    def set_current_command(self, command):
        self.redis.set(self.CURRENT_COMMAND_KEY, command)

    def get_config(self):
        raw = self.redis.get(self.CONFIG_KEY)
        if raw is None:
            return None
        from ...config import Config
        return Config.deserialize(raw)

    def set_config(self, config):
        self.redis.set(self.CONFIG_KEY, config if isinstance(config, (str, bytes)) else config.serialize())

    # ---- environment description ---------------------------------------------------------------------------------------
    def set_env_info(self, obs_shape, n_acts, n_agents):
        self.redis.set(self.ENV_INFO_KEY, self.serializer.pack([len(obs_shape), *obs_shape, n_acts, n_agents]))

    def get_env_info(self, poll_seconds=1.0):
        while True:
            raw = self.redis.get(self.ENV_INFO_KEY)
            if raw is not None:
                break
            time.sleep(poll_seconds)
        info = self.serializer.unpack(raw)
        rank = info[0]
        return info[1:1 + rank], info[1 + rank], info[2 + rank]

    # ---- model parameters ------------------------------------------------------------------------------------------------
    def set_latest_model(self, serialized_model, current_epoch):
        """Publishes the parameters under a new epoch; returns the collectors' running step count."""
        if isinstance(serialized_model, np.ndarray):
            message = self.serializer.pack_numbers([serialized_model])
        else:
            message = self.serializer.pack(serialized_model)
        collected = self._atomically(("set", self.CURRENT_EPOCH_KEY, current_epoch), ("set", self.MODEL_PARAMS_KEY, message),
                                     ("get", self.TOTAL_TIMESTEPS_COLLECTED_KEY))[-1]
        return int(collected) if collected is not None else 0

    def get_latest_model(self):
        """The parameter list if the learner published a newer epoch than the last one seen, else None."""
        epoch = self.redis.get(self.CURRENT_EPOCH_KEY)
        if epoch is None or int(epoch) == self.last_known_epoch:
            return None
        self.last_known_epoch = int(epoch)
        return self.serializer.unpack(self.redis.get(self.MODEL_PARAMS_KEY))

    # ---- collector steps -----------------------------------------------------------------------------------------------------
    def submit_timesteps(self, timesteps):
        message = self.serializer.pack_numbers(wire.timestep_segments(timesteps))
        self._atomically(("lpush", self.TIMESTEPS_KEY, message), ("ltrim", self.TIMESTEPS_KEY, 0, self.max_queue_size),
                         ("incrby", self.TOTAL_TIMESTEPS_COLLECTED_KEY, len(timesteps)))

    def get_timestep_arrays(self):
        """The queued blocks as float64 arrays, OLDEST FIRST.  (The reference concatenates them newest first, which its
        id-keyed link map does not mind; the stream decoder wants arrival order.)"""
        return [self.serializer.unpack_numbers(m) for m in reversed(self._take_all(self.TIMESTEPS_KEY))]

    def get_timesteps(self):
        """Reference form: one flat Python list, blocks in queue order (newest first)."""
        flat = []
        for m in self._take_all(self.TIMESTEPS_KEY):
            flat.extend(self.serializer.unpack(m))
        return flat

    # ---- training batches ------------------------------------------------------------------------------------------------------
    def add_batch(self, batch):
        self.redis.lpush(self.TRAINING_BATCH_KEY, self.serializer.pack(batch))

    def add_batch_segments(self, segments):
        self.redis.lpush(self.TRAINING_BATCH_KEY, self.serializer.pack_numbers(segments))

    def _take_batches(self):
        cap = self.max_queue_size
        return self._atomically(("lrange", self.TRAINING_BATCH_KEY, 0, cap), ("ltrim", self.TRAINING_BATCH_KEY, cap, -1))[0]

    def get_waiting_batch_arrays(self):
        messages = self._take_batches()
        return [self.serializer.unpack_numbers(m) for m in messages] if messages else None

    def get_waiting_batches(self):
        messages = self._take_batches()
        return [self.serializer.unpack(m) for m in messages] if messages else None

    def clear_redis(self):
        self.redis.flushall()
