"""Key layout and message flow of the Redis transport -- same keys, commands and method names as the reference's
``RedisInterface`` (prism/async_components/redis/redis_interface.py:6-137), so either side of the wire can be a
reference process.  Numeric payloads go through the native codec (``*_array`` methods); the list-returning reference
methods are kept on top of them.  ``client``: any object with the redis-py calls used here (get / set / lpush / lrange
/ ltrim / delete / incrby / pipeline / flushall); by default ``redis.Redis(host, port)`` -- the ``redis`` module is
required for that and its absence raises at construction."""
import time

import numpy as np

from .. import compression_methods


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
            from redis import Redis                  # no fallback: the transport IS Redis
            client = Redis(host=host, port=port)
        self.redis = client
        self.serializer = compression_methods.MessageSerializer()
        self.max_queue_size = 100_000_000
        self.last_known_epoch = None
        self.waiting_timestep_id_map = {}

    # ---- scalars -------------------------------------------------------------------------------------------------
    def get_training_reward(self):
        return self.redis.get(RedisInterface.TRAINING_REWARD_KEY)

    def set_training_reward(self, reward):
        self.redis.set(RedisInterface.TRAINING_REWARD_KEY, reward)

    def get_current_command(self):
        command = self.redis.get(RedisInterface.CURRENT_COMMAND_KEY)
        return command.decode() if isinstance(command, bytes) else command

    def set_current_command(self, command):
        self.redis.set(RedisInterface.CURRENT_COMMAND_KEY, command)

    def get_config(self):
        serialized_config = self.redis.get(RedisInterface.CONFIG_KEY)
        if serialized_config is not None:
            from ...config import Config
            return Config.deserialize(serialized_config)
        return None

    def set_config(self, config):
        self.redis.set(RedisInterface.CONFIG_KEY, config.serialize() if hasattr(config, "serialize") else config)

    # ---- training batches ------------------------------------------------------------------------------------------
    def _pop_batches(self):
        pipe = self.redis.pipeline()
        pipe.lrange(RedisInterface.TRAINING_BATCH_KEY, 0, self.max_queue_size)
        pipe.ltrim(RedisInterface.TRAINING_BATCH_KEY, self.max_queue_size, -1)
        data = pipe.execute()[0]
        return data if data else None

    def get_waiting_batch_arrays(self):
        data = self._pop_batches()
        return None if data is None else [self.serializer.unpack_numbers(batch) for batch in data]

    def get_waiting_batches(self):
        data = self._pop_batches()
        return None if data is None else [self.serializer.unpack(batch) for batch in data]

    def add_batch(self, batch):
        self.redis.lpush(RedisInterface.TRAINING_BATCH_KEY, self.serializer.pack(batch))

    def add_batch_segments(self, segments):
        self.redis.lpush(RedisInterface.TRAINING_BATCH_KEY, self.serializer.pack_numbers(segments))

    # ---- environment description -------------------------------------------------------------------------------------
    def get_env_info(self, poll_seconds=1.0):
        env_info_vector = self.redis.get(RedisInterface.ENV_INFO_KEY)
        while env_info_vector is None:
            time.sleep(poll_seconds)
            env_info_vector = self.redis.get(RedisInterface.ENV_INFO_KEY)
        env_info_vector = self.serializer.unpack(env_info_vector)
        n_elements_in_shape = env_info_vector[0]
        obs_shape = env_info_vector[1:n_elements_in_shape + 1]
        n_acts = env_info_vector[n_elements_in_shape + 1]
        n_agents = env_info_vector[n_elements_in_shape + 2]
        return obs_shape, n_acts, n_agents

    def set_env_info(self, obs_shape, n_acts, n_agents):
        env_info_vector = [len(obs_shape), *obs_shape, n_acts, n_agents]
        self.redis.set(RedisInterface.ENV_INFO_KEY, self.serializer.pack(env_info_vector))

    # ---- model parameters ----------------------------------------------------------------------------------------------
    def get_latest_model(self):
        current_epoch = self.redis.get(RedisInterface.CURRENT_EPOCH_KEY)
        if current_epoch is not None:
            current_epoch = int(current_epoch)
            if current_epoch != self.last_known_epoch:
                serialized_model_params = self.redis.get(RedisInterface.MODEL_PARAMS_KEY)
                self.last_known_epoch = current_epoch
                return self.serializer.unpack(serialized_model_params)
        return None

    def set_latest_model(self, serialized_model, current_epoch):
        if isinstance(serialized_model, np.ndarray):
            packed = self.serializer.pack_numbers([serialized_model])
        else:
            packed = self.serializer.pack(serialized_model)
        pipe = self.redis.pipeline()
        pipe.set(RedisInterface.CURRENT_EPOCH_KEY, current_epoch)
        pipe.set(RedisInterface.MODEL_PARAMS_KEY, packed)
        pipe.get(RedisInterface.TOTAL_TIMESTEPS_COLLECTED_KEY)
        total_timesteps = pipe.execute()[-1]
        return 0 if total_timesteps is None else int(total_timesteps)

    # ---- collector steps -------------------------------------------------------------------------------------------------
    def submit_timesteps(self, timesteps):
        from .. import wire
        packed_timesteps = self.serializer.pack_numbers(wire.timestep_segments(timesteps))
        pipe = self.redis.pipeline()
        pipe.lpush(RedisInterface.TIMESTEPS_KEY, packed_timesteps)
        pipe.ltrim(RedisInterface.TIMESTEPS_KEY, 0, self.max_queue_size)
        pipe.incrby(RedisInterface.TOTAL_TIMESTEPS_COLLECTED_KEY, len(timesteps))
        pipe.execute()

    def _pop_timestep_blocks(self):
        pipe = self.redis.pipeline()
        pipe.lrange(RedisInterface.TIMESTEPS_KEY, 0, -1)
        pipe.delete(RedisInterface.TIMESTEPS_KEY)
        return pipe.execute()[0] or []

    def get_timestep_arrays(self):
        """The queued blocks as float64 arrays, OLDEST FIRST.  (``lpush`` puts the newest block at the head of the
        list and the reference reads head to tail, i.e. newest first -- harmless for its id-keyed link map; the
        stream decoder wants arrival order.)"""
        return [self.serializer.unpack_numbers(block) for block in reversed(self._pop_timestep_blocks())]

    def get_timesteps(self):
        serialized_timesteps = []
        for packed_list in self._pop_timestep_blocks():
            serialized_timesteps += self.serializer.unpack(packed_list)
        return serialized_timesteps

    def clear_redis(self):
        self.redis.flushall()
