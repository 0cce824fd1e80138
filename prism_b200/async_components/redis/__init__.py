from .redis_interface import RedisInterface
