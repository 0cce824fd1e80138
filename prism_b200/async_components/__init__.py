"""Remote-actor transport: the reference's Redis wire formats in front of the device buffer."""
from .async_experience_buffer import AsyncExperienceBuffer, AsyncExperienceBufferInterface
