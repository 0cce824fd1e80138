from . import model_factory, agent_factory, exp_buffer_factory
from .model_factory import create_model
from .agent_factory import build_agent
from .exp_buffer_factory import build_exp_buffer
