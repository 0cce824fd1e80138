from .agent import Agent
from . import action_selectors
from .optim import FlatAdam, flatten_parameters
