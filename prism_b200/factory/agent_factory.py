"""build_agent(config, obs_shape, n_actions) -> Agent: the seam of prism/factory/agent_factory.py:7-61.

Same decisions from the same config fields (target network, IDS / e-greedy / greedy acting policy, Adam / RMSprop / SGD);
what differs is what gets built: Adam becomes ``FlatAdam`` (clip + Adam fused over one flat arena), and the online and
target parameters are laid out in identical arenas so the target sync is a single copy.
"""
import torch

from ..agents import Agent, action_selectors
from ..agents.optim import FlatAdam, flatten_parameters
from . import model_factory


def _acting_policy(cfg):
    """agent_factory.py:20-38.  Value squashing is unused by every BASELINE config: no unsquish function."""
    if cfg.use_ids:
        return action_selectors.IDSActionSelector(cfg.ids_lambda, cfg.ids_use_random_samples, cfg.ids_epsilon,
                                                  cfg.ids_rho_lower_bound, cfg.ids_beta, None)
    if cfg.use_e_greedy:
        return action_selectors.EGreedyActionSelector(cfg.e_greedy_initial_epsilon, cfg.e_greedy_final_epsilon,
                                                      cfg.e_greedy_decay_timesteps, cfg.seed)
    return action_selectors.GreedyActionSelector()


def _optimizer(cfg, model, arena, graphed):
    """agent_factory.py:40-58."""
    params = model.parameters()
    if cfg.use_adam:
        return FlatAdam(params, lr=cfg.learning_rate, betas=(cfg.adam_beta1, cfg.adam_beta2), eps=cfg.adam_epsilon,
                        max_grad_norm=cfg.max_grad_norm, arena=arena)
    if cfg.use_rmsprop:
        return torch.optim.RMSprop(params, lr=cfg.learning_rate, alpha=cfg.rmsprop_alpha, eps=cfg.rmsprop_epsilon,
                                   centered=True, capturable=graphed)
    return torch.optim.SGD(params, lr=cfg.learning_rate)


def build_agent(config, obs_shape, n_actions):
    shape, n_actions = [int(d) for d in obs_shape], int(n_actions)
    on_cuda = "cuda" in config.device
    graphed = bool(config.use_cuda_graph) and on_cuda

    online = model_factory.create_model(shape, n_actions, config)
    arena = flatten_parameters(online) if on_cuda else None

    target = None
    if config.use_target_network:                      # a frozen twin, same arena layout as the online network
        target = model_factory.create_model(shape, n_actions, config)
        target.load_state_dict(online.state_dict())
        if on_cuda:
            flatten_parameters(target)
            for p in target.parameters():
                p.requires_grad_(False)

    optimizer = _optimizer(config, online, arena, graphed)
    if isinstance(optimizer, FlatAdam):
        optimizer.layout_model = online        # checkpoints in torch.optim.Adam's format, reference parameter order
    return Agent(online, _acting_policy(config), action_selectors.GreedyActionSelector(), optimizer, target, graphed,
                 config.max_grad_norm)
