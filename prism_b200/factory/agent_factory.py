"""build_agent(config, obs_shape, n_actions) -> Agent, as prism/factory/agent_factory.py:7-61.
Adam becomes FlatAdam (fused clip + Adam over a flat arena); online and target parameters are laid
out in identical arenas so the target sync is one copy."""
import torch

from ..agents import Agent, action_selectors
from ..agents.optim import FlatAdam, flatten_parameters
from . import model_factory


def build_agent(config, obs_shape, n_actions):
    use_cuda_graph = config.use_cuda_graph and "cuda" in config.device
    obs_shape = [int(arg) for arg in obs_shape]
    n_actions = int(n_actions)
    model = model_factory.create_model(obs_shape, n_actions, config)
    eval_action_selector = action_selectors.GreedyActionSelector()

    target_model = None
    if config.use_target_network:
        target_model = model_factory.create_model(obs_shape, n_actions, config)
        target_model.load_state_dict(model.state_dict())

    on_cuda = "cuda" in config.device
    arena = flatten_parameters(model) if on_cuda else None
    if target_model is not None and on_cuda:
        flatten_parameters(target_model)
        for p in target_model.parameters():
            p.requires_grad_(False)

    if config.use_ids:
        action_selector = action_selectors.IDSActionSelector(config.ids_lambda, config.ids_use_random_samples,
                                                             config.ids_epsilon, config.ids_rho_lower_bound,
                                                             config.ids_beta, None)
    elif config.use_e_greedy:
        action_selector = action_selectors.EGreedyActionSelector(config.e_greedy_initial_epsilon,
                                                                 config.e_greedy_final_epsilon,
                                                                 config.e_greedy_decay_timesteps, config.seed)
    else:
        action_selector = action_selectors.GreedyActionSelector()

    if config.use_adam:
        optimizer = FlatAdam(model.parameters(), lr=config.learning_rate,
                             betas=(config.adam_beta1, config.adam_beta2), eps=config.adam_epsilon,
                             max_grad_norm=config.max_grad_norm, arena=arena)
    elif config.use_rmsprop:
        optimizer = torch.optim.RMSprop(model.parameters(), lr=config.learning_rate, alpha=config.rmsprop_alpha,
                                        centered=True, eps=config.rmsprop_epsilon, capturable=use_cuda_graph)
    else:
        optimizer = torch.optim.SGD(model.parameters(), lr=config.learning_rate)

    return Agent(model, action_selector, eval_action_selector, optimizer, target_model, use_cuda_graph,
                 config.max_grad_norm)
