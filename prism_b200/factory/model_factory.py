"""create_model(env_state_shape, env_n_actions, config) -> CompositeModel.
Same wiring as the reference (prism/factory/model_factory.py:49-153), reading the same Config fields."""
import torch
import torch.nn as nn

from ..agents.models import CompositeModel, FFNNModel, IQNModel, MinAtarModel, NatureAtariCnn, QEnsemble

_ACTIVATIONS = {"tanh": nn.Tanh, "gelu": nn.GELU, "sigmoid": nn.Sigmoid, "swish": nn.SiLU}


def _squish_id(config):
    sq = getattr(config, "loss_squish_fn_id", "none")
    # two reference configs carry a trailing comma and hold the 1-tuple ("none",)
    return sq[0] if isinstance(sq, tuple) else sq


def create_model(env_state_shape, env_n_actions, config):
    if _squish_id(config) in ("obs_look_further", "symlog"):
        raise NotImplementedError("value squashing is outside the hot-path scope (no BASELINE config uses it)")
    act_fn = _ACTIVATIONS.get(config.embedding_model_act_fn_id, nn.ReLU)
    device = config.device
    use_cuda_graph = config.use_cuda_graph and "cuda" in device

    embedding = None
    kind = config.embedding_model_type
    if "ffnn" in kind:
        embedding = FFNNModel(n_input_features=env_state_shape[-1],
                              n_output_features=config.embedding_model_final_dim,
                              n_layers=config.embedding_model_num_layers,
                              layer_width=config.embedding_model_layer_sizes,
                              use_layer_norm=config.use_layer_norm, apply_layer_norm_first_layer=False,
                              output_act_fn=act_fn, sparse_init_p=config.sparse_init_p, act_fn=act_fn, device=device)
    elif kind == "nature_atari_cnn":
        embedding = NatureAtariCnn(frame_stack=config.frame_stack_size, feature_dim=config.embedding_model_final_dim,
                                   act_fn=act_fn, device=device, sparse_init_p=config.sparse_init_p,
                                   use_layer_norm=config.use_layer_norm)
        config.embedding_model_final_dim = embedding.output_dim
    elif kind == "minatar_cnn":
        embedding = MinAtarModel(in_channels=env_state_shape[-1], act_fn=act_fn, device=device,
                                 sparse_init_p=config.sparse_init_p, use_layer_norm=config.use_layer_norm)
        config.embedding_model_final_dim = embedding.output_dim
    feat = config.embedding_model_final_dim

    iqn = None
    if config.use_iqn:
        propagate = (config.ids_allow_distributional_gradients and config.use_ids) or not config.use_ids
        iqn = IQNModel(n_input_features=feat, n_actions=env_n_actions, n_basis_elements=config.iqn_n_basis_elements,
                       use_layer_norm=config.use_layer_norm, n_model_layers=config.iqn_quantile_model_layers,
                       model_layer_size=config.iqn_quantile_model_feature_dim, model_activation=act_fn,
                       use_double_q_learning=config.use_double_q_learning, sparse_init_p=config.sparse_init_p,
                       huber_k=config.iqn_huber_loss_kappa, distributional_loss_weight=config.distributional_loss_weight,
                       n_current_quantile_samples=config.iqn_n_current_state_quantile_samples,
                       n_next_quantile_samples=config.iqn_n_next_state_quantile_samples,
                       n_quantile_samples_per_action=config.iqn_quantile_samples_per_action,
                       propagate_grad=propagate, device=device)

    q_model = None
    mse = nn.MSELoss(reduction="none")   # "huber" falls through to MSE in the reference too (:43-46)
    if config.use_ids:
        q_model = QEnsemble(n_input_features=feat, n_actions=env_n_actions, n_heads=config.ids_n_q_heads,
                            use_layer_norm=config.use_layer_norm, n_model_layers=config.ids_n_q_head_model_layers,
                            model_layer_size=config.ids_q_head_feature_dim, sparse_init_p=config.sparse_init_p,
                            model_activation=act_fn, use_double_q_learning=config.use_double_q_learning,
                            q_loss_function=mse, q_loss_weight=config.q_loss_weight,
                            ensemble_variation_coef=config.ids_ensemble_variation_coef, device=device)
    elif config.use_dqn:
        q_model = QEnsemble(n_input_features=feat, n_actions=env_n_actions, n_heads=1,
                            use_layer_norm=config.use_layer_norm, n_model_layers=config.dqn_n_model_layers,
                            model_layer_size=config.dqn_n_model_feature_dim, sparse_init_p=config.sparse_init_p,
                            model_activation=act_fn, use_double_q_learning=config.use_double_q_learning,
                            q_loss_function=mse, q_loss_weight=config.q_loss_weight, ensemble_variation_coef=0,
                            device=device)

    return CompositeModel(embedding_model=embedding, distribution_model=iqn, q_function_model=q_model,
                          device=device, use_cuda_graph=use_cuda_graph)
