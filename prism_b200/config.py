"""Hyper-parameters of the learner hot path.

The reference keeps one flat 91-field dataclass without defaults
(prism/config/algorithm_configuration.py:5-123) and builds named configs by copy-and-mutate.
The factories in prism_b200.factory read the SAME field names by attribute, so either a reference
``Config`` object or this dataclass can be passed.  Defaults below are the reference's
``DEFAULT_CONFIG`` values (prism/config/default_config.py:3-112) for the fields the hot path reads;
the named constructors reproduce the three BASELINE.json model configurations.
Fields beyond the reference's (device placement / storage / sampling mode of the device buffer)
default to reference behaviour.
"""
from dataclasses import dataclass, replace


@dataclass
class Config:
    # loss / optimiser
    distributional_loss_weight: float = 1
    q_loss_weight: float = 1
    batch_size: int = 32
    gamma: float = 0.99
    learning_rate: float = 6.25e-5
    max_grad_norm: float = 10.0
    sparse_init_p: float = 0.0
    loss_squish_fn_id: str = "none"
    use_adam: bool = True
    adam_beta1: float = 0.9
    adam_beta2: float = 0.999
    adam_epsilon: float = 1.5e-4
    use_rmsprop: bool = False
    rmsprop_alpha: float = 0.95
    rmsprop_epsilon: float = 0.01
    q_loss_fn: str = "mse"
    # embedding
    embedding_model_final_dim: int = 3136
    embedding_model_layer_sizes: int = 512
    embedding_model_num_layers: int = 0
    embedding_model_type: str = "nature_atari_cnn"
    embedding_model_act_fn_id: str = "relu"
    # IDS
    use_ids: bool = True
    ids_beta: float = 1
    ids_use_random_samples: bool = False
    ids_lambda: float = 0.1
    ids_n_q_heads: int = 10
    ids_q_head_feature_dim: int = 512
    ids_n_q_head_model_layers: int = 2
    ids_allow_distributional_gradients: bool = True
    ids_rho_lower_bound: float = 0.25
    ids_epsilon: float = 1e-10
    ids_ensemble_variation_coef: float = 1e-6
    # e-greedy
    use_e_greedy: bool = False
    e_greedy_initial_epsilon: float = 1.0
    e_greedy_final_epsilon: float = 0.01
    e_greedy_decay_timesteps: int = 50_000_000
    # returns / norm
    n_step_returns_length: int = 3
    use_layer_norm: bool = True
    # replay
    use_experience_replay: bool = True
    experience_replay_capacity: int = 1_000_000
    use_per: bool = False
    per_alpha: float = 0.5
    per_beta_start: float = 0.5
    per_beta_end: float = 0.5
    per_beta_anneal_timesteps: int = 1
    # IQN
    use_iqn: bool = True
    iqn_n_current_state_quantile_samples: int = 8
    iqn_n_next_state_quantile_samples: int = 8
    iqn_quantile_samples_per_action: int = 200
    iqn_n_basis_elements: int = 64
    iqn_quantile_model_feature_dim: int = 512
    iqn_quantile_model_layers: int = 1
    iqn_risk_policy_id: str = "neutral"
    iqn_huber_loss_kappa: float = 1.0
    # DQN
    use_dqn: bool = False
    dqn_n_model_layers: int = 1
    dqn_n_model_feature_dim: int = 512
    use_c51: bool = False
    use_double_q_learning: bool = False
    use_target_network: bool = False
    target_update_period: int = 8_000
    seed: int = 123
    device: str = "cuda:0"
    use_cuda_graph: bool = True
    frame_stack_size: int = 4
    run_through_redis: bool = False
    redis_host: str = "localhost"
    redis_port: int = 6379
    redis_side: str = "server"
    # ---- device-buffer knobs (new; defaults reproduce reference behaviour) ----
    per_sampling: str = "iid"               # "iid" (torchrl) | "stratified" (north star)
    replay_storage_dtype: str = "float32"   # "float32" | "uint8"
    replay_obs_scale_255: bool = False      # uint8 storage: emit v/255 (gymnasium scale_obs)
    replay_max_streams: int = 256
    replay_staging_rows: int = 256
    redis_local_buffer: bool = False        # learner drains the step blocks into its own device buffer (no batch hop)

    # ---- wire form (prism/config/algorithm_configuration.py:117-123): JSON of the fields ----
    def serialize(self):
        import json
        return json.dumps(self.__dict__)

    @classmethod
    def deserialize(cls, serialized_config):
        """Accepts a config serialized by either implementation: fields this dataclass does not declare (the
        reference has 91) are kept as plain attributes, since the factories read by attribute."""
        import json
        from dataclasses import fields
        if isinstance(serialized_config, (bytes, bytearray)):
            serialized_config = serialized_config.decode("utf-8")
        cfg_json = dict(json.loads(serialized_config))
        known = {f.name for f in fields(cls)}
        cfg = cls(**{k: v for k, v in cfg_json.items() if k in known})
        for k, v in cfg_json.items():
            if k not in known:
                setattr(cfg, k, v)
        return cfg


def minatar_ids_iqn_config(**kw):
    """BASELINE config 1: SUBTRACTIVE_ABLATION_BASE_CONFIG (prism/config/subtractive_ablation_base_config.py:2-56)
    + target net (period 4000, subtractive_ablation_experiment.py:37-40) + PER."""
    cfg = Config(
        use_iqn=True, iqn_n_current_state_quantile_samples=32, iqn_n_next_state_quantile_samples=32,
        iqn_quantile_samples_per_action=32, iqn_n_basis_elements=64, iqn_quantile_model_feature_dim=256,
        iqn_quantile_model_layers=1, use_e_greedy=False, use_ids=True, ids_n_q_head_model_layers=2,
        ids_n_q_heads=10, ids_q_head_feature_dim=256, ids_ensemble_variation_coef=1e-6, use_layer_norm=True,
        use_target_network=True, target_update_period=4000, use_double_q_learning=False, use_per=True,
        per_beta_start=0.5, per_beta_end=0.5, per_alpha=0.5, n_step_returns_length=3,
        embedding_model_type="minatar_cnn", experience_replay_capacity=3_000_000, learning_rate=1e-4, gamma=0.99,
        batch_size=64, frame_stack_size=1, adam_epsilon=0.0003125)
    return replace(cfg, **kw)


def minatar_dqn_per_config(**kw):
    """BASELINE config 2: DQN block of REVISITING_RAINBOW_MINATAR_CONFIG
    (prism/config/revisiting_rainbow_minitar_config.py:40-54) + double-Q + PER, 1M capacity, batch 256."""
    cfg = Config(
        use_per=True, use_ids=False, use_iqn=False, use_layer_norm=False, use_double_q_learning=True, use_dqn=True,
        use_target_network=True, target_update_period=1000, use_e_greedy=True, learning_rate=0.00025,
        e_greedy_decay_timesteps=250_000, e_greedy_final_epsilon=0.01, dqn_n_model_feature_dim=256,
        dqn_n_model_layers=2, n_step_returns_length=3, embedding_model_type="minatar_cnn",
        experience_replay_capacity=1_000_000, batch_size=256, frame_stack_size=1, gamma=0.99,
        adam_epsilon=0.0003125)
    return replace(cfg, **kw)


def atari_iqn_ids_config(**kw):
    """BASELINE config 5: DEFAULT_CONFIG (prism/config/default_config.py:3-112) with 64x64 quantile samples,
    batch 512, PER on."""
    cfg = Config(iqn_n_current_state_quantile_samples=64, iqn_n_next_state_quantile_samples=64, batch_size=512,
                 use_per=True)
    return replace(cfg, **kw)
