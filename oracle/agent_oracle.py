"""CPU fp32 restatement of the reference's agent-side hot path -- TEST INFRASTRUCTURE ONLY.

Plain PyTorch ops on the CPU, written to follow the reference line by line (so that it can be
checked against the reference's own outputs, frozen in tests/golden/agent_*.npz by
oracle/gen_golden.py, and then serve as the checker for the CUDA path where /root/reference is not
available):
  * IQN forward / quantile-Huber loss      prism/agents/models/iqn_model.py:48-93, 95-201
  * Q-ensemble forward / MSE double-Q loss prism/agents/models/q_ensemble.py:44-92
  * loss glue + new priorities             prism/agents/models/composite_model.py:94-144
  * update step (clip + Adam)              prism/agents/agent.py:53-79, factory/agent_factory.py:44-47
  * IDS / greedy selection                 prism/agents/action_selectors.py:70-82, 125-176
Modules keep the reference's parameter names so a reference state_dict loads with strict=True.
"""
import math

import torch
import torch.nn as nn


def _mlp(n_in, n_out, n_layers, width, use_ln, ln_first=True, out_act=False):
    mods, dims = [], [n_in] + [width] * (n_layers - 1) + [n_out]
    for i in range(n_layers):
        if use_ln and (i > 0 or ln_first):
            mods.append(nn.LayerNorm(dims[i]))
        mods.append(nn.Linear(dims[i], dims[i + 1]))
        if i < n_layers - 1:
            mods.append(nn.ReLU())
    if out_act:
        mods.append(nn.ReLU())
    return nn.Sequential(*mods)


class _Wrap(nn.Module):
    """Gives a Sequential the reference's `.model` attribute name."""

    def __init__(self, seq, pre=None):
        super().__init__()
        self.model = seq
        self._pre = pre

    def forward(self, x):
        if self._pre is not None:
            x = self._pre(x)
        return self.model(x)


def minatar_embedding(in_channels):
    seq = nn.Sequential(nn.Conv2d(in_channels, 16, kernel_size=3, stride=1), nn.ReLU(), nn.Flatten())
    return _Wrap(seq, pre=lambda x: x.permute(0, 3, 1, 2).float())


def nature_embedding(frame_stack, use_ln):
    mods = [nn.Conv2d(frame_stack, 32, 8, 4), nn.ReLU()]
    if use_ln:
        mods.append(nn.LayerNorm((32, 20, 20)))
    mods += [nn.Conv2d(32, 64, 4, 2), nn.ReLU()]
    if use_ln:
        mods.append(nn.LayerNorm((64, 9, 9)))
    mods += [nn.Conv2d(64, 64, 3, 1), nn.ReLU(), nn.Flatten()]
    return _Wrap(nn.Sequential(*mods))


class OracleIQN(nn.Module):
    def __init__(self, n_feat, n_actions, n_basis, use_ln, n_layers, width, kappa, T, Tp, Nq, double_q,
                 loss_weight=1.0, propagate_grad=True):
        super().__init__()
        self.n_actions, self.n_basis, self.kappa = n_actions, n_basis, kappa
        self.T, self.Tp, self.Nq, self.double_q = T, Tp, Nq, double_q
        self.loss_weight, self.propagate_grad = loss_weight, propagate_grad
        self.phi = nn.Sequential(nn.Linear(n_basis, n_feat), nn.ReLU())
        self.model = None
        if n_layers > 0:
            self.model = _Wrap(_mlp(n_feat, width, n_layers, width, use_ln, out_act=True))
            n_feat = width
        if use_ln:
            self.embedding_to_quantile_layer = nn.Sequential(nn.LayerNorm(n_feat), nn.Linear(n_feat, n_actions))
        else:
            self.embedding_to_quantile_layer = nn.Linear(n_feat, n_actions)
        self.tau_queue = []  # injected quantile draws, consumed in call order

    def _tau(self, rows):
        if self.tau_queue:
            t = self.tau_queue.pop(0)
            assert t.numel() == rows
            return t.reshape(rows, 1).float()
        return torch.rand([rows, 1]).float()

    def forward(self, x, n, for_action=False):
        if not self.propagate_grad:
            x = x.detach()
        x = x.reshape(x.shape[0], -1)
        if for_action:
            n = self.Nq
        tau = self._tau(n * x.shape[0])
        tiled_x = torch.tile(x, [n, 1])                                          # iqn_model.py:70
        basis = torch.arange(1, self.n_basis + 1)
        emb = torch.cos(torch.tile(tau, [1, self.n_basis]) * basis * math.pi)    # :89-92
        h = self.phi(emb) * tiled_x                                              # :73
        if self.model is not None:
            h = self.model(h)
        z = self.embedding_to_quantile_layer(h)
        return z.view(n, -1, self.n_actions) if for_action else (z, tau)

    def get_loss(self, emb, emb_next, acts, returns, gdn, target=None):
        if not self.propagate_grad:
            emb, emb_next = emb.detach(), emb_next.detach()
        target = self if target is None else target
        B, T, Tp, k = acts.shape[0], self.T, self.Tp, self.kappa
        z_cur, tau = self.forward(emb, T)
        with torch.no_grad():
            if target is self:
                z_on = self.forward(emb_next, Tp)[0]
                z_tg = z_on
            elif self.double_q:
                z_on = self.forward(emb_next, Tp)[0]
                z_tg = target.forward(emb_next, Tp)[0]
            else:
                z_tg = target.forward(emb_next, Tp)[0]
                z_on = z_tg
            best = z_on.view(Tp, B, -1).mean(dim=0).argmax(dim=-1)                 # :129-133
            z_sel = torch.gather(z_tg, 1, torch.tile(best.view(-1, 1), [Tp, 1]))   # :136-139
            y = torch.tile(returns.view(-1, 1), [Tp, 1]) + z_sel * torch.tile(gdn.view(-1, 1), [Tp, 1])
            y = y.view(Tp, B, 1).transpose(1, 0)                                   # (B, T', 1)
        theta = torch.gather(z_cur, 1, torch.tile(acts.view(-1, 1), [T, 1])).view(T, B, 1).transpose(1, 0)
        delta = y[:, :, None] - theta[:, None, :]                                  # (B, T', T, 1)
        small = (delta.abs() <= k).float()
        huber = small * 0.5 * delta.square() + (1 - small) * k * (delta.abs() - 0.5 * k)
        tq = tau.view(T, B, 1).transpose(1, 0)[:, None, :, :]
        rho = (tq - (delta < 0).float().detach()).abs() * huber / k
        return rho.sum(dim=2).mean(dim=1).view(-1) * self.loss_weight              # :196-201


class OracleEnsemble(nn.Module):
    def __init__(self, n_feat, n_actions, K, use_ln, n_layers, width, double_q, loss_weight=1.0, theil_coef=0.0):
        super().__init__()
        heads = []
        for _ in range(K):
            if n_layers > 0:
                heads.append(_Wrap(_mlp(n_feat, n_actions, n_layers, width, use_ln)))
            elif use_ln:
                heads.append(nn.Sequential(nn.LayerNorm(n_feat), nn.Linear(n_feat, n_actions)))
            else:
                heads.append(nn.Linear(n_feat, n_actions))
        self.q_heads = nn.ModuleList(heads)
        self.double_q, self.loss_weight, self.theil_coef = double_q, loss_weight, theil_coef
        self.theil = torch.tensor(0.0)

    def forward(self, x):
        return torch.stack([h(x) for h in self.q_heads], dim=-1)                   # (B, A, K)

    def get_loss(self, emb, emb_next, acts, returns, gdn, target=None):
        target = self if target is None else target
        B = acts.shape[0]
        q_cur = self.forward(emb)
        with torch.no_grad():
            if target is self:
                q_on = q_tg = self.forward(emb_next)
            elif self.double_q:
                q_on, q_tg = self.forward(emb_next), target.forward(emb_next)
            else:
                q_tg = q_on = target.forward(emb_next)
            best = q_on.argmax(dim=-2)                                            # (B, K)
            rows = torch.arange(B)[:, None]
            q_sel = q_tg[rows, best, torch.arange(q_tg.shape[-1])[None, :]]
            y = returns.view(-1, 1) + q_sel * gdn.view(-1, 1)
        loss = (q_cur[torch.arange(B), acts.view(-1), :] - y).square().mean(dim=-1)
        if self.theil_coef != 0:
            l2 = torch.stack([nn.utils.parameters_to_vector(h.parameters()).norm() for h in self.q_heads])
            ratio = l2 / l2.mean()
            self.theil = (ratio * torch.log(ratio)).mean()
        return self.loss_weight * (loss - self.theil * self.theil_coef)


class OracleComposite(nn.Module):
    def __init__(self, embedding, iqn, ens):
        super().__init__()
        self.embedding_model = embedding
        self.distribution_model = iqn
        self.q_function_model = ens

    def forward(self, x):
        emb = self.embedding_model(x)
        z = self.distribution_model(emb, None, for_action=True) if self.distribution_model is not None else None
        if self.q_function_model is not None:
            q = self.q_function_model(emb)
        else:
            q = z.mean(dim=0).unsqueeze(-1)
        return q, z

    def get_losses(self, batch, target):
        obs, nobs = batch["observation"], batch["next"]["observation"]
        if obs.shape[1] == 1:
            obs, nobs = obs.squeeze(1), nobs.squeeze(1)
        returns = batch["next"]["reward"].flatten()
        gdn = batch["gamma"].flatten().float() * batch["nonterminal"].flatten().float()
        acts = batch["action"].flatten().long()
        emb = self.embedding_model(obs)
        with torch.no_grad():
            emb_next = (target if target is not None else self).embedding_model(nobs)
        dist = q = td = None
        if self.distribution_model is not None:
            dist = self.distribution_model.get_loss(emb, emb_next, acts, returns, gdn,
                                                    None if target is None else target.distribution_model)
        if self.q_function_model is not None:
            q = self.q_function_model.get_loss(emb, emb_next, acts, returns, gdn,
                                               None if target is None else target.q_function_model)
        if dist is not None and q is not None:
            td = dist.detach() * 0.5 + q.detach() * 0.5
        elif dist is not None:
            td = dist.detach()
        elif q is not None:
            td = q.abs().detach()
        return dist, q, td


def build_oracle_model(cfg, obs_shape, n_actions):
    """Same wiring as prism/factory/model_factory.py:49-153 for the configurations in scope."""
    if cfg.embedding_model_type == "minatar_cnn":
        emb, feat = minatar_embedding(obs_shape[-1]), 16 * 8 * 8
    elif cfg.embedding_model_type == "nature_atari_cnn":
        emb, feat = nature_embedding(cfg.frame_stack_size, cfg.use_layer_norm), 3136
    else:
        feat = cfg.embedding_model_final_dim
        emb = _Wrap(_mlp(obs_shape[-1], feat, cfg.embedding_model_num_layers, cfg.embedding_model_layer_sizes,
                         cfg.use_layer_norm, ln_first=False, out_act=True))
    iqn = ens = None
    if cfg.use_iqn:
        propagate = (cfg.ids_allow_distributional_gradients and cfg.use_ids) or not cfg.use_ids
        iqn = OracleIQN(feat, n_actions, cfg.iqn_n_basis_elements, cfg.use_layer_norm, cfg.iqn_quantile_model_layers,
                        cfg.iqn_quantile_model_feature_dim, cfg.iqn_huber_loss_kappa,
                        cfg.iqn_n_current_state_quantile_samples, cfg.iqn_n_next_state_quantile_samples,
                        cfg.iqn_quantile_samples_per_action, cfg.use_double_q_learning,
                        cfg.distributional_loss_weight, propagate)
    if cfg.use_ids:
        ens = OracleEnsemble(feat, n_actions, cfg.ids_n_q_heads, cfg.use_layer_norm, cfg.ids_n_q_head_model_layers,
                             cfg.ids_q_head_feature_dim, cfg.use_double_q_learning, cfg.q_loss_weight,
                             cfg.ids_ensemble_variation_coef)
    elif cfg.use_dqn:
        ens = OracleEnsemble(feat, n_actions, 1, cfg.use_layer_norm, cfg.dqn_n_model_layers,
                             cfg.dqn_n_model_feature_dim, cfg.use_double_q_learning, cfg.q_loss_weight, 0.0)
    return OracleComposite(emb, iqn, ens)


class OracleAgent:
    """Eager update of prism/agents/agent.py:53-79 with torch.optim.Adam (agent_factory.py:44-47)."""

    def __init__(self, cfg, obs_shape, n_actions):
        self.cfg = cfg
        self.model = build_oracle_model(cfg, obs_shape, n_actions)
        self.target = None
        if cfg.use_target_network:
            self.target = build_oracle_model(cfg, obs_shape, n_actions)
            self.target.load_state_dict(self.model.state_dict())
        self.opt = torch.optim.Adam(self.model.parameters(), lr=cfg.learning_rate,
                                    betas=(cfg.adam_beta1, cfg.adam_beta2), eps=cfg.adam_epsilon)

    def inject_taus(self, taus):
        """taus: list in draw order (current, next-online, next-target as applicable)."""
        taus = [torch.as_tensor(t, dtype=torch.float32) for t in taus]
        d = self.model.distribution_model
        if d is None:
            return
        if self.target is None:
            d.tau_queue = list(taus)
        elif self.cfg.use_double_q_learning:
            d.tau_queue = [taus[0], taus[1]]
            self.target.distribution_model.tau_queue = [taus[2]]
        else:
            d.tau_queue = [taus[0]]
            self.target.distribution_model.tau_queue = [taus[1]]

    def update(self, batch, per_weights=1):
        dist, q, td = self.model.get_losses(batch, self.target)
        total = 0
        if dist is not None:
            total = total + (dist * per_weights).mean()
        if q is not None:
            total = total + (q * per_weights).mean()
        self.opt.zero_grad()
        total.backward()
        norm = torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.cfg.max_grad_norm)
        self.opt.step()
        return {"dist": dist, "q": q, "td": td, "total": total.detach(), "grad_norm": norm}

    def sync_target(self):
        self.target.load_state_dict(self.model.state_dict())


def ids_actions(q, z, lmbda=0.1, eps=1e-10, rho_lb=0.25, return_scores=False):
    """IDSActionSelector, deterministic branch (action_selectors.py:125-167, 169-176).
    q: (N, A, K); z: (Nq, N, A).  Keeps the std-as-variance / sqrt(std)-as-std quirk."""
    mean = q.mean(dim=-1)
    variance = q.std(dim=-1)
    std = torch.sqrt(variance)
    regret = torch.max(mean + lmbda * std, dim=-1).values.view(-1, 1) - (mean - lmbda * std)
    regret_sq = regret.square()
    var_z = z.var(dim=0)
    rho = torch.clamp(var_z / (eps + var_z.mean(dim=-1, keepdim=True)), min=rho_lb)
    info_gain = torch.log(1 + variance / rho) + eps
    scores = regret_sq / info_gain
    act = torch.argmin(scores, dim=-1)
    return (act, scores) if return_scores else act


def greedy_actions(q):
    return torch.argmax(q.mean(dim=-1), dim=-1)
