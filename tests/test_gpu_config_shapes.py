"""Parity of the product update step AT THE BASELINE CONFIGURATIONS' REAL SHAPES against the CPU oracle
(oracle/agent_oracle.py, itself pinned to the reference's outputs by tests/golden/agent_*.npz).

The reference goldens are tiny (batch 5-9, feature width 16-32), so the dispatch in prism_b200/agents/ops.py routes them
to the FFMA kernel; these tests run the shapes of BASELINE.json configs[0], [1] and [4] -- where the tcgen05 3xTF32
GEMMs, the fused phi(tau) (.) x epilogue and the fused LayerNorm kernels are what actually executes -- with identical
weights, batch and injected quantile draws on both sides, and ASSERT which routes fired.

Reference update being matched: prism/agents/agent.py:53-79 over prism/agents/models/iqn_model.py:95-201,
q_ensemble.py:50-92, composite_model.py:94-144.  Bar (north star): loss / TD / clipped gradients / post-Adam parameters
within 1e-4 relative (max-norm)."""
import dataclasses

import numpy as np
import pytest
import torch

from helpers import rel_err, target_transform

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4


def _spec(name):
    import prism_b200
    if name == "configs0":      # MinAtar Breakout IDS + IQN + LayerNorm + 3-step + target, B 64, T = T' = 32, F 1024, H 256, K 10
        return prism_b200.minatar_ids_iqn_config, (10, 10, 4), 3, 64, 1
    if name == "configs1":      # MinAtar SpaceInvaders DQN + double-Q, B 256
        return prism_b200.minatar_dqn_per_config, (10, 10, 6), 4, 256, 1
    if name == "configs4":      # Atari-shaped 84x84x4, IQN 64x64 + IDS, B 512, F 3136, H 512, A 18, K 10
        return prism_b200.atari_iqn_ids_config, (4, 84, 84), 18, 512, 4
    raise KeyError(name)


def _make_batch(obs_shape, A, B, fs, seed):
    from prism_b200.experience.batch import Batch
    g = torch.Generator().manual_seed(seed)
    if len(obs_shape) == 3 and obs_shape[0] == fs and fs > 1:          # Atari: (fs, 84, 84) frames in [0, 1]
        shape = (B,) + tuple(obs_shape)
        obs = torch.randint(0, 256, shape, generator=g).float() / 255.0
        nobs = torch.randint(0, 256, shape, generator=g).float() / 255.0
    else:                                                               # MinAtar: (1, 10, 10, C) binary planes
        shape = (B, 1) + tuple(obs_shape)
        obs = (torch.rand(shape, generator=g) < 0.1).float()
        nobs = (torch.rand(shape, generator=g) < 0.1).float()
    k = torch.randint(1, 4, (B, 1), generator=g)
    batch = {"observation": obs, "next": {"observation": nobs, "reward": torch.randn(B, 1, generator=g)},
             "nonterminal": torch.rand(B, 1, generator=g) > 0.1, "gamma": (0.99 ** k.float()),
             "action": torch.randint(0, A, (B, 1), generator=g)}
    w = torch.rand(B, generator=g) * 0.7 + 0.3
    to = lambda d: Batch({k_: (to(v) if isinstance(v, dict) else v.to(DEV)) for k_, v in d.items()})
    return batch, to(batch), w


def _relu_inputs(model, run):
    """{ReLU module: (preceding Linear / Conv2d, max |z - bias|, min |z|)} over one oracle forward pass ``run()``."""
    import torch.nn as nn
    prev_of, stats, hooks = {}, {}, []
    for seq in (m for m in model.modules() if isinstance(m, nn.Sequential)):
        mods = list(seq)
        for i, m in enumerate(mods):
            if isinstance(m, nn.ReLU) and i > 0 and isinstance(mods[i - 1], (nn.Linear, nn.Conv2d)):
                prev_of[m] = mods[i - 1]
    for relu, lin in prev_of.items():
        def hook(mod, inp, _lin=lin, _relu=relu):
            z = inp[0].detach()
            b = _lin.bias.detach().view(1, -1, *([1] * (z.dim() - 2)))
            hi, lo = float((z - b).abs().max()), float(z.abs().min())
            old = stats.get(_relu)
            stats[_relu] = (_lin, max(hi, old[1]) if old else hi, min(lo, old[2]) if old else lo)
        hooks.append(relu.register_forward_pre_hook(hook))
    try:
        with torch.no_grad():
            run()
    finally:
        for h in hooks:
            h.remove()
    return stats


def _move_off_the_kinks(oracle, cpu_batch, taus, margin=3.0):
    """A ReLU mask is a step function of its pre-activation: where z is within an fp32 rounding of 0, the CPU and the
    GPU may legitimately disagree about the mask, and ONE such element moves a whole row of a weight gradient (measured
    with default initialisation: 1-3 flips per update at configs[0] shapes, rows off by up to 6e-3 -- profiles/
    diag_parity.py; the loss and the TD errors are continuous and stay exact).  Parity of GRADIENTS is therefore
    judged away from the kinks: every layer that feeds a ReLU gets biases +-margin * max|z - bias| (alternating by
    unit), so each unit is firmly on or firmly off for every row, and both mask states are exercised.  Layers are
    settled front to back (a layer's inputs change when the layer before it is moved): a few passes.  Returns the
    smallest min|z| / max|z - bias| ratio over the layers."""
    def run():
        oracle.inject_taus([t.clone() for t in taus])
        oracle.model.get_losses(cpu_batch, oracle.target)
    ratio = 0.0
    for _ in range(6):
        stats = _relu_inputs(oracle.model, run)
        ratio = min(lo / max(hi, 1e-30) for _, hi, lo in stats.values())
        if ratio > 1.0:
            break
        with torch.no_grad():
            for relu, (lin, hi, lo) in stats.items():
                sign = torch.ones_like(lin.bias)
                sign[1::2] = -1.0
                lin.bias.copy_(sign * (margin * hi))
    return ratio


def _build_pair(name, use_cuda_graph=False, dekink_with=None):
    """Product agent on the GPU, oracle agent on the CPU, identical weights (target = a transformed copy).
    ``dekink_with`` = (cpu_batch, taus): see _move_off_the_kinks."""
    import prism_b200
    from oracle.agent_oracle import OracleAgent
    make, obs_shape, A, B, fs = _spec(name)
    cfg = make(device=DEV, use_cuda_graph=use_cuda_graph)
    torch.manual_seed(123)
    agent = prism_b200.build_agent(cfg, obs_shape, A)
    sd = {k: v.detach().cpu().clone() for k, v in agent.model.state_dict().items()}
    # A freshly initialised network values all actions almost equally: the bootstrap argmax over actions
    # (iqn_model.py:129-133, q_ensemble.py:70) would then be decided by the last bits of an fp32 sum, and ONE flipped
    # row moves a gradient by ~1/B -- a discontinuity of the function, not an error of either implementation.  Spread
    # the output biases so the argmax has a margin (the trained regime), identically on both sides.
    for k, v in sd.items():
        if k.endswith(".bias") and v.dim() == 1 and v.numel() == A and ("embedding_to_quantile_layer" in k or "q_heads" in k):
            head = int(k.split("q_heads.")[1].split(".")[0]) if "q_heads." in k else 0
            v += torch.linspace(-0.6, 0.6, A).roll(head) * (1.0 + 0.05 * head)
    ocfg = dataclasses.replace(cfg, device="cpu", use_cuda_graph=False)
    oracle = OracleAgent(ocfg, obs_shape, A)
    oracle.model.load_state_dict(sd, strict=True)
    if oracle.target is not None:
        oracle.target.load_state_dict({k: target_transform(v) for k, v in sd.items()}, strict=True)
    if dekink_with is not None:
        ratio = _move_off_the_kinks(oracle, *dekink_with)
        assert ratio > 1.0, "pre-activations still near a ReLU kink: min |z| / max |z - bias| = %.3g" % ratio
        sd = {k: v.detach().clone() for k, v in oracle.model.state_dict().items()}
        if oracle.target is not None:
            oracle.target.load_state_dict({k: target_transform(v) for k, v in sd.items()}, strict=True)
    agent.model.load_state_dict(sd, strict=True)
    if agent.target_model is not None:
        agent.target_model.load_state_dict({k: target_transform(v) for k, v in sd.items()}, strict=True)
    return cfg, agent, oracle, obs_shape, A, B, fs


def _taus(cfg, B, seed):
    if not cfg.use_iqn:
        return []
    g = torch.Generator().manual_seed(seed)
    T, Tp = cfg.iqn_n_current_state_quantile_samples, cfg.iqn_n_next_state_quantile_samples
    n_next = 2 if (cfg.use_target_network and cfg.use_double_q_learning) else 1
    return [torch.rand(T * B, 1, generator=g)] + [torch.rand(Tp * B, 1, generator=g) for _ in range(n_next)]


def _inject(cfg, agent, taus):
    from test_gpu_agent import inject_taus
    inject_taus(cfg, agent, [t.clone() for t in taus])


def _named_grads(agent):
    from test_gpu_agent import reference_named_grads
    return reference_named_grads(agent)


def _compare(agent, oracle, out, td, name, tol=TOL):
    errs = {}
    if out["dist"] is not None:
        errs["loss.dist"] = rel_err(agent._static_distribution_loss.detach().cpu().numpy(), out["dist"].detach().numpy())
    if out["q"] is not None:
        errs["loss.q"] = rel_err(agent._static_q_loss.detach().cpu().numpy(), out["q"].detach().numpy())
    errs["td"] = rel_err(td.cpu().numpy(), out["td"].numpy())
    errs["loss.total"] = rel_err(agent._static_total_loss.detach().cpu().numpy(), out["total"].numpy())
    # clip coefficient applied by the fused clip+Adam kernel vs clip_grad_norm_ (the oracle's .grad are already clipped)
    ograds = {k: p.grad for k, p in oracle.model.named_parameters()}
    coef = float(agent.optimizer.norm_out[1])
    for k, g in _named_grads(agent).items():
        errs["grad." + k] = rel_err(g.cpu().numpy() * coef, ograds[k].numpy())
    osd = oracle.model.state_dict()
    for k, v in agent.model.state_dict().items():
        errs["param." + k] = rel_err(v.cpu().numpy(), osd[k].numpy())
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    td_rows = int((np.abs(td.cpu().numpy() - out["td"].numpy()) > 1e-4 * np.abs(out["td"].numpy()).max()).sum())
    print("%s: worst errors %s; td rows off: %d of %d" % (name, ", ".join("%s %.2e" % kv for kv in worst), td_rows, td.numel()))
    bad = [(k, e) for k, e in errs.items() if not e < tol]
    assert not bad, "%s: %d quantities beyond %.0e, worst %s" % (name, len(bad), tol, worst[:3])
    return max(e for k, e in errs.items() if k.startswith("grad."))


@pytest.mark.parametrize("name", ["configs0", "configs1", "configs4"])
def test_update_at_baseline_shapes_matches_oracle(name):
    from prism_b200.agents import ops
    make, obs_shape, A, B, fs = _spec(name)
    cpu_batch, dev_batch, w = _make_batch(obs_shape, A, B, fs, seed=7)
    taus = _taus(make(), B, seed=11)
    cfg, agent, oracle, obs_shape, A, B, fs = _build_pair(name, dekink_with=(cpu_batch, taus))
    oracle.inject_taus([t.clone() for t in taus])
    out = oracle.update(cpu_batch, w)
    _inject(cfg, agent, taus)
    ops.route_counts(reset=True)
    td = agent.update(dev_batch, w.to(DEV))
    torch.cuda.synchronize()
    routes = ops.route_counts()
    worst = _compare(agent, oracle, out, td, name)
    print("%s: worst clipped-gradient error %.2e; routes %s" % (name, worst, routes))
    # the fast kernels must be what ran: this is the point of testing at these shapes
    assert ops.fallthrough_count() == 0, "library fall-throughs on the %s path: %s" % (name, routes)
    if name in ("configs0", "configs4"):
        assert routes.get("linear:tc_gemm", 0) >= 2, routes          # IQN hidden layer(s) + K-head ensemble on tcgen05
        assert routes.get("phi_x:tc_gemm", 0) >= 2, routes           # cos-embedding GEMM with the fused phi (.) x epilogue
        assert routes.get("ln:fused", 0) >= 2, routes                # LayerNorm kernels of this library


@pytest.mark.parametrize("name", ["configs0", "configs4"])
def test_default_initialisation_values_exact_gradients_up_to_relu_flips(name):
    """The same comparison WITHOUT moving the weights off the ReLU kinks (the reference's own initialisation): losses and
    TD errors -- continuous in every pre-activation -- must still agree to 1e-4; a gradient tensor may differ only
    through a handful of flipped masks, i.e. in a small part of its energy (relative L2 error below 2 %)."""
    cfg, agent, oracle, obs_shape, A, B, fs = _build_pair(name)
    cpu_batch, dev_batch, w = _make_batch(obs_shape, A, B, fs, seed=40)
    taus = _taus(cfg, B, seed=50)
    oracle.inject_taus([t.clone() for t in taus])
    out = oracle.update(cpu_batch, w)
    _inject(cfg, agent, taus)
    td = agent.update(dev_batch, w.to(DEV))
    torch.cuda.synchronize()
    assert rel_err(td.cpu().numpy(), out["td"].numpy()) < TOL
    assert rel_err(agent._static_total_loss.detach().cpu().numpy(), out["total"].numpy()) < TOL
    ograds = {k: p.grad for k, p in oracle.model.named_parameters()}
    coef = float(agent.optimizer.norm_out[1])
    worst = 0.0
    for k, g in _named_grads(agent).items():
        a, b = g.cpu().double().numpy() * coef, ograds[k].double().numpy()
        worst = max(worst, float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)))
    print("%s default init: worst relative L2 gradient error %.2e" % (name, worst))
    assert worst < 2e-2


@pytest.mark.parametrize("name", ["configs0", "configs1"])
def test_graph_update_at_baseline_shapes_matches_oracle(name):
    """Same comparison with the update captured as a CUDA graph (Agent._update_with_cuda_graph): two steps, so the
    replayed graph -- not the capture's warm-up -- is what is checked."""
    make, obs_shape, A, B, fs = _spec(name)
    probe = (_make_batch(obs_shape, A, B, fs, seed=20)[0], _taus(make(), B, seed=30))
    cfg, agent, oracle, obs_shape, A, B, fs = _build_pair(name, use_cuda_graph=True, dekink_with=probe)
    for it in range(2):
        cpu_batch, dev_batch, w = _make_batch(obs_shape, A, B, fs, seed=20 + it)
        taus = _taus(cfg, B, seed=30 + it)
        oracle.inject_taus([t.clone() for t in taus])
        out = oracle.update(cpu_batch, w)
        if cfg.use_iqn:
            # a captured graph draws its quantiles on the device: feed the same draws through static tensors
            _set_static_taus(cfg, agent, taus)
        td = agent.update(dev_batch, w.to(DEV))
        torch.cuda.synchronize()
        _compare(agent, oracle, out, td, "%s step %d" % (name, it))


def _set_static_taus(cfg, agent, taus):
    """Captured graphs cannot pop a Python queue per replay: IQNModel.static_taus holds device tensors that every
    forward of that module reads, in the reference's draw order (current, next-online | next-target)."""
    d = agent.model.distribution_model
    dev = [t.to(DEV).float().contiguous() for t in taus]
    if agent.target_model is None:
        d.set_static_taus(dev)
    elif cfg.use_double_q_learning:
        d.set_static_taus(dev[:2])
        agent.target_model.distribution_model.set_static_taus(dev[2:])
    else:
        d.set_static_taus(dev[:1])
        agent.target_model.distribution_model.set_static_taus(dev[1:])


def test_configs0_parity_holds_over_consecutive_steps():
    """Three consecutive eager product updates vs three oracle updates at configs[0] shapes: the Adam state carries
    over, so a systematic difference between the two implementations would compound here.  Parameters are compared by
    their UPDATE (value - initial value): Adam's step is ~lr per element whatever the gradient's size, so the change of
    a parameter, not its magnitude, is the scale on which an fp32 rounding difference of a near-zero gradient shows."""
    make, obs_shape, A, B, fs = _spec("configs0")
    probe = (_make_batch(obs_shape, A, B, fs, seed=40)[0], _taus(make(), B, seed=50))
    cfg, agent, oracle, obs_shape, A, B, fs = _build_pair("configs0", dekink_with=probe)
    init = {k: v.detach().cpu().clone() for k, v in agent.model.state_dict().items()}
    for it in range(3):
        cpu_batch, dev_batch, w = _make_batch(obs_shape, A, B, fs, seed=40 + it)
        taus = _taus(cfg, B, seed=50 + it)
        oracle.inject_taus([t.clone() for t in taus])
        out = oracle.update(cpu_batch, w)
        _inject(cfg, agent, taus)
        td = agent.update(dev_batch, w.to(DEV))
        torch.cuda.synchronize()
        _compare(agent, oracle, out, td, "configs0 step %d" % it, tol=2e-4)
    osd = oracle.model.state_dict()
    frac_off = []
    for k, v in agent.model.state_dict().items():
        du, do = (v.cpu() - init[k]).numpy(), (osd[k] - init[k]).numpy()
        # elements whose 3-step update differs by more than 2 % of the largest update of the tensor
        frac_off.append(float((np.abs(du - do) > 0.02 * np.abs(do).max()).mean()))
    assert max(frac_off) < 1e-3, max(frac_off)


# ------------------------------------------------------------------------------------------------
# priority store at BASELINE sizes against the C oracle (oracle/per_oracle.c)
# ------------------------------------------------------------------------------------------------
def test_tree_2_24_batch_4096_is_bit_exact_vs_oracle():
    """configs[2]: 2^24 leaves, batch 4096 -- sampled indices, masses, whole sum / min trees and the state block after
    sample -> update rounds, against the oracle (128 MiB per tree on the host)."""
    from oracle.per_oracle import OracleTree
    from prism_b200 import PrioritizedTree
    N, B = 1 << 24, 4096
    rng = np.random.default_rng(24)
    leaves = np.sqrt(rng.exponential(1.0, N).astype(np.float32) + np.float32(1e-8))
    g, o = PrioritizedTree(N, device=DEV, mode="stratified"), OracleTree(N)
    g.build(torch.from_numpy(leaves).to(DEV))
    o.build(leaves)
    for it in range(3):
        u = rng.random(B)
        mass = torch.empty(B, dtype=torch.float32, device=DEV)
        idx, w = g.sample(B, u=torch.from_numpy(u).to(DEV), mass_out=mass)
        oi, ow, om, _, _ = o.sample(u, 0.5, mode=1)
        assert np.array_equal(idx.cpu().numpy(), oi), "round %d: sampled indices" % it
        assert np.array_equal(mass.cpu().numpy(), om)
        assert np.allclose(w.cpu().numpy(), ow, rtol=1e-6, atol=0)
        prio = rng.exponential(2.0, B).astype(np.float32)
        g.update_priority(idx, torch.from_numpy(prio).to(DEV), sorted=True)
        o.update_priority(oi, prio)
    # one iid round (torchrl's mode): unsorted indices with duplicates through the general update path
    u = rng.random(B)
    idx, w = g.sample(B, u=torch.from_numpy(u).to(DEV), mode="iid")
    oi, ow, _, _, _ = o.sample(u, 0.5, mode=0)
    assert np.array_equal(idx.cpu().numpy(), oi)
    prio = rng.exponential(2.0, B).astype(np.float32)
    g.update_priority(idx, torch.from_numpy(prio).to(DEV), sorted=False)
    o.update_priority(oi, prio)
    torch.cuda.synchronize()
    assert np.array_equal(g.sum.cpu().numpy()[1:], o.sum[1:]), "sum tree differs"
    assert np.array_equal(g.min.cpu().numpy()[1:], o.min[1:]), "min tree differs"
    st = g.state_host()
    assert st["max_priority"] == np.float32(o.max_priority) and st["p_sum"] == o.sum[1] and st["p_min"] == o.min[1]


@pytest.mark.parametrize("K", [4, 16, 64])
def test_tree_2_24_batches_in_flight_are_bit_exact_vs_oracle(K):
    """configs[2] with K batches of 4096 in flight (one sampling launch, one write-back call) against K sequential
    oracle samples + K sequential oracle updates: indices, whole trees, state."""
    from oracle.per_oracle import OracleTree
    from prism_b200 import PrioritizedTree
    N, B = 1 << 24, 4096
    rng = np.random.default_rng(240 + K)
    leaves = np.sqrt(rng.exponential(1.0, N).astype(np.float32) + np.float32(1e-8))
    g, o = PrioritizedTree(N, device=DEV, mode="stratified"), OracleTree(N)
    g.build(torch.from_numpy(leaves).to(DEV))
    o.build(leaves)
    for rnd in range(2):
        u = rng.random((K, B))
        idx, w = g.sample(B, u=torch.from_numpy(u.reshape(-1)).to(DEV), n_batches=K)
        gi = idx.cpu().numpy().reshape(K, B)
        newp = rng.exponential(2.0, (K, B)).astype(np.float32)
        ois = []
        for b in range(K):
            oi, ow, _, _, _ = o.sample(u[b], 0.5, mode=1)
            assert np.array_equal(gi[b], oi), (rnd, b)
            ois.append(oi)
        g.update_priority(idx, torch.from_numpy(newp.reshape(-1)).to(DEV), sorted=False)
        for b in range(K):
            o.update_priority(ois[b], newp[b])
    torch.cuda.synchronize()
    gs, gm = g.export()
    assert np.array_equal(gs.cpu().numpy()[1:], o.sum[1:]), "sum tree differs"
    assert np.array_equal(gm.cpu().numpy()[1:], o.min[1:]), "min tree differs"
    st = g.state_host()
    assert st["max_priority"] == np.float32(o.max_priority) and st["p_sum"] == o.sum[1] and st["p_min"] == o.min[1]
    assert bool(torch.isfinite(g.leaves()).all()) and int(g.counters.abs().sum()) == 0


@pytest.mark.parametrize("G", [2, 4, 8])
def test_sharded_2_23_per_gpu_equals_one_big_tree(G):
    """configs[3]: G emulated shards of 2^23 leaves on one device against ONE oracle tree of G * 2^23 leaves: global
    stratified sampling (batch 4096) bit-exact, then every shard's priority update, then a second sample."""
    from oracle.per_oracle import OracleTree
    from prism_b200 import PrioritizedTree
    C, B = 1 << 23, 4096
    free, _ = torch.cuda.mem_get_info()
    if free < G * C * 4 * 6:
        pytest.skip("not enough device memory for %d shards" % G)
    rng = np.random.default_rng(100 + G)
    leaves = np.sqrt(rng.exponential(1.0, G * C).astype(np.float32) + np.float32(1e-8))
    big = OracleTree(G * C)
    big.build(leaves)
    shards = []
    for r in range(G):
        t = PrioritizedTree(C, device=DEV, mode="stratified")
        t.build(torch.from_numpy(leaves[r * C:(r + 1) * C]).to(DEV))
        shards.append(t)
    for it in range(2):
        torch.cuda.synchronize()
        all_state = torch.stack([t.state for t in shards]).contiguous()
        u = rng.random(B)
        oi, ow, _, _, _ = big.sample(u, 0.5, mode=1)
        ud = torch.from_numpy(u).to(DEV)
        got_idx, got_w = np.full(B, -1, np.int64), np.zeros(B, np.float32)
        prio = rng.exponential(2.0, B).astype(np.float32)
        for r in range(G):
            strat = torch.empty(B, dtype=torch.int64, device=DEV)
            idx, w = shards[r].sample_global(G, r, all_state, B, ud, stratum_out=strat)
            st = shards[r].state_host()
            n, lo = st["owned_n"], st["owned_lo"]
            k = strat[:n].cpu().numpy()
            assert np.array_equal(k, np.arange(lo, lo + n))
            got_idx[k] = idx[:n].cpu().numpy() + r * C
            got_w[k] = w[:n].cpu().numpy()
            # owner-computes write-back: the shard updates exactly the strata it drew
            shards[r].update_priority(idx[:n], torch.from_numpy(prio[lo:lo + n]).to(DEV), sorted=True)
        assert np.array_equal(got_idx, oi), "G=%d round %d" % (G, it)
        assert np.allclose(got_w, ow, rtol=1e-6)
        big.update_priority(oi, prio)
    torch.cuda.synchronize()
    for r in range(G):                                  # every shard's leaves == its slice of the big tree
        leaf = shards[r].leaves().cpu().numpy()
        assert np.array_equal(leaf, big.sum[big.capacity + r * C: big.capacity + (r + 1) * C]), "shard %d" % r
