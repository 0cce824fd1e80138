#!/usr/bin/env python
"""check_torchrl.py -- pin oracle/per_oracle.c against a REAL torchrl, the day one is importable.

TEST INFRASTRUCTURE.  The sum-tree / sampler arithmetic of the reference lives in torchrl
(prism/factory/exp_buffer_factory.py:3-4,22-28; call sites prism/experience/timestep_buffer.py:33,37,54), which is
absent from /root/reference, from this image and from the wheelhouse, and the reference pins no version.  The oracle
restates torchrl's published algorithm with every version-dependent choice behind a flag:

    strict_pow2                tree capacity = first power of two STRICTLY greater than N (value-equivalent)
    weight_eps_in_denominator  w = (p / (p_min + eps)) ** -beta   instead of (p / p_min) ** -beta
    default_priority_fp64      (max_priority + eps) ** alpha evaluated in float64 (Python scalars) vs float32

This script replays the inputs of tests/golden/per_tree.npz (priorities, uniforms, new priorities) through
``torchrl.data.PrioritizedReplayBuffer`` and reports, for every combination of the flags, whether the oracle agrees
with torchrl on sampled indices (bit-exact), importance weights (1e-6) and max_priority.  Without torchrl it prints
"torchrl absent -- parity unpinned" and exits 0 (``__graft_entry__.smoke()`` prints the same line, non-fatally).

    python oracle/check_torchrl.py            # probes sys.path and baseline/_ref
"""
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def find_torchrl():
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref) and ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        import tensordict  # noqa: F401
        import torchrl
        from torchrl.data import ListStorage, PrioritizedReplayBuffer  # noqa: F401
        return torchrl
    except Exception as e:                                    # ImportError, or a wheel built for another torch
        return e


def replay_torchrl(fx):
    """The golden's script through the real sampler.  torchrl draws its own masses with numpy
    (``np.random.uniform(0, p_sum, batch)``): the uniforms are injected by patching that one call."""
    import torch
    from torchrl.data import ListStorage, PrioritizedReplayBuffer
    N = int(fx["N"])
    rb = PrioritizedReplayBuffer(storage=ListStorage(max_size=N), collate_fn=lambda x: x, alpha=0.5, beta=0.5,
                                 batch_size=64)
    rb.extend(list(range(N)))
    rb.update_priority(torch.arange(N), torch.from_numpy(fx["p0"]))
    out = []
    real_uniform = np.random.uniform
    for r in range(4):
        if r % 2 == 1:
            continue                                           # stratified rounds are ours, not torchrl's
        u = fx["r%d.u" % r]
        np.random.uniform = lambda lo, hi, size=None, _u=u: lo + (hi - lo) * _u
        try:
            _, info = rb.sample(return_info=True)
        finally:
            np.random.uniform = real_uniform
        idx = torch.as_tensor(info["index"]).numpy().astype(np.int64)
        w = torch.as_tensor(info["_weight"]).numpy().astype(np.float32)
        rb.update_priority(torch.from_numpy(idx), torch.from_numpy(fx["r%d.newp" % r]))
        out.append((r, idx, w, float(rb._sampler._max_priority if not hasattr(rb._sampler._max_priority, "__len__")
                                     else rb._sampler._max_priority[0])))
    return out


def replay_oracle(fx, **flags):
    from oracle.per_oracle import OracleTree
    N = int(fx["N"])
    t = OracleTree(N, **flags)
    t.extend(N)
    t.update_priority(np.arange(N), fx["p0"])
    out = []
    for r in range(4):
        if r % 2 == 1:
            continue
        idx, w, _, _, _ = t.sample(fx["r%d.u" % r], 0.5, mode=0)
        t.update_priority(idx, fx["r%d.newp" % r])
        out.append((r, idx, w, float(t.max_priority)))
    return out


def main(verbose=True):
    found = find_torchrl()
    if isinstance(found, Exception):
        msg = "torchrl absent -- parity unpinned (%s: %s)" % (type(found).__name__, str(found).splitlines()[0][:80])
        if verbose:
            print(msg)
        return {"torchrl": None, "message": msg}
    fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "per_tree.npz"), allow_pickle=False))
    try:
        real = replay_torchrl(fx)
    except Exception as e:
        msg = "torchrl %s imports but the replay failed (%r) -- parity unpinned" % (getattr(found, "__version__", "?"), e)
        if verbose:
            print(msg)
        return {"torchrl": getattr(found, "__version__", "?"), "message": msg}
    report = {"torchrl": getattr(found, "__version__", "?"), "flags": []}
    for sp, we, dp in itertools.product((False, True), repeat=3):
        flags = dict(strict_pow2=sp, weight_eps_in_denominator=we, default_priority_fp64=dp)
        mine = replay_oracle(fx, **flags)
        ok_idx = all(np.array_equal(a[1], b[1]) for a, b in zip(real, mine))
        ok_w = all(np.allclose(a[2], b[2], rtol=1e-6, atol=0) for a, b in zip(real, mine))
        ok_mp = all(np.float32(a[3]) == np.float32(b[3]) for a, b in zip(real, mine))
        report["flags"].append({**flags, "indices": ok_idx, "weights": ok_w, "max_priority": ok_mp})
        if verbose:
            print("strict_pow2=%d weight_eps=%d default_fp64=%d : indices %s  weights %s  max_priority %s"
                  % (sp, we, dp, ok_idx, ok_w, ok_mp))
    full = [f for f in report["flags"] if f["indices"] and f["weights"] and f["max_priority"]]
    report["message"] = ("torchrl %s: oracle PINNED with flags %s" % (report["torchrl"], {k: full[0][k] for k in
                         ("strict_pow2", "weight_eps_in_denominator", "default_priority_fp64")})) if full else \
        "torchrl %s: NO flag combination reproduces it -- restatement needs fixing" % report["torchrl"]
    if verbose:
        print(report["message"])
    return report


if __name__ == "__main__":
    main()
