"""Shared helpers of the parity tests: build (CUDA agent, oracle agent) pairs on identical weights and data, step
them with identically seeded global RNGs, compare info dicts and parameters."""
from __future__ import annotations

import numpy as np
import torch

from oracle import rl_oracle as O


class Space:
    def __init__(self, A):
        self.low, self.high, self.shape = -np.ones(A, np.float32), np.ones(A, np.float32), (A,)


def make_pair(alg, S, A, kw, rows, precision="tf32", use_cuda_graph=True, seed=0, oracle_kw=None, agent_kw=None):
    from rlrep_b200 import ReplayBuffer
    from rlrep_b200.agents import AGENTS
    init = O.init_state(alg, S, A, kw, seed=seed)
    oracle_kw = dict(oracle_kw or {})
    extra_state = {}
    if alg == "vlsac":  # the critic's fixed noise is construction-time state on both sides
        g = torch.Generator().manual_seed(1234)
        noise = torch.randn(20, kw.get("feature_dim", 256), generator=g)
        oracle_kw["critic_noise"] = noise
        extra_state["critic.noise"] = noise
    oracle = O.ORACLES[alg](S, A, init, discount=0.99, tau=0.005, **kw, **oracle_kw)
    oring = O.synthetic_ring(S, A, rows, seed=0)
    agent = AGENTS[alg](state_dim=S, action_dim=A, action_space=Space(A), discount=0.99, tau=0.005,
                        precision=precision, use_cuda_graph=use_cuda_graph, **kw, **(agent_kw or {}))
    agent.load_state_dict({**init, **extra_state})
    buf = ReplayBuffer(S, A, max_size=rows)
    buf.load(oring.state, oring.action, oring.next_state, oring.reward, oring.done)
    return agent, buf, oracle, oring


def step_both(agent, buf, oracle, oring, B, n, seed=1):
    """n train() calls on each side under the same global seeds -> (infos_cuda, infos_oracle)."""
    np.random.seed(seed)
    torch.manual_seed(seed)
    oi = [oracle.train(oring, B) for _ in range(n)]
    np.random.seed(seed)
    torch.manual_seed(seed)
    ci = [agent.train(buf, B) for _ in range(n)]
    return ci, oi


def worst_info_error(ci, oi, atol=1e-6):
    worst, where = 0.0, None
    for step, (c, o) in enumerate(zip(ci, oi)):
        assert set(c) == set(o), (sorted(c), sorted(o))
        for k in o:
            err = max(0.0, abs(c[k] - o[k]) - atol) / (abs(o[k]) + 1e-12)
            if err > worst:
                worst, where = err, (step, k, c[k], o[k])
    return worst, where


def worst_param_error(agent, oracle):
    """Per-tensor relative L2 distance (norm-wise: Adam turns 1-ulp gradient noise into +-2 lr on single elements,
    SURVEY.md 7.2 #1) -> (worst value, tensor name, fraction of elements beyond 1e-3 relative)."""
    csd, osd = agent.state_dict(), oracle.state_dict()
    worst, where = 0.0, None
    outliers = 0
    total = 0
    for k, v in osd.items():
        if k == "log_alpha":
            assert abs(float(csd[k]) - float(v)) < 1e-4 * max(1.0, abs(float(v))), (float(csd[k]), float(v))
            continue
        assert k in csd, f"{k} missing from the CUDA agent's state_dict"
        c = csd[k].double().reshape(-1)
        v = v.double().reshape(-1)
        d = (c - v).norm().item() / (v.norm().item() + 1e-30)
        outliers += int(((c - v).abs() > 1e-3 * v.abs() + 1e-6).sum())
        total += v.numel()
        if d > worst:
            worst, where = d, k
    return worst, where, outliers / max(total, 1)


def worst_param_error_conditioned(agent, oracle, before, g_floor=1e-3):
    """Per-tensor relative L2 distance after ONE train() call, leaving out the elements whose Adam step is ill-conditioned:
    on the first step m / (sqrt(v) + eps) = g / (|g| + eps') is +-1 whatever |g| is, so where the oracle's gradient is below
    `g_floor` of the tensor's RMS gradient the SIGN of a +-lr move is decided by rounding.  The gradients are the ones the
    oracle recorded at each tensor's most recent optimiser step (`last_step_grads`)  -> (worst, name, fraction left out)."""
    csd, osd = agent.state_dict(), oracle.state_dict()
    worst, where, skipped, total = 0.0, None, 0, 0
    for k, v in osd.items():
        if k == "log_alpha":
            assert abs(float(csd[k]) - float(v)) < 1e-5 * max(1.0, abs(float(v))), (float(csd[k]), float(v))
            continue
        c, v, b = csd[k].double().reshape(-1), v.double().reshape(-1), before[k].double().reshape(-1)
        g = getattr(oracle, "last_step_grads", {}).get(k)  # gradient at the tensor's most recent optimiser step
        g = g.double().reshape(-1) if g is not None and g.numel() == v.numel() else None
        keep = torch.ones_like(v, dtype=torch.bool)
        if g is not None and (v - b).abs().max() > 0:
            rms = g.pow(2).mean().sqrt()
            keep = g.abs() >= g_floor * rms
        skipped += int((~keep).sum())
        total += v.numel()
        d = ((c - v)[keep].norm() / (v.norm() + 1e-30)).item()
        if d > worst:
            worst, where = d, k
    return worst, where, skipped / max(total, 1)
