"""Synthetic inputs of the benchmarks and parity tests -- DATA ONLY, no update arithmetic.

Shared by `bench.py` (both arms), the parity tests and the CPU oracle, so that every side is fed the same replay rows and
the same initial weights: a host ring with the reference's fp64 column layout (utils/buffer.py:13-48), the synthetic
transition generator, the per-agent layer tables (names = the reference's state_dict names) and deterministic initial
weights drawn from a private generator.
"""
from __future__ import annotations

import collections
import math

import numpy as np
import torch

Batch = collections.namedtuple("Batch", ["state", "action", "reward", "next_state", "done"])  # utils/buffer.py:7-10


# --------------------------------------------------------------------------------------------------
# Replay ring (utils/buffer.py:13-48): fp64 host arrays, uniform-with-replacement index draw from the
# global legacy numpy RNG, fp32 cast at sample time.
# --------------------------------------------------------------------------------------------------
class HostRing:
    def __init__(self, state_dim, action_dim, max_size=int(1e6)):
        self.max_size, self.ptr, self.size = max_size, 0, 0
        self.state = np.zeros((max_size, state_dim))
        self.action = np.zeros((max_size, action_dim))
        self.next_state = np.zeros((max_size, state_dim))
        self.reward = np.zeros((max_size, 1))
        self.done = np.zeros((max_size, 1))

    def add(self, state, action, next_state, reward, done):  # buffer.py:28-36
        i = self.ptr
        self.state[i], self.action[i], self.next_state[i] = state, action, next_state
        self.reward[i], self.done[i] = reward, done
        self.ptr = (i + 1) % self.max_size
        self.size = min(self.size + 1, self.max_size)

    def take(self, ind) -> Batch:  # buffer.py:42-48 (the fp64 -> fp32 cast happens here)
        f = lambda a: torch.from_numpy(a[ind]).float()
        return Batch(state=f(self.state), action=f(self.action), reward=f(self.reward),
                     next_state=f(self.next_state), done=f(self.done))

    def sample(self, batch_size) -> Batch:  # buffer.py:39-40
        return self.take(np.random.randint(0, self.size, size=batch_size))


def synthetic_ring(state_dim, action_dim, n_rows, seed=0, ring_cls=HostRing):
    """Synthetic replay data of SURVEY.md 8d: s, s' ~ N(0,1); a ~ U(-1,1); r ~ N(0,1); done ~ Bernoulli(1e-3)."""
    rng = np.random.default_rng(seed)
    ring = ring_cls(state_dim, action_dim, max_size=n_rows)
    ring.state[:] = rng.standard_normal((n_rows, state_dim))
    ring.action[:] = rng.uniform(-1.0, 1.0, (n_rows, action_dim))
    ring.next_state[:] = rng.standard_normal((n_rows, state_dim))
    ring.reward[:] = rng.standard_normal((n_rows, 1))
    ring.done[:] = (rng.random((n_rows, 1)) < 1e-3).astype(np.float64)
    ring.size, ring.ptr = n_rows, 0
    return ring


# --------------------------------------------------------------------------------------------------
# Parameter tables.  Names are the reference's state_dict names, order is the reference's optimizer order.
# --------------------------------------------------------------------------------------------------
def _mlp_names(prefix, dims):
    """util.mlp (utils/util.py:85-96): Linear at Sequential indices 0, 2, 4, ..."""
    return [(f"{prefix}.{2 * i}", dims[i + 1], dims[i]) for i in range(len(dims) - 1)]


def layer_table(alg: str, S: int, A: int, cfg: dict):
    """[(module, [(layer_name, out, in), ...])] for one agent, in reference construction/optimizer order."""
    H, D = cfg.get("hidden_dim", 256), cfg.get("feature_dim", 256)
    actor_h = {"sac": H, "ctrlsac": 256, "vlsac": H, "diffsrsac": H,
               "spedersac": cfg.get("critic_and_actor_hidden_dim", 256)}[alg]
    actor = ("actor", _mlp_names("actor.trunk", [S, actor_h, actor_h, 2 * A]))  # actor.py:66-74
    if alg == "sac":  # critic.py:15-24
        crit = ("critic", _mlp_names("critic.Q1", [S + A, H, H, 1]) + _mlp_names("critic.Q2", [S + A, H, H, 1]))
        return [crit, actor]
    if alg == "ctrlsac":  # ctrlsac_agent.py:18-120
        return [
            ("phi", [("phi.l1", H, S + A), ("phi.l2", H, H), ("phi.l3", D, H)]),
            ("mu", [("mu.l1", H, S), ("mu.l2", H, H), ("mu.l3", D, H)]),
            ("theta", [("theta.l", 1, D)]),
            actor,
            ("critic", [("critic.l1", H, D), ("critic.l2", 1, H), ("critic.l4", H, D), ("critic.l5", 1, H)]),
        ]
    if alg == "vlsac":  # networks/vae.py:13-120, vlsac_agent.py:17-41
        return [
            ("encoder", [("encoder.l1", 256, 2 * S + A), ("encoder.l2", 256, 256), ("encoder.mean_linear", D, 256),
                         ("encoder.log_std_linear", D, 256)]),
            ("decoder", [("decoder.l1", 256, D), ("decoder.state_linear", S, 256), ("decoder.reward_linear", 1, 256)]),
            ("f", [("f.l1", 256, S + A), ("f.l2", 256, 256), ("f.mean_linear", D, 256), ("f.log_std_linear", D, 256)]),
            actor,
            ("critic", [("critic.l1", H, D), ("critic.l2", H, H), ("critic.l3", 1, H), ("critic.l4", H, D),
                        ("critic.l5", H, H), ("critic.l6", 1, H)]),
        ]
    if alg == "spedersac":  # spedersac_agent.py:21-98,147-158
        ph, pd = cfg["phi_hidden_dim"], cfg["phi_hidden_depth"]
        mh, md = cfg["mu_hidden_dim"], cfg["mu_hidden_depth"]
        CH = cfg["critic_and_actor_hidden_dim"]
        return [
            ("phi", _mlp_names("phi.trunk", [S + A] + [ph] * pd + [D])),
            ("mu", _mlp_names("mu.trunk", [S] + [mh] * md + [D])),
            ("theta", [("theta.l", 1, D)]),
            actor,
            ("critic", [("critic.l1", CH, D), ("critic.l2", CH, CH), ("critic.l3", 1, CH), ("critic.l4", CH, D),
                        ("critic.l5", CH, CH), ("critic.l6", 1, CH)]),
        ]
    if alg == "diffsrsac":  # diffsrsac_agent.py:14-60
        ph, pd = cfg.get("phi_hidden_dim", 256), cfg.get("phi_hidden_depth", 1)
        nh, nd = cfg.get("nabla_mu_hidden_dim", 512), cfg.get("nabla_mu_hidden_depth", 1)
        return [
            ("phi", _mlp_names("critic_feed_feature.z_vector", [S + A] + [ph] * pd + [D])),
            ("nablamu", _mlp_names("nablamu_net.Mu_z_by_s_layer", [S + 1] + [nh] * nd + [D * S])),
            actor,
            ("critic", [("critic.l1", H, D), ("critic.l2", H, H), ("critic.l3", 1, H), ("critic.l4", H, D),
                        ("critic.l5", H, H), ("critic.l6", 1, H)]),
        ]
    raise ValueError(alg)


def init_state(alg: str, S: int, A: int, cfg: dict, seed: int = 0) -> "collections.OrderedDict[str, torch.Tensor]":
    """Deterministic initial weights from a private generator (NOT the reference's init order): U(+-1/sqrt(fan_in))
    like nn.Linear's default.  Parity tests load the same table into the reference, the oracle and the CUDA
    agent, so only determinism matters here."""
    g = torch.Generator().manual_seed(seed)
    sd = collections.OrderedDict()
    for _, layers in layer_table(alg, S, A, cfg):
        for name, out, inp in layers:
            bound = 1.0 / math.sqrt(inp)
            sd[name + ".weight"] = (torch.rand(out, inp, generator=g) * 2 - 1) * bound
            sd[name + ".bias"] = (torch.rand(out, generator=g) * 2 - 1) * bound
    return sd
