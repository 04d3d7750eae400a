"""Generates tests/golden/*.npz by running the REAL reference agents (imported from /root/reference).

Only runs in the build container (the reference tree is not shipped to the GPU box); the fixtures it writes
are committed and are what pins oracle/rl_oracle.py:

    python -m oracle.make_golden            # regenerate every fixture and check the oracle against each

For every case: deterministic initial weights (rl_oracle.init_state) are loaded into the reference agent,
the reference replay buffer is filled with the synthetic rows of SURVEY.md 8d, the global RNGs are seeded
(np.random.seed(1); torch.manual_seed(1)) and `agent.train(buffer, B)` is called n times.  Recorded: the info
dict of every call, and for every parameter / target tensor afterwards [sum, l2, absmax] plus a strided
sample of its elements.
"""
from __future__ import annotations

import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parents[1] / "tests" / "golden"

HC = dict(S=17, A=6)
HUM = dict(S=376, A=17)

SPEDER_MAIN = dict(extra_feature_steps=5, phi_and_mu_lr=0.00001, phi_hidden_dim=512, phi_hidden_depth=1,
                   mu_hidden_dim=512, mu_hidden_depth=0, critic_and_actor_lr=0.0003, critic_and_actor_hidden_dim=256)

# name -> (alg, shapes, ctor kwargs beyond main.py's {discount, tau, hidden_dim}, batch, ring rows, n train() calls)
CASES = {
    "sac_hc_b256": ("sac", HC, dict(hidden_dim=256), 256, 5000, 4),
    "ctrlsac_small": ("ctrlsac", HC, dict(hidden_dim=64, feature_dim=128, extra_feature_steps=3), 32, 2000, 4),
    "ctrlsac_hc_b256": ("ctrlsac", HC, dict(hidden_dim=1024, feature_dim=2048, extra_feature_steps=3), 256, 5000, 2),
    "vlsac_hc_b64": ("vlsac", HC, dict(hidden_dim=256, feature_dim=256, extra_feature_steps=3), 64, 2000, 4),
    "vlsac_hum_b128": ("vlsac", HUM, dict(hidden_dim=256, feature_dim=256, extra_feature_steps=3), 128, 2000, 2),
    "spedersac_hc_b64": ("spedersac", HC, dict(hidden_dim=256, feature_dim=256, **SPEDER_MAIN), 64, 2000, 4),
    "diffsrsac_hc_b64": ("diffsrsac", HC, dict(hidden_dim=256), 64, 2000, 4),
}


def import_reference():
    """Shims of SURVEY.md 8c: `gym` (import only) and `torchinfo.summary` (printing only)."""
    if "gym" not in sys.modules:
        sys.modules["gym"] = types.ModuleType("gym")
    if "torchinfo" not in sys.modules:
        ti = types.ModuleType("torchinfo")
        ti.summary = lambda *a, **k: None
        sys.modules["torchinfo"] = ti
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    from agent.ctrlsac import ctrlsac_agent
    from agent.diffsrsac import diffsrsac_agent
    from agent.sac import sac_agent
    from agent.spedersac import spedersac_agent
    from agent.vlsac import vlsac_agent
    from utils import buffer
    return {"sac": sac_agent.SACAgent, "ctrlsac": ctrlsac_agent.CTRLSACAgent, "vlsac": vlsac_agent.VLSACAgent,
            "spedersac": spedersac_agent.SPEDERSACAgent, "diffsrsac": diffsrsac_agent.DIFFSRSACAgent}, buffer


class Space:
    def __init__(self, A):
        self.low, self.high, self.shape = -np.ones(A, np.float32), np.ones(A, np.float32), (A,)


TARGETS = {"critic": "critic_target", "phi": "phi_target", "f": "f_target"}


def load_into_reference(agent, state):
    groups = {}
    for k, v in state.items():
        mod, rest = k.split(".", 1)
        groups.setdefault(mod, {})[rest] = v.clone()
    for mod, sd in groups.items():
        getattr(agent, mod).load_state_dict(sd)
        tgt = TARGETS.get(mod)
        if tgt and hasattr(agent, tgt):
            getattr(agent, tgt).load_state_dict(sd)


def reference_state(agent, state_keys):
    out = {}
    mods = sorted({k.split(".", 1)[0] for k in state_keys})
    for mod in mods:
        for k, v in getattr(agent, mod).state_dict().items():
            out[f"{mod}.{k}"] = v.detach().clone()
        tgt = TARGETS.get(mod)
        if tgt and hasattr(agent, tgt):
            for k, v in getattr(agent, tgt).state_dict().items():
                out[f"{tgt}.{k}"] = v.detach().clone()
    out["log_alpha"] = agent.log_alpha.detach().clone()
    return out


def tensor_record(t: torch.Tensor):
    f = t.detach().double().flatten()
    stride = max(1, f.numel() // 256)
    return np.array([f.sum().item(), f.norm().item(), f.abs().max().item()]), f[::stride][:256].numpy()


def run_case(name, classes, buffer_mod):
    from oracle import rl_oracle as O
    alg, shp, kw, B, rows, n = CASES[name]
    S, A = shp["S"], shp["A"]
    init = O.init_state(alg, S, A, kw, seed=0)
    torch.manual_seed(0)
    np.random.seed(0)
    agent = classes[alg](state_dim=S, action_dim=A, action_space=Space(A), discount=0.99, tau=0.005, **kw)
    load_into_reference(agent, init)
    extra = {}
    if alg == "vlsac":
        extra["critic_noise"] = agent.critic.noise.clone()
        agent.critic_target.noise = agent.critic.noise  # deepcopy already shares values; keep explicit
    ring = O.synthetic_ring(S, A, rows, seed=0, ring_cls=buffer_mod.ReplayBuffer)
    np.random.seed(1)
    torch.manual_seed(1)
    infos = []
    for _ in range(n):
        info = agent.train(ring, B)
        infos.append({k: float(v) for k, v in info.items()})
    ref_sd = reference_state(agent, init.keys())

    # --- pin the oracle against what the reference just produced
    oracle = O.ORACLES[alg](S, A, init, discount=0.99, tau=0.005, **kw, **extra)
    oring = O.synthetic_ring(S, A, rows, seed=0)
    np.random.seed(1)
    torch.manual_seed(1)
    oinfos = [oracle.train(oring, B) for _ in range(n)]
    worst_info = 0.0
    for ri, oi in zip(infos, oinfos):
        assert set(ri) == set(oi), (sorted(ri), sorted(oi))
        for k in ri:
            # fp32 scalars: allow 1-ulp noise on near-zero means (atol) on top of a tight relative bound
            worst_info = max(worst_info, max(0.0, abs(ri[k] - oi[k]) - 1e-7) / (abs(ri[k]) + 1e-12))
    osd = oracle.state_dict()
    worst_param = 0.0
    for k, v in ref_sd.items():
        assert k in osd, k
        d = (osd[k].double() - v.double()).norm().item() / (v.double().norm().item() + 1e-30)
        worst_param = max(worst_param, d)
    print(f"{name}: oracle vs reference  worst info rel {worst_info:.2e}  worst param rel-l2 {worst_param:.2e}")
    # params: 1-ulp gradient noise moves Adam steps by up to 2*lr on near-zero-gradient elements
    assert worst_info < 1e-5 and worst_param < 5e-6, "oracle does not restate the reference"

    arrays = {"infos_json": np.frombuffer(json.dumps(infos).encode(), dtype=np.uint8)}
    meta = dict(alg=alg, S=S, A=A, kwargs=kw, batch=B, rows=rows, n=n, keys=list(ref_sd.keys()))
    arrays["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    for k, v in ref_sd.items():
        stats, sample = tensor_record(v)
        arrays["stats/" + k] = stats
        arrays["sample/" + k] = sample
    for k, v in extra.items():
        arrays["extra/" + k] = v.numpy()
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / f"{name}.npz", **arrays)


def main(argv):
    classes, buffer_mod = import_reference()
    names = argv or list(CASES)
    for name in names:
        run_case(name, classes, buffer_mod)


if __name__ == "__main__":
    main(sys.argv[1:])
