"""Golden fixtures for the muLV-Rep DrQ-v2 pixel update: runs the REAL reference class (agent/mulvdrq/drqv2.py) in the
build container, checks oracle/mulv_oracle.py against it (parameters after the updates -- the reference's update()
returns no metrics with use_tb=False) and writes tests/golden/mulvdrq_*.npz.

    python -m oracle.make_golden_mulv          (needs /root/reference; never runs on the GPU box)"""
from __future__ import annotations

import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference/agent/mulvdrq")
OUT = ROOT / "tests" / "golden"
CASES = {"mulvdrq_b4": (9, 4, 100, 64, 4, 4)}  # channels, action_dim, feat_dim, hid_dim, batch, update() calls


def import_reference():
    for name in ("hydra", "omegaconf", "matplotlib", "matplotlib.pyplot", "termcolor"):
        sys.modules.setdefault(name, types.ModuleType(name))  # imported, never used on the update path
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["omegaconf"].OmegaConf = type("OmegaConf", (), {})
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    import drqv2
    from mulv_config import config
    return drqv2.DrQV2Agent, config


def tensor_record(t):
    f = t.detach().double().flatten()
    stride = max(1, f.numel() // 256)
    return np.array([f.sum().item(), f.norm().item(), f.abs().max().item()]), f[::stride][:256].numpy()


def run_case(name, Agent, cfg):
    from oracle import mulv_oracle as M
    C, A, Fd, H, B, n = CASES[name]
    cfg = type(cfg)(cfg)  # the reference's attribute-access dict
    cfg.update(device="cpu", feat_dim=Fd, hid_dim=H, use_tb=True)
    init = M.init_state(C, A, Fd, H, seed=0)
    torch.manual_seed(0)
    ref = Agent((C, 84, 84), (A,), cfg)
    mods = ("encoder", "decoder", "actor", "critic", "predict_encoder", "feat_encoder", "feat_decoder", "feat_f")
    for mod in mods:
        getattr(ref, mod).load_state_dict({k[len(mod) + 1:]: v.clone() for k, v in init.items() if k.startswith(mod + ".")})
    ref.encoder_target.load_state_dict(ref.encoder.state_dict())
    ref.critic_target.load_state_dict(ref.critic.state_dict())
    ref.feat_f_target.load_state_dict(ref.feat_f.state_dict())
    batches = [M.synthetic_pixel_batch(B, C, 84, A, seed=20 + i) for i in range(n)]
    torch.manual_seed(1)
    infos = [{k: float(v) for k, v in ref.update(iter([tuple(b)]), step=2 * i).items()} for i, b in enumerate(batches)]
    ref_sd = {}
    for mod in mods + ("encoder_target", "critic_target", "feat_f_target"):
        for k, v in getattr(ref, mod).state_dict().items():
            ref_sd[f"{mod}.{k}"] = v.detach().clone()

    oracle = M.OracleMuLVDrQ(A, init)
    torch.manual_seed(1)
    oinfos = [oracle.update(b, step=2 * i) for i, b in enumerate(batches)]
    worst_info = max(abs(ri["actor_loss"] - oi["actor_loss"]) / (abs(ri["actor_loss"]) + 1e-12) for ri, oi in zip(infos, oinfos))
    osd = oracle.state_dict()
    assert set(osd) == set(ref_sd), sorted(set(osd) ^ set(ref_sd))
    worst_param = max((osd[k].double() - v.double()).norm().item() / (v.double().norm().item() + 1e-30)
                      for k, v in ref_sd.items())
    print(f"{name}: oracle vs reference  actor_loss rel {worst_info:.2e}  worst param rel-l2 {worst_param:.2e}")
    assert worst_info < 1e-5 and worst_param < 5e-6, "oracle does not restate the reference"
    arrays = {"infos_json": np.frombuffer(json.dumps(oinfos).encode(), dtype=np.uint8),
              "meta_json": np.frombuffer(json.dumps(dict(C=C, A=A, feat_dim=Fd, hid_dim=H, batch=B, n=n,
                                                         keys=list(ref_sd))).encode(), dtype=np.uint8)}
    for k, v in ref_sd.items():
        arrays["stats/" + k], arrays["sample/" + k] = tensor_record(v)
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / f"{name}.npz", **arrays)


if __name__ == "__main__":
    Agent, cfg = import_reference()
    for name in (sys.argv[1:] or list(CASES)):
        run_case(name, Agent, cfg)
