"""Golden fixtures for the plain DrQ-v2 pixel update: runs the REAL reference class (agent/diffsrdrq/drqv2.py) in the
build container, checks oracle/drq_oracle.py against it and writes tests/golden/drqv2_*.npz.

    python -m oracle.make_golden_drq          (needs /root/reference; never runs on the GPU box)"""
from __future__ import annotations

import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference/agent/diffsrdrq")
OUT = ROOT / "tests" / "golden"

# name -> (channels, action_dim, bn_dim, hidden_dim, batch, number of train_step calls)
CASES = {"drqv2_b8": (9, 4, 50, 256, 8, 4), "drqv2_b16_c3": (3, 6, 32, 128, 16, 2)}


def import_reference():
    u, m, d = types.ModuleType("UtilsRL"), types.ModuleType("UtilsRL.misc"), types.ModuleType("UtilsRL.misc.decorator")
    d.profile = lambda f: f  # imported, never applied (drqv2.py:7)
    sys.modules.update({"UtilsRL": u, "UtilsRL.misc": m, "UtilsRL.misc.decorator": d})
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    import drqv2
    return drqv2.DrQv2


class Box:
    def __init__(self, shape):
        self.shape = shape
        self.minimum, self.maximum = -np.ones(shape, np.float32), np.ones(shape, np.float32)


def tensor_record(t):
    f = t.detach().double().flatten()
    stride = max(1, f.numel() // 256)
    return np.array([f.sum().item(), f.norm().item(), f.abs().max().item()]), f[::stride][:256].numpy()


def run_case(name, DrQv2):
    from oracle import drq_oracle as D
    C, A, bn, H, B, n = CASES[name]
    args = types.SimpleNamespace(tau=0.01, update_every=2, device="cpu", critic_loss="mse",
                                 stddev_schedule="linear(1.0,0.1,500000)", stddev_clip=0.3, bn_dim=bn,
                                 actor_hidden_dim=H, critic_hidden_dim=H, encoder_lr=1e-4, actor_lr=1e-4, critic_lr=1e-4)
    init = D.init_state(C, A, bn, H, seed=0)
    torch.manual_seed(0)
    ref = DrQv2(Box((C, 84, 84)), Box((A,)), args)
    for mod in ("encoder", "actor", "critic"):
        sd = {k[len(mod) + 1:]: v.clone() for k, v in init.items() if k.startswith(mod + ".")}
        getattr(ref, mod).load_state_dict(sd)
    ref.critic_target.load_state_dict(ref.critic.state_dict())
    batches = [D.synthetic_pixel_batch(B, C, 84, A, seed=10 + i) for i in range(n)]
    torch.manual_seed(1)
    infos = [{k: float(v) for k, v in ref.train_step(iter([tuple(b)]), step=1000 * i).items()} for i, b in enumerate(batches)]
    ref_sd = {}
    for mod in ("encoder", "actor", "critic", "critic_target"):
        for k, v in getattr(ref, mod).state_dict().items():
            ref_sd[f"{mod}.{k}"] = v.detach().clone()

    oracle = D.OracleDrQv2(A, init)
    torch.manual_seed(1)
    oinfos = [oracle.train_step(b, step=1000 * i) for i, b in enumerate(batches)]
    worst_info = 0.0
    for ri, oi in zip(infos, oinfos):
        assert set(ri) == set(oi), (sorted(ri), sorted(oi))
        for k in ri:
            worst_info = max(worst_info, max(0.0, abs(ri[k] - oi[k]) - 1e-7) / (abs(ri[k]) + 1e-12))
    osd = oracle.state_dict()
    worst_param = max((osd[k].double() - v.double()).norm().item() / (v.double().norm().item() + 1e-30)
                      for k, v in ref_sd.items())
    print(f"{name}: oracle vs reference  worst info rel {worst_info:.2e}  worst param rel-l2 {worst_param:.2e}")
    assert worst_info < 1e-5 and worst_param < 5e-6, "oracle does not restate the reference"

    arrays = {"infos_json": np.frombuffer(json.dumps(infos).encode(), dtype=np.uint8),
              "meta_json": np.frombuffer(json.dumps(dict(C=C, A=A, bn_dim=bn, hidden_dim=H, batch=B, n=n,
                                                         keys=list(ref_sd))).encode(), dtype=np.uint8)}
    for k, v in ref_sd.items():
        arrays["stats/" + k], arrays["sample/" + k] = tensor_record(v)
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / f"{name}.npz", **arrays)


if __name__ == "__main__":
    cls = import_reference()
    for name in (sys.argv[1:] or list(CASES)):
        run_case(name, cls)
