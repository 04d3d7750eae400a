"""Golden fixture for the latent Diff-SR DrQ-v2 pixel update: runs the REAL reference class
(agent/diffsrdrq/latent_diff_sr.py) in the build container, checks oracle/ldiffsr_oracle.py against it and writes
tests/golden/ldiffsr_*.npz.

    python -m oracle.make_golden_ldiffsr          (needs /root/reference; never runs on the GPU box)"""
from __future__ import annotations

import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

from oracle.make_golden_drq import Box, import_reference, tensor_record

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "tests" / "golden"
# name -> (Dims(A, L, feat, bn, psi_h, psi_d, zeta_h, zeta_d, H), batch, number of train_step calls)
CASES = {"ldiffsr_b4": ((4, 16, 32, 24, 32, 2, 32, 4, 64), 4, 4)}
MODS = {"vae": "vae", "score": "score", "actor": "actor", "critic": "critic", "vae_target": "vae_target",
        "score_target": "score_target", "critic_target": "critic_target"}


def run_case(name):
    from oracle import ldiffsr_oracle as O
    import_reference()
    import latent_diff_sr as L
    dims, B, n = CASES[name]
    d = O.Dims(*dims)
    args = types.SimpleNamespace(
        device="cpu", use_repr_target=True, kl_coef=1.0, reg_coef=0.0, ae_coef=1.0, repr_coef=1.0, tau=0.01, grad_norm=None,
        extra_repr_step=1, update_every=2, stddev_schedule="linear(1.0,0.1,500000)", stddev_clip=0.3, pretrain_steps=10000,
        ae_pretrain_steps=5000, back_critic_grad=True, critic_loss="mse", num_noises=1000, noise_param1=1e-4,
        noise_param2=0.02, noise_schedule="linear", latent_dim=d.L, feature_dim=d.feat, bn_dim=d.bn, ae_num_filters=32,
        ae_num_layers=4, do_scale=False, psi_hidden_dim=d.psi_h, psi_hidden_depth=d.psi_d, zeta_hidden_dim=d.zeta_h,
        zeta_hidden_depth=d.zeta_d, actor_hidden_dim=d.H, critic_hidden_dim=d.H, ae_lr=3e-4, score_lr=3e-4, actor_lr=1e-4,
        critic_lr=1e-4)
    init = O.init_state(d, seed=0)
    torch.manual_seed(0)
    ref = L.LatentDiffSRDrQv2(Box((9, 84, 84)), Box((d.A,)), args)
    for mod in ("vae", "score", "actor", "critic"):
        getattr(ref, mod).load_state_dict({k[len(mod) + 1:]: v.clone() for k, v in init.items() if k.startswith(mod + ".")})
    for mod in ("vae", "score", "critic"):
        getattr(ref, mod + "_target").load_state_dict(getattr(ref, mod).state_dict())
    batches = [O.synthetic_pixel_batch(B, 9, 84, d.A, seed=40 + i) for i in range(n)]
    torch.manual_seed(1)
    infos = [{k: float(v) for k, v in ref.train_step(iter([tuple(b)]), step=1000 * i).items()} for i, b in enumerate(batches)]
    ref_sd = {}
    for mod in MODS:
        for k, v in getattr(ref, mod).state_dict().items():
            ref_sd[f"{mod}.{k}"] = v.detach().clone()

    oracle = O.OracleLatentDiffSR(d, init)
    torch.manual_seed(1)
    oinfos = [oracle.train_step(b, step=1000 * i) for i, b in enumerate(batches)]
    assert [bool(x) for x in oinfos] == [bool(x) for x in infos], "update_every gating differs"
    worst_info = 0.0
    for ri, oi in zip(infos, oinfos):
        assert set(ri) == set(oi), sorted(set(ri) ^ set(oi))
        for k, v in ri.items():
            worst_info = max(worst_info, abs(v - oi[k]) / (abs(v) + 1e-6))
    osd = oracle.state_dict()
    assert set(osd) == set(ref_sd), sorted(set(osd) ^ set(ref_sd))[:10]
    worst_param, wname = max(((osd[k].double() - v.double()).norm().item() / (v.double().norm().item() + 1e-30), k)
                             for k, v in ref_sd.items())
    print(f"{name}: oracle vs reference  worst info rel {worst_info:.2e}  worst param rel-l2 {worst_param:.2e} ({wname})")
    assert worst_info < 2e-5 and worst_param < 5e-6, "oracle does not restate the reference"
    arrays = {"infos_json": np.frombuffer(json.dumps(infos).encode(), dtype=np.uint8),
              "meta_json": np.frombuffer(json.dumps(dict(dims=list(dims), batch=B, n=n, keys=list(ref_sd))).encode(),
                                         dtype=np.uint8)}
    for k, v in ref_sd.items():
        arrays["stats/" + k], arrays["sample/" + k] = tensor_record(v)
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / f"{name}.npz", **arrays)


if __name__ == "__main__":
    for name in (sys.argv[1:] or list(CASES)):
        run_case(name)
