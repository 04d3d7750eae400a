"""Per-launch profile of one muLV-Rep DrQ-v2 update at full size: prints every launch above 40 us with its algorithmic
flops / bytes (2MNK identifies a GEMM's shape) and the per-kernel totals.  python tests/gpu_mulv_profile.py [B] [H]"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import mulv_oracle as M  # noqa: E402  (synthetic batch + initial weights only)
from rlrep_b200 import _lib  # noqa: E402
from rlrep_b200.pixel import MuLVDrQv2  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
cfg = dict(aug=True, pre_aug=False, back_q2feat=True, tanh=True, both_q=False, q_activ="relu", q_loss="huber", q_up_n=1,
           l2_norm=0.0, c_targ_tau=0.01, up_every=1, feat_dim=100, hid_dim=H)
agent = MuLVDrQv2((9, 84, 84), (4,), cfg)
agent.load_state_dict(M.init_state(9, 4, 100, H, seed=0))
b = tuple(M.synthetic_pixel_batch(B, 9, 84, 4, seed=0))
torch.manual_seed(0)
for i in range(3):
    agent.update(iter([b]), step=i)
cap = 4096
names, ms = (C.c_char_p * cap)(), (C.c_float * cap)()
by, fl, n = (C.c_double * cap)(), (C.c_double * cap)(), C.c_int()
_lib.check(agent.lib.rlrep_mulv_profile_update(agent._h, 1.0, cap, names, ms, by, fl, C.byref(n)))
tot = {}
total = 0.0
for i in range(n.value):
    nm = names[i].decode()
    t = tot.setdefault(nm, [0.0, 0])
    t[0] += ms[i]
    t[1] += 1
    total += ms[i]
    if ms[i] > 0.04:
        print(f"{i:4d} {nm:24s} {ms[i] * 1e3:9.1f} us  flops {fl[i]:.3e}  bytes {by[i]:.3e}"
              + (f"  -> {fl[i] / ms[i] / 1e9:7.1f} TF/s" if fl[i] else "") + (f"  {by[i] / ms[i] / 1e6:7.0f} GB/s" if by[i] else ""))
print(f"total {total:.3f} ms over {n.value} launches")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:24s} {v[0] * 1e3:9.1f} us  x{v[1]}")
