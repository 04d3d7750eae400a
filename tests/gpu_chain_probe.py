"""In-kernel timeline of a GEMM chain (csrc/gemm_chain.cu debug stamps): per GEMM of the chain, when its items were picked
up, had their dependencies, had operands, finished accumulating, were drained and published -- relative to the kernel's
first stamp.   python tests/gpu_chain_probe.py [fwd|bwd] [bn] [split]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import ctypes as C
import numpy as np
import torch
from rlrep_b200 import _lib

lib = _lib.load()
which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
bn = int(sys.argv[2]) if len(sys.argv) > 2 else 0
split = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = "cuda"
torch.manual_seed(0)
B, S, H, D = 256, 24, 1024, 2048
r = lambda *s: torch.randn(*s, device=dev)
specs = []
if which == "fwd":
    x, y = r(B, S), r(B, S)
    for inp in (x, y):
        W = [r(H, S) / 5, r(H, H) / 32, r(D, H) / 32]
        b = [r(H), r(H), r(D)]
        o = [torch.empty(B, H, device=dev), torch.empty(B, H, device=dev), torch.empty(B, D, device=dev)]
        src = inp
        for l in range(3):
            specs.append((src, W[l], o[l], dict(epi=_lib.make_epilogue(bias=b[l], act="elu"))))
            src = o[l]
        if inp is x:
            zp = o[2]
        else:
            zm = o[2]
    logits = torch.empty(B, B, device=dev)
    specs.append((zp, zm, logits, {}))
else:  # the backward GEMM group of one CTRL feature step: dz (2) + per net [dgrad3, wgrad3, dgrad2, wgrad2, wgrad1]
    G = r(B, B) / 16
    zp, zm = r(B, D), torch.tanh(r(B, D))
    dzp, dzm = torch.empty(B, D, device=dev), torch.empty(B, D, device=dev)
    specs.append((G, zm, dzp, dict(b_mn=True)))
    specs.append((G, zp, dzm, dict(a_mn=True, b_mn=True, epi=_lib.make_epilogue(aux=zm, dact="tanh_out"))))
    for dz in (dzp, dzm):
        x = r(B, 32)
        h1, h2 = torch.nn.functional.elu(r(B, H)), torch.nn.functional.elu(r(B, H))
        W3, W2 = r(D, H) / 32, r(H, H) / 32
        dh2, dh1 = torch.empty(B, H, device=dev), torch.empty(B, H, device=dev)
        dW3, dW2, dW1 = torch.empty(D, H, device=dev), torch.empty(H, H, device=dev), torch.empty(H, 32, device=dev)
        specs.append((dz, W3, dh2, dict(b_mn=True, epi=_lib.make_epilogue(aux=h2, dact="elu_out"))))
        specs.append((dz, h2, dW3, dict(a_mn=True, b_mn=True)))
        specs.append((dh2, W2, dh1, dict(b_mn=True, epi=_lib.make_epilogue(aux=h1, dact="elu_out"))))
        specs.append((dh2, h1, dW2, dict(a_mn=True, b_mn=True)))
        specs.append((dh1, x, dW1, dict(a_mn=True, b_mn=True)))

ITEMS, EV = 16, 10
dbg = torch.zeros(148 * ITEMS * EV, dtype=torch.int64, device=dev)
_lib.check(lib.rlrep_gemm_chain_set_debug(dbg.data_ptr()))
ms, levels = _lib.gemm_chain(specs, bn=bn, split_k=split, iters=4)
torch.cuda.synchronize()
_lib.check(lib.rlrep_gemm_chain_set_debug(None))
d = dbg.cpu().numpy().reshape(148, ITEMS, EV)
entry, exit_ = d[:, ITEMS - 1, 2].copy(), d[:, ITEMS - 1, 3].copy()
d[:, ITEMS - 1, 2:4] = 0
valid = d[:, :, 0] > 0
t0 = d[:, :, 0][valid].min()
print(f"{which} chain: {len(specs)} GEMMs, {levels} levels, {ms * 1e3:.1f} us per launch (bn={bn}, split={split}); "
      f"{int(valid.sum())} items stamped, last publish at {(d[:, :, [5, 7]].max() - t0) / 1e3:.1f} us")
print(f"  kernel entry {(entry[entry > 0].min() - t0) / 1e3:.1f}..{(entry.max() - t0) / 1e3:.1f} us, "
      f"exit {(exit_[exit_ > 0].min() - t0) / 1e3:.1f}..{(exit_.max() - t0) / 1e3:.1f} us (relative to the first item pick)")
names = ["pick", "deps", "tma", "opnd", "acc", "epi", "arrive", "publish", None, "stored"]
for g in range(len(specs)):
    sel = valid & ((d[:, :, 8] & 0xFFFF) == g)
    if not sel.any():
        continue
    rows = []
    for e in (0, 1, 2, 3, 4, 5, 6, 9, 7):
        v = d[:, :, e][sel]
        v = v[v > 0]
        rows.append(f"{names[e]} {((v.min() - t0) / 1e3):6.1f}..{((v.max() - t0) / 1e3):6.1f}" if len(v) else f"{names[e]}   -")
    A, Bm, Cm, kw = specs[g]
    print(f"  gemm {g:2d} items {int(sel.sum()):3d} | " + " | ".join(rows))
# per-phase medians over all items (us)
def med(a, b):
    x = (d[:, :, b] - d[:, :, a])[valid & (d[:, :, a] > 0) & (d[:, :, b] > 0)]
    return float(np.median(x)) / 1e3 if len(x) else float("nan")
print(f"  medians: deps wait {med(0, 1):.2f} | deps->tma {med(1, 2):.2f} | deps->operands {med(1, 3):.2f} | operands->acc {med(3, 4):.2f} "
      f"| acc->epi {med(4, 5):.2f} | epi->arrive {med(5, 6):.2f} | epi->stored {med(5, 9):.2f} | stored->publish {med(9, 7):.2f}")
