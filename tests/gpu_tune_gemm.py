"""GPU tuning sweep: time per launch for the ctrlsac GEMM shapes over (bn, split_k). Run under gpurun."""
import sys
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
import torch
from rlrep_b200 import _lib

dev = "cuda"
# (label, M, N, K, a_mn, b_mn)
shapes = [
    ("fwd  l2   [256,1024]x[1024,1024]^T", 256, 1024, 1024, 0, 0),
    ("fwd  l3   [256,1024]x[2048,1024]^T", 256, 2048, 1024, 0, 0),
    ("logits    [256,2048]x[256,2048]^T ", 256, 256, 2048, 0, 0),
    ("dgrad l3  [256,2048]x[2048,1024]  ", 256, 1024, 2048, 0, 1),
    ("dgrad l2  [256,1024]x[1024,1024]  ", 256, 1024, 1024, 0, 1),
    ("wgrad l3  [2048,256]x[256,1024]   ", 2048, 1024, 256, 1, 1),
    ("wgrad l2  [1024,256]x[256,1024]   ", 1024, 1024, 256, 1, 1),
    ("dzphi     [256,256]x[256,2048]    ", 256, 2048, 256, 0, 1),
    ("dzmu      [256,256]^Tx[256,2048]  ", 256, 2048, 256, 1, 1),
    ("critic l1l4 [512,2048]x[2048,2048]^T", 512, 2048, 2048, 0, 0),
    ("big   [2048,2048]x[16384,2048]^T  ", 2048, 16384, 2048, 0, 0),
]
for label, M, N, K, a_mn, b_mn in shapes:
    A = torch.randn((K, M) if a_mn else (M, K), device=dev)
    B = torch.randn((K, N) if b_mn else (N, K), device=dev)
    Cm = torch.empty((M, N), device=dev)
    ws = torch.empty(16 * M * N if M * N < (1 << 22) else 1, device=dev)
    res = []
    for bn in (32, 64, 128, 256):
        for sk in (1, 2, 4, 8):
            if sk > 1 and M * N >= (1 << 22):
                continue
            ms, bno, so = _lib.gemm_bench(A, B, Cm, a_mn=bool(a_mn), b_mn=bool(b_mn), bn=bn, split_k=sk, ws=ws, iters=200)
            if so != sk:
                continue
            res.append((ms * 1e3, bn, sk))
    res.sort()
    ms_auto, bn_a, sk_a = _lib.gemm_bench(A, B, Cm, a_mn=bool(a_mn), b_mn=bool(b_mn), ws=ws, iters=200)
    flops = 2.0 * M * N * K
    best = res[0]
    print(f"{label}: best {best[0]:.1f}us (bn={best[1]},sk={best[2]}) {flops / best[0] / 1e6:.1f} TF/s | auto {ms_auto * 1e3:.1f}us (bn={bn_a},sk={sk_a}) | top3 "
          + ", ".join(f"{t:.1f}us@bn{b}/sk{s}" for t, b, s in res[:3]) + f" | worst {res[-1][0]:.1f}us", flush=True)
