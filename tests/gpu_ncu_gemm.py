"""A handful of single GEMM launches for `ncu --metrics gpu__time_duration.sum` (kernel-only durations)."""
import sys
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
import torch
from rlrep_b200 import _lib
dev = "cuda"
cases = [(256, 2048, 1024, 0, 0, 32, 1), (256, 2048, 1024, 0, 0, 128, 1), (256, 2048, 1024, 0, 0, 128, 4),
         (256, 2048, 1024, 0, 0, 128, 8), (256, 1024, 2048, 0, 1, 64, 4), (256, 1024, 2048, 0, 1, 128, 8),
         (2048, 1024, 256, 1, 1, 128, 1), (256, 256, 2048, 0, 0, 32, 8), (128, 128, 32, 0, 0, 128, 1),
         (128, 128, 1024, 0, 0, 128, 1)]
for (M, N, K, a_mn, b_mn, bn, sk) in cases:
    A = torch.randn((K, M) if a_mn else (M, K), device=dev)
    B = torch.randn((K, N) if b_mn else (N, K), device=dev)
    Cm = torch.empty((M, N), device=dev)
    ws = torch.empty(16 * M * N, device=dev)
    for _ in range(3):
        _lib.gemm(A, B, Cm, a_mn=bool(a_mn), b_mn=bool(b_mn), bn=bn, split_k=sk, ws=ws)
    torch.cuda.synchronize()
