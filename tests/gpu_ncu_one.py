"""One GEMM launch (after warm-up) for `ncu --set full --import-source on`:  python tests/gpu_ncu_one.py tc|simt M N K a_mn b_mn bn sk"""
import sys
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
import torch
from rlrep_b200 import _lib
path, M, N, K, a_mn, b_mn, bn, sk = sys.argv[1], *map(int, sys.argv[2:9])
A = torch.randn((K, M) if a_mn else (M, K), device="cuda")
B = torch.randn((K, N) if b_mn else (N, K), device="cuda")
C = torch.empty((M, N), device="cuda")
ws = torch.empty(16 * M * N, device="cuda")
for _ in range(5):
    _lib.gemm(A, B, C, a_mn=bool(a_mn), b_mn=bool(b_mn), path=path, bn=bn, split_k=sk, ws=ws)
torch.cuda.synchronize()
