"""Warm (graph-replayed) timings of the CUDA-core GEMM on the skinny shapes of the update step."""
import sys
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
import torch
from rlrep_b200 import _lib
shapes = [("fwd l1 K=23", 256, 1024, 23, 0, 0), ("wgrad l1 N=23", 1024, 23, 256, 1, 1), ("d_action N=6", 256, 6, 1024, 0, 1),
          ("actor head N=12", 256, 12, 256, 0, 0), ("actor head wgrad M=12", 12, 256, 256, 1, 1),
          ("actor l0 K=17", 256, 256, 17, 0, 0), ("actor head dgrad K=12", 256, 256, 12, 0, 1),
          ("actor l1 256^3 simt", 256, 256, 256, 0, 0)]
for label, M, N, K, a_mn, b_mn in shapes:
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda")
    B = torch.randn((K, N) if b_mn else (N, K), device="cuda")
    C = torch.empty((M, N), device="cuda")
    ms, _, _ = _lib.gemm_bench(A, B, C, a_mn=bool(a_mn), b_mn=bool(b_mn), path="simt", iters=200)
    print(f"{label}: {ms*1e3:.1f} us")
A = torch.randn(256, 256, device="cuda"); B = torch.randn(256, 256, device="cuda"); C = torch.empty(256, 256, device="cuda")
ws = torch.empty(16*256*256, device="cuda")
for bn, sk in ((32, 1), (64, 1), (128, 1), (64, 2), (32, 2)):
    ms, bno, so = _lib.gemm_bench(A, B, C, bn=bn, split_k=sk, ws=ws, iters=200)
    print(f"tc 256^3 bn={bno} sk={so}: {ms*1e3:.1f} us")
