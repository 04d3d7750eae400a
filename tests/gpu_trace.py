import sys
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
import ctypes as C, numpy as np, torch
from rlrep_b200 import _lib
lib=_lib.load()
names=["entry","setup done","first operands landed","last MMA issued","accum complete","staged","sync passed","stores issued","exit","store fn entry","dsmem loaded","arrived"]
for (M,N,K,a_mn,b_mn,bn,sk) in [(256,256,256,0,0,128,1),(256,1024,1024,0,0,64,2),(256,1024,1024,0,0,128,8),(256,1024,1024,0,0,64,4),(256,1024,2048,0,1,64,4),(2048,1024,256,1,1,128,1)]:
    A=torch.randn((K,M) if a_mn else (M,K),device="cuda"); B=torch.randn((K,N) if b_mn else (N,K),device="cuda"); Cm=torch.empty(M,N,device="cuda")
    for _ in range(10): _lib.gemm(A,B,Cm,a_mn=bool(a_mn),b_mn=bool(b_mn),bn=bn,split_k=sk)
    t=np.zeros(16,dtype=np.uint64); _lib.check(lib.rlrep_gemm_trace(t.ctypes.data))
    t=t.astype(np.int64); base=t[0]
    print(f"M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} bn={bn} sk={sk}: "+", ".join(f"{n}={int(t[i]-base)}ns" for i,n in enumerate(names)))
