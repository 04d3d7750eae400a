import sys
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
import ctypes as C, torch
from rlrep_b200 import _lib
lib=_lib.load()
A=torch.zeros(64,64,device="cuda"); epi=_lib.make_epilogue()
for path,name in ((2,"null"),(3,"null+PDL")):
    for grid in (1,148,1184):
        ms,bn,so=C.c_float(),C.c_int(),C.c_int()
        _lib.check(lib.rlrep_gemm_bench(torch.cuda.current_stream().cuda_stream, path, grid, 64, 64, A.data_ptr(), 64, 0, A.data_ptr(), 64, 0, A.data_ptr(), 64, C.byref(epi), 0, 0, None, 0, 500, C.byref(ms), C.byref(bn), C.byref(so)))
        print(f"{name} grid={grid}: {ms.value*1e3:.2f} us per launch (graph of 500 dependent launches)")
