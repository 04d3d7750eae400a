"""ctypes binding of librlrep_b200.so (C ABI declared in include/rlrep_b200.h).

There is exactly one compute path: the CUDA library.  If the shared object is missing or a call fails the
caller gets an exception -- nothing here falls back to PyTorch or to the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "librlrep_b200.so"


class RlrepError(RuntimeError):
    pass


class Epilogue(C.Structure):
    """Mirror of `rlrep_epilogue` (include/rlrep_b200.h)."""
    _fields_ = [
        ("bias", C.c_void_p), ("r1_u", C.c_void_p), ("r1_v", C.c_void_p), ("aux", C.c_void_p),
        ("pre_out", C.c_void_p), ("ld_aux", C.c_int), ("ld_pre", C.c_int), ("act", C.c_int),
        ("dact", C.c_int), ("accumulate", C.c_int), ("scale", C.c_float),
    ]


ACT = {"none": 0, "elu": 1, "relu": 2, "tanh": 3, "sin": 4}
DACT = {"none": 0, "elu_out": 1, "relu_out": 2, "tanh_out": 3, "cos_pre": 4}

_lib = None


def load() -> C.CDLL:
    """Load the library once.  Raises if it has not been built (python -m rlrep_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RlrepError(
            f"{LIB_PATH} not found: build it with `python -m rlrep_b200.build` (needs nvcc). "
            "rlrep_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(str(LIB_PATH))
    lib.rlrep_last_error.restype = C.c_char_p
    lib.rlrep_abi_version.restype = C.c_int
    _declare(lib)
    _lib = lib
    return lib


def _declare(lib: C.CDLL) -> None:
    vp, i, f, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    sigs = {
        "rlrep_gemm": [vp, i, i, i, i, vp, i, i, vp, i, i, vp, i, i, vp, i, C.POINTER(Epilogue), i, i, vp, sz],
        "rlrep_gemm_bench": [vp, i, i, i, i, vp, i, i, vp, i, i, vp, i, C.POINTER(Epilogue), i, i, vp, sz, i,
                             C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_int)],
    }
    for name, argtypes in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int


def check(rc: int) -> None:
    if rc != 0:
        msg = load().rlrep_last_error()
        raise RlrepError(msg.decode() if msg else f"rlrep call failed with code {rc}")


def current_stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def make_epilogue(bias=None, r1=None, aux=None, pre_out=None, act="none", dact="none", accumulate=False,
                  scale=1.0) -> Epilogue:
    e = Epilogue()
    e.bias = bias.data_ptr() if bias is not None else None
    if r1 is not None:
        e.r1_u, e.r1_v = r1[0].data_ptr(), r1[1].data_ptr()
    if aux is not None:
        e.aux, e.ld_aux = aux.data_ptr(), aux.stride(0)
    if pre_out is not None:
        e.pre_out, e.ld_pre = pre_out.data_ptr(), pre_out.stride(0)
    e.act, e.dact = ACT[act], DACT[dact]
    e.accumulate = int(accumulate)
    e.scale = scale
    return e


def gemm(A, B, C_out, *, a_mn=False, b_mn=False, path="tc", epi: Epilogue | None = None, A2=None, bn=0, split_k=0,
         ws=None):
    """C = epilogue(sum_k A(m,k) B(n,k)) on torch CUDA tensors (kernel-level entry, used by tests).

    K-major operand: tensor [rows, K]; MN-major operand: tensor [K, rows]."""
    lib = load()
    M = A.shape[1] if a_mn else A.shape[0]
    K = A.shape[0] if a_mn else A.shape[1]
    N = B.shape[1] if b_mn else B.shape[0]
    K1 = K
    if A2 is not None:
        K += A2.shape[1]
    assert (B.shape[0] if b_mn else B.shape[1]) == K
    epi = epi or make_epilogue()
    rc = lib.rlrep_gemm(current_stream_ptr(), 0 if path == "tc" else 1, M, N, K, A.data_ptr(), A.stride(0), int(a_mn),
                        A2.data_ptr() if A2 is not None else None, A2.stride(0) if A2 is not None else 0, K1,
                        B.data_ptr(), B.stride(0), int(b_mn), C_out.data_ptr(), C_out.stride(0), C.byref(epi), bn,
                        split_k, ws.data_ptr() if ws is not None else None, ws.numel() if ws is not None else 0)
    check(rc)
    return C_out


def gemm_bench(A, B, C_out, *, a_mn=False, b_mn=False, path="tc", epi=None, bn=0, split_k=0, ws=None, iters=50):
    """Average milliseconds per launch of a planned GEMM (tuning aid); returns (ms, bn, split_k)."""
    lib = load()
    M = A.shape[1] if a_mn else A.shape[0]
    K = A.shape[0] if a_mn else A.shape[1]
    N = B.shape[1] if b_mn else B.shape[0]
    epi = epi or make_epilogue()
    ms, bno, so = C.c_float(), C.c_int(), C.c_int()
    rc = lib.rlrep_gemm_bench(current_stream_ptr(), 0 if path == "tc" else 1, M, N, K, A.data_ptr(), A.stride(0),
                              int(a_mn), B.data_ptr(), B.stride(0), int(b_mn), C_out.data_ptr(), C_out.stride(0),
                              C.byref(epi), bn, split_k, ws.data_ptr() if ws is not None else None,
                              ws.numel() if ws is not None else 0, iters, C.byref(ms), C.byref(bno), C.byref(so))
    check(rc)
    return ms.value, bno.value, so.value
