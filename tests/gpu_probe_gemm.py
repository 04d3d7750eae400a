"""GPU probe for the GEMM kernels: prints the relative error of every operand-major variant and tile shape.

Run on the GPU box:  python tests/gpu_probe_gemm.py            (drives one subprocess per variant, 120 s cap each)
One variant only:    python tests/gpu_probe_gemm.py tc 0 1 128 0
A kernel that traps kills only its own subprocess, so the remaining variants still report.
"""
import subprocess
import sys
import time

sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))


def run_variant(path, a_mn, b_mn, bn, split_k):
    import torch
    from rlrep_b200 import _lib
    torch.manual_seed(0)
    dev = "cuda"
    shapes = [(256, 2048, 1024), (256, 256, 2048), (2048, 1024, 256), (128, 128, 32), (224, 160, 100),
              (256, 1024, 2048), (1024, 23 * 4, 256)]
    out = []
    for (M, N, K) in shapes:
        A = torch.randn((K, M) if a_mn else (M, K), device=dev)
        B = torch.randn((K, N) if b_mn else (N, K), device=dev)
        Cm = torch.full((M, N), float("nan"), device=dev)
        ws = torch.empty(16 * M * N, device=dev) if split_k != 1 else None
        bias = torch.randn(N, device=dev)
        epi = _lib.make_epilogue(bias=bias, act="elu")
        _lib.gemm(A, B, Cm, a_mn=a_mn, b_mn=b_mn, path=path, epi=epi, bn=bn, split_k=split_k, ws=ws)
        torch.cuda.synchronize()
        Am = (A.t() if a_mn else A).double()
        Bm = (B.t() if b_mn else B).double()
        ref = torch.nn.functional.elu(Am @ Bm.t() + bias.double())
        err = ((Cm.double() - ref).norm() / ref.norm()).item()
        maxabs = (Cm.double() - ref).abs().max().item()
        out.append(f"  M={M} N={N} K={K}: rel_fro={err:.3e} max_abs={maxabs:.3e} nan={int(torch.isnan(Cm).sum())}")
    # timing of the weight-streaming shape
    M, N, K = 256, 2048, 1024
    A = torch.randn((K, M) if a_mn else (M, K), device=dev)
    B = torch.randn((K, N) if b_mn else (N, K), device=dev)
    Cm = torch.empty((M, N), device=dev)
    ws = torch.empty(16 * M * N, device=dev) if split_k != 1 else None
    for _ in range(5):
        _lib.gemm(A, B, Cm, a_mn=a_mn, b_mn=b_mn, path=path, bn=bn, split_k=split_k, ws=ws)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        _lib.gemm(A, B, Cm, a_mn=a_mn, b_mn=b_mn, path=path, bn=bn, split_k=split_k, ws=ws)
    e1.record()
    torch.cuda.synchronize()
    out.append(f"  time 256x2048x1024 (incl. per-call tensor-map encode): {e0.elapsed_time(e1) / 50 * 1e3:.1f} us")
    print("\n".join(out))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_variant(sys.argv[1], bool(int(sys.argv[2])), bool(int(sys.argv[3])), int(sys.argv[4]), int(sys.argv[5]))
        sys.exit(0)
    variants = []
    for a_mn in (0, 1):
        for b_mn in (0, 1):
            variants.append(("tc", a_mn, b_mn, 128, 1))
    variants += [("tc", 0, 0, 32, 1), ("tc", 0, 0, 64, 1), ("tc", 0, 0, 256, 1), ("tc", 0, 1, 64, 1),
                 ("tc", 1, 1, 256, 1), ("tc", 0, 0, 0, 0), ("tc", 0, 1, 0, 0), ("tc", 1, 1, 0, 0),
                 ("simt", 0, 0, 0, 1), ("simt", 0, 1, 0, 1), ("simt", 1, 1, 0, 1), ("simt", 1, 0, 0, 1)]
    for v in variants:
        print(f"== path={v[0]} a_mn={v[1]} b_mn={v[2]} bn={v[3]} split_k={v[4]}", flush=True)
        t = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, *map(str, v)], capture_output=True, text=True, timeout=120)
            print(r.stdout.rstrip())
            if r.returncode != 0:
                print(f"  FAILED rc={r.returncode}\n  " + "\n  ".join(r.stderr.strip().splitlines()[-6:]))
        except subprocess.TimeoutExpired:
            print("  TIMEOUT")
        print(f"  ({time.time() - t:.1f}s)", flush=True)
