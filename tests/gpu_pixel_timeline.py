"""Per-launch CUPTI timeline of one pixel-agent update at the bench size (through bench.py's own agent setup): every kernel
launch above a threshold with its duration, in launch order, plus per-kernel totals.
    python tests/gpu_pixel_timeline.py [ldiffsr_pixels_b256|mulvdrq_pixels_b256|drqv2_pixels_b256] [min_us]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from torch.profiler import ProfilerActivity, profile
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "ldiffsr_pixels_b256"
min_us = float(sys.argv[2]) if len(sys.argv) > 2 else 100.0
w = bench.WORKLOADS[name]
arm = bench._PixelArm(w, "tf32", 0)
batches = [tuple(arm.D.synthetic_pixel_batch(w["B"], w["C"], 84, w["A"], seed=i)) for i in range(2)]
torch.manual_seed(1)
for i in range(3):
    arm.step(batches[i % 2], i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    arm.step(batches[0], 4)
    torch.cuda.synchronize()
ev = sorted((e for e in prof.events() if e.device_type.name == "CUDA"), key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
tot = {}
for i, e in enumerate(ev):
    nm = bench.kernel_label(e.name) if hasattr(bench, "kernel_label") else e.name[:40]
    d = e.time_range.end - e.time_range.start
    a = tot.setdefault(nm, [0.0, 0])
    a[0] += d; a[1] += 1
    if d >= min_us:
        print(f"{i:4d} {(e.time_range.start - t0):10.1f} {d:9.1f} us  {nm}  {e.name[-60:] if 'gemm' in nm else ''}")
print("totals:")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0])[:30]:
    print(f"  {k:32s} {v[0]:10.1f} us  x{v[1]}")
