"""Timing of the pixel encoder (forward + backward, B = 256, 9x84x84 uint8 frames) with CUDA events, next to cuDNN
through PyTorch (reference only) and the reference's own CPU path on the host cores:  python tests/gpu_bench_conv.py"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1])); sys.path.insert(0, str(Path(__file__).resolve().parent))
import torch
from rlrep_b200.pixel import ConvEncoder
from test_gpu_conv import ref_aug, ref_encoder

B, C = 256, 9
g = torch.Generator().manual_seed(0)
sd = {}
for i, cin in zip((0, 2, 4, 6), (C, 32, 32, 32)):
    sd[f"convnet.{i}.weight"] = torch.randn(32, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5
    sd[f"convnet.{i}.bias"] = torch.zeros(32)
obs = torch.randint(0, 256, (B, C, 84, 84), generator=g, dtype=torch.uint8).cuda()
shift = torch.randint(0, 9, (B, 1, 1, 2), generator=g).cuda()
dfeat = torch.randn(B, 32 * 35 * 35, generator=g).cuda()


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / n


for prec in ("tf32", "fp32"):
    enc = ConvEncoder((C, 84, 84), batch=B, precision=prec)
    enc.load_state_dict(sd)
    s2 = shift.reshape(B, 2).int().contiguous()
    f = timed(lambda: enc.forward(obs, s2))
    fb = timed(lambda: (enc.forward(obs, s2), enc.backward(dfeat)))
    print(f"rlrep conv encoder {prec}: forward {f:.3f} ms, forward+backward {fb:.3f} ms (B={B})")
    enc.close()

rsd = {k: v.cuda().requires_grad_() for k, v in sd.items()}
def torch_fb():
    for v in rsd.values():
        v.grad = None
    ref_encoder(rsd, ref_aug(obs.float(), shift)).backward(dfeat)
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    print(f"torch/cuDNN (allow_tf32={tf32}): forward+backward {timed(torch_fb):.3f} ms")

csd = {k: v.detach().cpu().requires_grad_() for k, v in sd.items()}
ocpu, scpu, dcpu = obs.cpu(), shift.cpu(), dfeat.cpu()
t0 = time.perf_counter()
for _ in range(2):
    ref_encoder(csd, ref_aug(ocpu.float(), scpu)).backward(dcpu)
print(f"reference modules on {torch.get_num_threads()} host threads: forward+backward {(time.perf_counter() - t0) / 2 * 1e3:.0f} ms")
