#!/usr/bin/env python
"""Benchmark of the rl-rep update step on B200 (contract: see the task statement / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one `agent.train(buffer, batch_size)` call = K_f feature iterations + critic + actor/alpha + Polyak.
Default workload (BASELINE.json configs[1]): ctrlsac, HalfCheetah shapes (S=17, A=6), batch 256, feature_dim 2048,
hidden 1024, extra_feature_steps 3; synthetic replay ring of 1,000,000 rows.

Lines printed (rank 0, one JSON object):
  value      updates/s, whole job, device-timed (CUDA events), inputs resident in HBM before the timed region
  e2e        updates/s through the public Python API (`agent.train`) with host-drawn indices/noise: H2D of the
             step's inputs and D2H of its metrics inside the timed region
  roofline   the dominant kernel, event-timed per launch in an eager profiled pass right after the timed region
  cpu_baseline  the oracle port (plain PyTorch CPU restatement of the reference) on the box's host cores
N > 1 (torchrun): the path shards as independent agents ("replicas only", one agent per GPU, no collective in the
data path); time = max over ranks, value = sum of updates / that time.
`--impl reference` times the reference's own CPU implementation of the path (the oracle port with the as-written
[B,B,D] broadcast; /root/reference itself is not available on the GPU box) with all host threads.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SPEDER_MAIN = dict(extra_feature_steps=5, phi_and_mu_lr=1e-5, phi_hidden_dim=512, phi_hidden_depth=1, mu_hidden_dim=512,
                   mu_hidden_depth=0, critic_and_actor_lr=3e-4, critic_and_actor_hidden_dim=256, feature_dim=2048)

WORKLOADS = {
    # BASELINE.json configs[1] -- the configuration the metric is quoted on
    "ctrlsac_hc_b256": dict(alg="ctrlsac", S=17, A=6, B=256, rows=1_000_000,
                            kw=dict(hidden_dim=1024, feature_dim=2048, extra_feature_steps=3)),
    # BASELINE.json configs[0] (the reference's own CPU-runnable case)
    "sac_hc_b256": dict(alg="sac", S=17, A=6, B=256, rows=1_000_000, kw=dict(hidden_dim=256)),
    # BASELINE.json configs[2]
    "vlsac_hum_b1024": dict(alg="vlsac", S=376, A=17, B=1024, rows=200_000,
                            kw=dict(hidden_dim=256, feature_dim=256, extra_feature_steps=3)),
    # the other two state-based agents at what main.py passes (main.py:93-104)
    "spedersac_hc_b256": dict(alg="spedersac", S=17, A=6, B=256, rows=1_000_000, kw=SPEDER_MAIN),
    "diffsrsac_hc_b256": dict(alg="diffsrsac", S=17, A=6, B=256, rows=1_000_000, kw=dict(hidden_dim=256)),
    # first pixel agent (plain DrQ-v2 `train_step`, configs/drqv2.yaml shapes); one replica per GPU
    # at N > 1 is the population harness of BASELINE.json configs[4]
    "drqv2_pixels_b256": dict(alg="drqv2", C=9, A=4, B=256, bn=50, H=1024, rows=0, kw={}),
    # muLV-Rep DrQ-v2 `update` at mulv_config.py's shapes (b_size 256, feat_dim 100, hid_dim 1024) --
    # the per-member update of BASELINE.json configs[4]; one replica per GPU at N > 1
    "mulvdrq_pixels_b256": dict(alg="mulvdrq", C=9, A=4, B=256, F=100, H=1024, rows=0, kw={}),
    # BASELINE.json configs[3]: large-batch CTRL, GLOBAL batch 16384 split by rows over the ranks (strong scaling:
    # the total work is fixed; 2048 rows per GPU at N = 8), mu(s') all-gathered over NVLink
    "ctrlsac_b16384_sharded": dict(alg="ctrlsac", S=17, A=6, B=16384, rows=1_000_000, sharded=True,
                                   kw=dict(hidden_dim=1024, feature_dim=2048, extra_feature_steps=3)),
}


class Space:
    def __init__(self, A):
        self.low, self.high, self.shape = -np.ones(A, np.float32), np.ones(A, np.float32), (A,)


def module_params(w):
    from oracle import rl_oracle as O  # layer table only (shapes), no arithmetic
    return {m: sum(o * i + o for _, o, i in layers) for m, layers in O.layer_table(w["alg"], w["S"], w["A"], w["kw"])}


def algorithmic_bytes(w):
    """HBM bytes one update must move (SURVEY.md 8d / BASELINE.md): Adam 28 B/param (p, g, m, v read; p, m, v
    written), Polyak 12 B/param, plus every fp32 weight streamed once per GEMM that uses it."""
    n = module_params(w)
    alg = w["alg"]
    K = w["kw"].get("extra_feature_steps", {"sac": -1, "diffsrsac": 3}.get(alg, 1)) + 1  # class defaults
    feat = {"ctrlsac": ("phi", "mu", "theta"), "vlsac": ("encoder", "decoder", "f"), "spedersac": ("phi", "mu", "theta"),
            "diffsrsac": ("phi", "nablamu"), "sac": ()}[alg]
    feat_target = {"ctrlsac": "phi", "vlsac": "f", "spedersac": "phi"}.get(alg)
    feat_n = sum(n[m] for m in feat)
    critic_adam = 0 if alg == "diffsrsac" else n["critic"]
    opt = 28 * (K * feat_n + critic_adam + n["actor"]) + 12 * (K * n.get(feat_target, 0) + n["critic"] / 2)
    # weight streaming: forward + dgrad per feature step; critic step: feature net x2 + critic + target critic (+ critic
    # dgrad inside wgrad chain); actor step: feature net fwd + dgrad, critic fwd + dgrad, actor fwd + dgrad (+ a' fwd)
    used = {"ctrlsac": "phi", "vlsac": "f", "spedersac": "phi", "diffsrsac": "phi"}.get(alg)
    used_n = n.get(used, 0)
    stream = 4 * (K * 2 * feat_n + 2 * used_n + 2 * n["critic"] + 2 * used_n + 2 * n["critic"] + 3 * n["actor"])
    return dict(optimizer=opt, weights=stream, total=opt + stream)


# ---------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.proc, self.path = gpu_index, None, None

    def __enter__(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.path or not os.path.exists(self.path):
            return out
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def measure_tf32_peak():
    """Dense TF32 tensor-core peak the way MEASURED_PEAKS.json measures bf16: torch.matmul (cuBLAS) on 8192^3, best of
    10 with CUDA events.  Only a roofline denominator -- never on the measured path."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(8192, 8192, device="cuda")
        b = torch.randn(8192, 8192, device="cuda")
        c = torch.empty(8192, 8192, device="cuda")
        for _ in range(3):
            torch.matmul(a, b, out=c)
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2 * 8192 ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


# ---------------------------------------------------------------------------------------------------- CPU arm
def make_oracle(w, as_written=True):
    from oracle import rl_oracle as O
    kw = dict(w["kw"])
    init = O.init_state(w["alg"], w["S"], w["A"], kw, seed=0)
    extra = dict(as_written=as_written) if w["alg"] == "ctrlsac" else {}
    if w["alg"] == "vlsac":
        import torch
        extra["critic_noise"] = torch.randn(20, kw.get("feature_dim", 256), generator=torch.Generator().manual_seed(1234))
    agent = O.ORACLES[w["alg"]](w["S"], w["A"], init, discount=0.99, tau=0.005, **kw, **extra)
    ring = O.synthetic_ring(w["S"], w["A"], min(w["rows"], 200_000), seed=0)
    return agent, ring


def time_oracle(w, steps, warmup, as_written=True):
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    agent, ring = make_oracle(w, as_written)
    np.random.seed(1)
    torch.manual_seed(1)
    for _ in range(warmup):
        agent.train(ring, w["B"])
    t0 = time.perf_counter()
    for _ in range(steps):
        agent.train(ring, w["B"])
    dt = time.perf_counter() - t0
    return steps / dt, dt / steps * 1e3, torch.get_num_threads()


def run_reference(args, w, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port, as written) on host cores."""
    if rank != 0:
        return
    as_written = not w.get("sharded")  # the [B,B,D] broadcast needs 2.2 TB at B = 16384: matmul-restated there
    ups, ms, cores = time_oracle(w, args.steps, args.warmup, as_written=as_written)
    sample = f"{args.steps} full train() calls after {args.warmup} warm-up, reference arithmetic " + \
        ("as written ([B,B,D] broadcast logits)" if w["alg"] == "ctrlsac" and as_written else
         "with the logits restated as a matmul (SURVEY.md 8c)" if w["alg"] == "ctrlsac" else "as written")
    line = {
        "impl": "reference", "metric": "agent updates/sec", "value": ups, "unit": "updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, **{k: w[k] for k in ("alg", "S", "A", "B")}, **w["kw"]},
        "cpu_baseline": {"value": ups, "unit": "updates/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ups, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------- pixel arm
def drq_args(w):
    import types
    return types.SimpleNamespace(tau=0.01, update_every=1, critic_loss="mse", stddev_schedule="linear(1.0,0.1,500000)",
                                 stddev_clip=0.3, bn_dim=w["bn"], actor_hidden_dim=w["H"], critic_hidden_dim=w["H"],
                                 encoder_lr=1e-4, actor_lr=1e-4, critic_lr=1e-4)


class _Box:
    def __init__(self, shape):
        self.shape = shape


MULV_CFG = dict(aug=True, pre_aug=False, back_q2feat=True, tanh=True, both_q=False, q_activ="relu", q_loss="huber",
                q_up_n=1, l2_norm=0.0, c_targ_tau=0.01, up_every=1, stddev_schedule="linear(1.0,0.1,500000)",
                stddev_clip=0.3, lr=1e-4, vae_w=0.5, mse_w=1.0, c_noise=0.1)


class _PixelArm:
    """The two pixel agents behind one face: step(i) = one updating call through the public API with host batches."""

    def __init__(self, w, precision, seed):
        from rlrep_b200 import _lib
        self.w, self.lib_mod = w, _lib
        C_, A = w["C"], w["A"]
        if w["alg"] == "drqv2":
            from oracle import drq_oracle as D  # synthetic batch + deterministic initial weights (data, not arithmetic)
            from rlrep_b200.pixel import DrQv2
            self.agent = DrQv2(_Box((C_, 84, 84)), _Box((A,)), drq_args(w), precision=precision)
            self.agent.load_state_dict(D.init_state(C_, A, w["bn"], w["H"], seed=seed))
            self.step = lambda batch, i: self.agent.train_step(iter([batch]), step=i)
            self.prefix = "rlrep_drq"
            self.h2d = 2 * w["B"] * C_ * 84 * 84 + 4 * (4 * w["B"] + 3 * w["B"] * A + 2 * w["B"])
        else:
            from oracle import mulv_oracle as D
            from rlrep_b200.pixel import MuLVDrQv2
            self.agent = MuLVDrQv2((C_, 84, 84), (A,), dict(MULV_CFG, feat_dim=w["F"], hid_dim=w["H"]), precision=precision)
            self.agent.load_state_dict(D.init_state(C_, A, w["F"], w["H"], seed=seed))
            self.step = lambda batch, i: self.agent.update(iter([batch]), step=i)
            self.prefix = "rlrep_mulv"
            self.h2d = (2 * C_ + 3) * w["B"] * 84 * 84 + 4 * (4 * w["B"] + w["B"] * w["F"] + 3 * w["B"] * A +
                                                              3 * 20 * w["F"] + 2 * w["B"])
        self.D = D

    def fn(self, name):
        return getattr(self.agent.lib, f"{self.prefix}_{name}")

    def make_oracle(self):
        w, D = self.w, self.D
        if w["alg"] == "drqv2":
            o = D.OracleDrQv2(w["A"], D.init_state(w["C"], w["A"], w["bn"], w["H"], seed=0), update_every=1)
            return lambda b: o.train_step(b, 0)
        o = D.OracleMuLVDrQ(w["A"], D.init_state(w["C"], w["A"], w["F"], w["H"], seed=0), up_every=1)
        return lambda b: o.update(b, 0)


def run_drq(args, w, rank, world, local_rank):
    """Pixel update (plain DrQ-v2 or muLV-Rep DrQ-v2): one step = one updating call on a [B, 9, 84, 84] uint8 batch
    (synthetic frames)."""
    import torch
    import torch.distributed as dist
    from rlrep_b200 import _lib
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    C_, A, B = w["C"], w["A"], w["B"]
    arm = _PixelArm(w, args.precision, rank)
    agent, D = arm.agent, arm.D
    batches = [tuple(D.synthetic_pixel_batch(B, C_, 84, A, seed=100 * rank + i)) for i in range(4)]
    torch.manual_seed(1 + rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        arm.step(batches[i % 4], i)
    ms = C.c_float()
    barrier()
    with ClockSampler(local_rank) as clk:  # (1) device-timed on the batch already resident in HBM
        _lib.check(arm.fn("update_resident")(agent._h, args.steps, 1.0, C.byref(ms)))
        barrier()
    dev_ms, clocks = float(ms.value), clk.summary()
    barrier()
    t0 = time.perf_counter()  # (2) end to end: host batch -> pinned staging -> H2D -> update -> metrics D2H
    info = None
    for i in range(args.steps):
        info = arm.step(batches[i % 4], i)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    launches = agent.gpu_launches_last_update
    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = t.tolist()
    roofline, top, cpu = None, [], None
    if rank == 0:
        cap = 4096
        names, kms = (C.c_char_p * cap)(), (C.c_float * cap)()
        kby, kfl, n = (C.c_double * cap)(), (C.c_double * cap)(), C.c_int()
        agg = {}
        for _ in range(3):
            _lib.check(arm.fn("profile_update")(agent._h, 1.0, cap, names, kms, kby, kfl, C.byref(n)))
            for i in range(min(n.value, cap)):
                a = agg.setdefault(names[i].decode(), [0.0, 0, 0.0, 0.0])
                a[0] += kms[i]; a[1] += 1; a[2] += kby[i]; a[3] += kfl[i]
        total = sum(v[0] for v in agg.values())
        top = sorted(((k, v[0] / 3, v[1] // 3) for k, v in agg.items()), key=lambda x: -x[1])
        pk = ROOT / "MEASURED_PEAKS.json"
        hbm_peak = float(json.loads(pk.read_text()).get("hbm_gbs", 6650.0)) if pk.exists() else 6650.0
        tf32_peak = measure_tf32_peak()
        step_s = dev_ms / args.steps * 1e-3
        by, fl = sum(v[2] for v in agg.values()) / 3, sum(v[3] for v in agg.values()) / 3
        k0 = top[0][0]
        t0k = agg[k0][0] * 1e-3
        f_h, f_t = agg[k0][2] / t0k / 1e9 / hbm_peak, agg[k0][3] / t0k / 1e12 / tf32_peak
        roofline = {"kernel": k0, "bound": "hbm" if f_h >= f_t else "tensor",
                    "achieved": agg[k0][2] / t0k / 1e9 if f_h >= f_t else agg[k0][3] / t0k / 1e12,
                    "peak": hbm_peak if f_h >= f_t else tf32_peak, "unit": "GB/s" if f_h >= f_t else "TFLOP/s",
                    "frac": max(f_h, f_t), "traffic": None, "share_of_step": agg[k0][0] / total,
                    "launches_per_step": agg[k0][1] / 3, "tf32_peak_tflops": tf32_peak,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs / cuBLAS TF32 8192^3 measured in this run",
                    "note": "v1 lowers the convolutions onto explicit im2col + GEMM: the column matrices are ~10x the "
                            "algorithmic traffic of the convolutions (DESIGN.md section 8)",
                    "step": {"algorithmic_bytes_of_launches": by, "algorithmic_flops": fl,
                             "roofline_ms": max(by / (hbm_peak * 1e9), fl / (tf32_peak * 1e12)) * 1e3,
                             "frac": max(by / (hbm_peak * 1e9), fl / (tf32_peak * 1e12)) / step_s}}
        if world == 1 and not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            oracle_step = arm.make_oracle()
            ob = D.synthetic_pixel_batch(B, C_, 84, A, seed=0)
            oracle_step(ob)
            t1 = time.perf_counter()
            n_cpu = 20 if w["alg"] == "drqv2" else 8
            for _ in range(n_cpu):
                oracle_step(ob)
            dt = (time.perf_counter() - t1) / n_cpu
            cpu = {"value": 1.0 / dt, "unit": "updates/s", "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"{n_cpu} updates of the same workload after 1 warm-up ({dt * 1e3:.0f} ms each), reference "
                             f"arithmetic as written (grid_sample augmentation, F.conv2d, autograd, torch.optim.Adam)"}
        line = {"metric": "agent updates/sec", "value": world * args.steps / (dev_ms * 1e-3), "unit": "updates/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "tf32" if args.precision == "tf32" else "f32", "data": "synthetic",
                "config": {"workload": args.workload, "alg": w["alg"], "obs": [C_, 84, 84], "A": A, "B": B,
                           "feature_dim": w.get("bn", w.get("F")), "hidden_dim": w["H"], "parallelism": f"replicas x{world} (no collective)",
                           "l2": "no flush: the update streams ~3 GB of column matrices and activations (>> 126 MB L2)"},
                "e2e": {"value": world * args.steps / (e2e_ms * 1e-3), "unit": "updates/s", "ms_per_step": e2e_ms / args.steps,
                        "h2d_bytes_per_step": arm.h2d, "d2h_bytes_per_step": 32},
                "gpu_launches": launches * args.steps, "gpu_launches_per_step": launches, "clocks": clocks,
                "roofline": roofline, "cpu_baseline": cpu,
                "top_kernels_us_per_step": [[k, round(v * 1e3, 1), c] for k, v, c in top[:8]], "last_info": info}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_drq_reference(args, w, rank):
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    if w["alg"] == "drqv2":
        from oracle import drq_oracle as D
        o = D.OracleDrQv2(w["A"], D.init_state(w["C"], w["A"], w["bn"], w["H"], seed=0), update_every=1)
        oracle_step = lambda b: o.train_step(b, 0)
    else:
        from oracle import mulv_oracle as D
        o = D.OracleMuLVDrQ(w["A"], D.init_state(w["C"], w["A"], w["F"], w["H"], seed=0), up_every=1)
        oracle_step = lambda b: o.update(b, 0)
    ob = D.synthetic_pixel_batch(w["B"], w["C"], 84, w["A"], seed=0)
    torch.manual_seed(1)
    for _ in range(args.warmup):
        oracle_step(ob)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(ob)
    dt = (time.perf_counter() - t0) / args.steps
    emit({"impl": "reference", "metric": "agent updates/sec", "value": 1.0 / dt, "unit": "updates/s", "n_gpus": args.gpus,
          "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": args.workload, "alg": w["alg"]},
          "cpu_baseline": {"value": 1.0 / dt, "unit": "updates/s", "cores": torch.get_num_threads(), "kind": "port",
                           "sample": f"{args.steps} updates after {args.warmup} warm-up"},
          "e2e": {"value": 1.0 / dt, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


# ---------------------------------------------------------------------------------------------------- GPU arm
def run_ours(args, w, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from rlrep_b200 import ReplayBuffer, _lib
    from rlrep_b200.agents import AGENTS
    from oracle import rl_oracle as O  # synthetic data generator + deterministic initial weights only

    torch.cuda.set_device(local_rank)
    sharded = bool(w.get("sharded"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    elif sharded:  # the sharded agent at N = 1 still goes through a (single-rank) process group and communicator
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29517")
        dist.init_process_group("gloo", rank=0, world_size=1)
    S, A, B, kw = w["S"], w["A"], w["B"], w["kw"]
    if sharded:
        from rlrep_b200.agents import ShardedCTRLSACAgent
        # one logical agent over all ranks: identical weights, identical seeds, the global batch split by rows
        agent = ShardedCTRLSACAgent(S, A, Space(A), discount=0.99, tau=0.005, precision=args.precision, **kw)
        seed = 0
    else:
        agent = AGENTS[w["alg"]](S, A, Space(A), discount=0.99, tau=0.005, precision=args.precision, **kw)
        seed = rank  # independent replicas: own weights, own data, own seeds
    agent.load_state_dict(O.init_state(w["alg"], S, A, kw, seed=seed))
    ring = O.synthetic_ring(S, A, w["rows"], seed=seed)
    buf = ReplayBuffer(S, A, max_size=w["rows"])
    buf.load(ring.state, ring.action, ring.next_state, ring.reward, ring.done)
    del ring
    np.random.seed(1 + seed)
    torch.manual_seed(1 + seed)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up through the public API (eager call, graph capture, replays)
    for _ in range(max(args.warmup, 3)):
        agent.train(buf, B)
    h = agent._h

    # ---- (1) device-timed, inputs resident: all indices / noise uploaded before the timed region
    draws = [agent._draw(buf, B) for _ in range(args.steps)]
    idx_all = np.ascontiguousarray(np.concatenate([d[0] for d in draws]), dtype=np.int64)
    eps_all = np.ascontiguousarray(np.concatenate([d[1] for d in draws]), dtype=np.float32)
    ms = C.c_float()
    barrier()
    with ClockSampler(local_rank) as clk:
        _lib.check(h.lib.rlrep_agent_train_resident(h.h, buf._h, idx_all.ctypes.data, eps_all.ctypes.data, args.steps,
                                                    C.byref(ms)))
        barrier()
    dev_ms = float(ms.value)
    clocks = clk.summary()

    # ---- (2) end to end through agent.train(): host RNG draws, H2D of inputs, D2H of metrics, every step
    barrier()
    t0 = time.perf_counter()
    info = None
    for _ in range(args.steps):
        info = agent.train(buf, B)
    barrier()
    e2e_s = time.perf_counter() - t0
    launches = agent.gpu_launches_last_train

    if world > 1:
        t = torch.tensor([dev_ms, e2e_s * 1e3], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = t.tolist()
    else:
        e2e_ms = e2e_s * 1e3

    # ---- (3) per-kernel profile (eager, one stream, an event behind every launch).  Every launch carries its
    # algorithmic bytes / flops (operands read once, results written once), so each kernel gets a roofline fraction.
    roofline, top = None, []
    if rank != 0 and sharded:  # the sharded update contains collectives: every rank has to take part in the profiled calls
        for _ in range(3):
            i1, e1 = agent._draw(buf, B)
            i1 = np.ascontiguousarray(i1, dtype=np.int64)
            e1 = np.ascontiguousarray(e1, dtype=np.float32)
            _lib.check(h.lib.rlrep_agent_profile_train(h.h, buf._h, i1.ctypes.data, e1.ctypes.data, 0, None, None, None,
                                                       None, C.byref(C.c_int())))
    if rank == 0:
        cap = 8192
        names = (C.c_char_p * cap)()
        kms = (C.c_float * cap)()
        kby = (C.c_double * cap)()
        kfl = (C.c_double * cap)()
        n = C.c_int()
        agg = {}
        reps = 3
        for _ in range(reps):
            i1, e1 = agent._draw(buf, B)
            i1 = np.ascontiguousarray(i1, dtype=np.int64)
            e1 = np.ascontiguousarray(e1, dtype=np.float32)
            _lib.check(h.lib.rlrep_agent_profile_train(h.h, buf._h, i1.ctypes.data, e1.ctypes.data, cap, names, kms, kby,
                                                       kfl, C.byref(n)))
            for i in range(min(n.value, cap)):
                a = agg.setdefault(names[i].decode(), [0.0, 0, 0.0, 0.0])
                a[0] += kms[i]
                a[1] += 1
                a[2] += kby[i]
                a[3] += kfl[i]
        total = sum(v[0] for v in agg.values())
        top = sorted(((k, v[0] / reps, v[1] // reps) for k, v in agg.items()), key=lambda x: -x[1])
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "MEASURED_PEAKS.json hbm_gbs" if pk.exists() else "fallback 6650 GB/s (B200_PROFILING.md)"
        tf32_peak = measure_tf32_peak()  # cuBLAS TF32 8192^3, measured here the way MEASURED_PEAKS.json measures bf16
        traffic = {}
        tp = ROOT / "profiles" / "ncu_traffic.json"  # dram__bytes_read+write per launch from the committed ncu captures
        if tp.exists():
            traffic = json.loads(tp.read_text()).get(args.workload, {})

        def kernel_roofline(name):
            ms_sum, count, by, fl = agg[name]
            t = ms_sum * 1e-3
            gbs, tfs = by / t / 1e9, fl / t / 1e12
            f_hbm, f_tc = gbs / hbm_peak, tfs / tf32_peak
            bound = "hbm" if f_hbm >= f_tc else "tensor"
            return {"kernel": name, "bound": bound, "achieved": gbs if bound == "hbm" else tfs,
                    "peak": hbm_peak if bound == "hbm" else tf32_peak, "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
                    "frac": max(f_hbm, f_tc), "traffic": traffic.get(name), "launches_per_step": count / reps,
                    "avg_launch_us": ms_sum / count * 1e3, "algorithmic_bytes_per_launch": by / count,
                    "algorithmic_flops_per_launch": fl / count, "achieved_gbs": gbs, "achieved_tflops": tfs,
                    "share_of_step": ms_sum / total}

        ranked = [k for k, _, _ in top if agg[k][2] > 0 or agg[k][3] > 0]
        if ranked:
            alg_bytes = algorithmic_bytes(w)
            step_s = dev_ms / args.steps * 1e-3
            step_flops = sum(v[3] for v in agg.values()) / reps
            t_hbm, t_tc = alg_bytes["total"] / (hbm_peak * 1e9), step_flops / (tf32_peak * 1e12)
            roofline = kernel_roofline(ranked[0])  # the kernel with the largest share of the step
            roofline["peak_source"] = hbm_src if roofline["bound"] == "hbm" else "cuBLAS TF32 8192^3 measured in this run"
            roofline["tf32_peak_tflops"] = tf32_peak
            roofline["kernels"] = [kernel_roofline(k) for k in ranked[1:6]]
            roofline["step"] = {"algorithmic_bytes": alg_bytes["total"], "algorithmic_flops": step_flops,
                                "bound": "hbm" if t_hbm >= t_tc else "tensor",
                                "roofline_ms": max(t_hbm, t_tc) * 1e3, "frac": max(t_hbm, t_tc) / step_s,
                                "achieved_gbs": alg_bytes["total"] / step_s / 1e9,
                                "achieved_tflops": step_flops / step_s / 1e12}

    # ---- (4) CPU baseline: the oracle port on this box's host cores (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not sharded:
        n_cpu = {"ctrlsac": 6, "vlsac": 30, "spedersac": 60}.get(w["alg"], 200)  # ~10-30 s of CPU work
        ups, ms_cpu, cores = time_oracle(w, n_cpu, 1, as_written=True)
        cpu = {"value": ups, "unit": "updates/s", "cores": cores, "kind": "port",
               "sample": f"{n_cpu} full train() calls of the same workload after 1 warm-up ({ms_cpu:.0f} ms each), "
                         f"reference arithmetic as written" + (" (broadcast logits)" if w["alg"] == "ctrlsac" else "")}

    if rank == 0:
        ni, ne = h.n_idx, h.n_eps
        n_agents = 1 if sharded else world  # a sharded run is ONE agent's update, however many GPUs compute it
        line = {
            "metric": "agent updates/sec", "value": n_agents * args.steps / (dev_ms * 1e-3), "unit": "updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
            "dtype": "tf32" if args.precision == "tf32" else "f32", "data": "synthetic",
            "config": {"workload": args.workload, **{k: w[k] for k in ("alg", "S", "A", "B")}, **kw,
                       "ring_rows": w["rows"],
                       "parallelism": (f"one agent, batch sharded by rows over {world} GPU(s): {B // world} rows/GPU, NCCL "
                                       f"all-gather of mu(s'), reduce-scatter of d mu, all-reduce of gradients"
                                       if sharded else f"replicas x{world} (no collective)"),
                       "l2": ("no flush: per-update working set (params+grads+Adam moments+targets ~200 MB) exceeds the 126 MB L2"
                              if w["alg"] == "ctrlsac" else "no flush between updates (working set below L2: see DESIGN.md)")},
            "e2e": {"value": n_agents * args.steps / (e2e_ms * 1e-3), "unit": "updates/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": ni * 8 + ne * 4, "d2h_bytes_per_step": 32 * 4},
            "gpu_launches": launches * args.steps,
            "gpu_launches_per_step": launches,
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "top_kernels_us_per_step": [[k, round(v * 1e3, 1), c] for k, v, c in top[:8]],
            "last_info": info,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    # stdout carries the JSON line and nothing else: native libraries (NCCL's version banner ...) write to fd 1 directly,
    # so fd 1 is pointed at stderr for the duration of the run and the JSON line goes to the saved real stdout.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ctrlsac_hc_b256", choices=list(WORKLOADS))
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if w["alg"] in ("drqv2", "mulvdrq"):
        if args.impl == "reference":
            args.steps, args.warmup = min(args.steps, 20 if w["alg"] == "drqv2" else 8), min(args.warmup, 2)
            run_drq_reference(args, w, rank)
        else:
            run_drq(args, w, rank, world, local_rank)
        return
    if args.impl == "reference":
        if w.get("sharded"):
            args.steps, args.warmup = min(args.steps, 2), 0  # ~17 TFLOP per update on the host cores: a bounded sample
        elif args.steps > 40 and w["alg"] == "ctrlsac":
            args.steps = 40  # bounded sample: ~0.5-2 s of CPU work per update
        run_reference(args, w, rank, world)
        return
    run_ours(args, w, rank, world, local_rank)


if __name__ == "__main__":
    main()
