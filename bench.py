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

WORKLOADS = {
    # BASELINE.json configs[1]
    "ctrlsac_hc_b256": dict(alg="ctrlsac", S=17, A=6, B=256, rows=1_000_000,
                            kw=dict(hidden_dim=1024, feature_dim=2048, extra_feature_steps=3)),
    # BASELINE.json configs[0] (the reference's own CPU-runnable case)
    "sac_hc_b256": dict(alg="sac", S=17, A=6, B=256, rows=1_000_000, kw=dict(hidden_dim=256)),
}


class Space:
    def __init__(self, A):
        self.low, self.high, self.shape = -np.ones(A, np.float32), np.ones(A, np.float32), (A,)


def algorithmic_bytes(w):
    """HBM bytes one update must move (SURVEY.md 8d / BASELINE.md): Adam 28 B/param (p, g, m, v read; p, m, v
    written), Polyak 12 B/param, plus every fp32 weight streamed once per GEMM that uses it."""
    S, A, kw = w["S"], w["A"], w["kw"]
    if w["alg"] == "ctrlsac":
        H, D, K = kw["hidden_dim"], kw["feature_dim"], kw["extra_feature_steps"] + 1
        phi = (S + A) * H + H + H * H + H + H * D + D
        mu = S * H + H + H * H + H + H * D + D
        theta = D + 1
        critic = 2 * (D * H + H + H + 1)
        actor = S * 256 + 256 + 256 * 256 + 256 + 256 * 2 * A + 2 * A
        feat = phi + mu + theta
        opt = 28 * (K * feat + critic + actor) + 12 * (K * phi + critic / 2)
        # weight streaming: fwd + dgrad per feature step (phi, mu), critic step: phi x2, critic/target; actor: phi fwd
        # + dgrad, critic fwd + dgrad, actor fwd + dgrad
        stream = 4 * (K * 2 * (phi + mu) + 2 * phi + 2 * critic + 2 * phi + 2 * critic + 3 * actor)
        return dict(optimizer=opt, weights=stream, total=opt + stream, adam_feature_launch=28 * feat + 12 * phi)
    if w["alg"] == "sac":
        H = kw["hidden_dim"]
        critic = 2 * ((S + A) * H + H + H * H + H + H + 1)
        actor = S * H + H + H * H + H + H * 2 * A + 2 * A
        opt = 28 * (critic + actor) + 12 * critic / 2
        stream = 4 * (4 * critic + 3 * actor)
        return dict(optimizer=opt, weights=stream, total=opt + stream, adam_feature_launch=28 * critic + 12 * critic)
    raise ValueError(w["alg"])


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


# ---------------------------------------------------------------------------------------------------- CPU arm
def make_oracle(w, as_written=True):
    from oracle import rl_oracle as O
    kw = dict(w["kw"])
    init = O.init_state(w["alg"], w["S"], w["A"], kw, seed=0)
    extra = dict(as_written=as_written) if w["alg"] == "ctrlsac" else {}
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
    ups, ms, cores = time_oracle(w, args.steps, args.warmup, as_written=True)
    sample = f"{args.steps} full train() calls after {args.warmup} warm-up, as-written [B,B,D] broadcast logits"
    line = {
        "impl": "reference", "metric": "agent updates/sec", "value": ups, "unit": "updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, **{k: w[k] for k in ("alg", "S", "A", "B")}, **w["kw"]},
        "cpu_baseline": {"value": ups, "unit": "updates/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ups, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------- GPU arm
def run_ours(args, w, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from rlrep_b200 import ReplayBuffer, _lib
    from rlrep_b200.agents import AGENTS
    from oracle import rl_oracle as O  # synthetic data generator + deterministic initial weights only

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    S, A, B, kw = w["S"], w["A"], w["B"], w["kw"]
    agent = AGENTS[w["alg"]](S, A, Space(A), discount=0.99, tau=0.005, precision=args.precision, **kw)
    agent.load_state_dict(O.init_state(w["alg"], S, A, kw, seed=rank))  # independent seeds per replica
    ring = O.synthetic_ring(S, A, w["rows"], seed=rank)
    buf = ReplayBuffer(S, A, max_size=w["rows"])
    buf.load(ring.state, ring.action, ring.next_state, ring.reward, ring.done)
    del ring
    np.random.seed(1 + rank)
    torch.manual_seed(1 + rank)

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

    # ---- (3) per-kernel profile (eager, event behind every launch) -> dominant kernel for the roofline
    roofline, top = None, []
    if rank == 0:
        cap = 4096
        names = (C.c_char_p * cap)()
        kms = (C.c_float * cap)()
        n = C.c_int()
        agg = {}
        reps = 3
        for _ in range(reps):
            i1, e1 = agent._draw(buf, B)
            i1 = np.ascontiguousarray(i1, dtype=np.int64)
            e1 = np.ascontiguousarray(e1, dtype=np.float32)
            _lib.check(h.lib.rlrep_agent_profile_train(h.h, buf._h, i1.ctypes.data, e1.ctypes.data, cap, names, kms,
                                                       C.byref(n)))
            for i in range(min(n.value, cap)):
                a = agg.setdefault(names[i].decode(), [0.0, 0])
                a[0] += kms[i]
                a[1] += 1
        total = sum(v[0] for v in agg.values())
        top = sorted(((k, v[0] / reps, v[1] // reps) for k, v in agg.items()), key=lambda x: -x[1])
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        alg_bytes = algorithmic_bytes(w)
        # dominant bandwidth-bound kernel: the fused Adam+Polyak launch over the feature group
        adam = agg.get("adam_polyak")
        if adam:
            n_launch = adam[1] / reps
            per_launch_ms = adam[0] / adam[1]
            # bytes of an average adam_polyak launch: optimiser traffic of the update / launches per update
            per_launch_bytes = alg_bytes["optimizer"] / n_launch
            ach = per_launch_bytes / (per_launch_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": "adam_polyak_kernel", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                        "frac": ach / hbm_peak, "traffic": None, "launches_per_step": n_launch,
                        "avg_launch_us": per_launch_ms * 1e3,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if pk.exists() else "fallback 6650",
                        "share_of_step": adam[0] / total,
                        "step": {"algorithmic_bytes": alg_bytes["total"],
                                 "achieved": alg_bytes["total"] / (dev_ms / args.steps * 1e-3) / 1e9,
                                 "frac": alg_bytes["total"] / (dev_ms / args.steps * 1e-3) / 1e9 / hbm_peak}}

    # ---- (4) CPU baseline: the oracle port on this box's host cores (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_cpu = 6 if w["alg"] == "ctrlsac" else 200
        ups, ms_cpu, cores = time_oracle(w, n_cpu, 1, as_written=True)
        cpu = {"value": ups, "unit": "updates/s", "cores": cores, "kind": "port",
               "sample": f"{n_cpu} full train() calls of the same workload after 1 warm-up ({ms_cpu:.0f} ms each), "
                         f"reference arithmetic as written (broadcast logits)"}

    if rank == 0:
        ni, ne = h.n_idx, h.n_eps
        line = {
            "metric": "agent updates/sec", "value": world * args.steps / (dev_ms * 1e-3), "unit": "updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32" if args.precision == "tf32" else "f32", "data": "synthetic",
            "config": {"workload": args.workload, **{k: w[k] for k in ("alg", "S", "A", "B")}, **kw,
                       "ring_rows": w["rows"], "parallelism": f"replicas x{world} (no collective)",
                       "l2": "no flush: per-update working set (params+grads+Adam moments+targets ~200 MB) exceeds the 126 MB L2"},
            "e2e": {"value": world * args.steps / (e2e_ms * 1e-3), "unit": "updates/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": ni * 8 + ne * 4, "d2h_bytes_per_step": 32 * 4},
            "gpu_launches": launches * args.steps,
            "gpu_launches_per_step": launches,
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "top_kernels_us_per_step": [[k, round(v * 1e3, 1), c] for k, v, c in top[:8]],
            "last_info": info,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
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
    if args.impl == "reference":
        if args.steps > 40 and w["alg"] == "ctrlsac":
            args.steps = 40  # bounded sample: ~1-2 s of CPU work per update
        run_reference(args, w, rank, world)
        return
    run_ours(args, w, rank, world, local_rank)


if __name__ == "__main__":
    main()
