"""Kernel timeline of one graph-replayed train() (CUPTI through torch.profiler): start / duration / stream of every
kernel, so overlap between the graph's branches and the gaps between dependent kernels can be read directly.
    python tests/gpu_timeline.py [workload] > gpurun_out/timeline.csv"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile
import bench
from oracle import rl_oracle as O
from rlrep_b200 import ReplayBuffer
from rlrep_b200.agents import AGENTS

name = sys.argv[1] if len(sys.argv) > 1 else "ctrlsac_hc_b256"
w = bench.WORKLOADS[name]
S, A, B, kw = w["S"], w["A"], w["B"], w["kw"]
agent = AGENTS[w["alg"]](S, A, bench.Space(A), discount=0.99, tau=0.005, **kw)
agent.load_state_dict(O.init_state(w["alg"], S, A, kw, seed=0))
rows = min(w["rows"], 100_000)
ring = O.synthetic_ring(S, A, rows, seed=0)
buf = ReplayBuffer(S, A, max_size=rows)
buf.load(ring.state, ring.action, ring.next_state, ring.reward, ring.done)
np.random.seed(1)
torch.manual_seed(1)
for _ in range(6):
    agent.train(buf, B)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        agent.train(buf, B)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
print("start_us,dur_us,stream,name")
for e in ev:
    print(f"{e.time_range.start - t0:.2f},{e.time_range.end - e.time_range.start:.2f},{getattr(e, 'stream', -1)},{e.name[:70]}")
