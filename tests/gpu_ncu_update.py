"""A few train() calls of one bench workload for ncu (launch list or `--set full` on one kernel):
    python tests/gpu_ncu_update.py [workload] [n_calls] [eager]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import bench
from oracle import rl_oracle as O
from rlrep_b200 import ReplayBuffer
from rlrep_b200.agents import AGENTS

name = sys.argv[1] if len(sys.argv) > 1 else "ctrlsac_hc_b256"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eager = len(sys.argv) > 3 and sys.argv[3] == "eager"
w = bench.WORKLOADS[name]
S, A, B, kw = w["S"], w["A"], w["B"], w["kw"]
agent = AGENTS[w["alg"]](S, A, bench.Space(A), discount=0.99, tau=0.005, use_cuda_graph=not eager, **kw)
agent.load_state_dict(O.init_state(w["alg"], S, A, kw, seed=0))
rows = min(w["rows"], 100_000)
ring = O.synthetic_ring(S, A, rows, seed=0)
buf = ReplayBuffer(S, A, max_size=rows)
buf.load(ring.state, ring.action, ring.next_state, ring.reward, ring.done)
np.random.seed(1)
torch.manual_seed(1)
for _ in range(n):
    agent.train(buf, B)
torch.cuda.synchronize()
