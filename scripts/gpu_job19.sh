timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5
cat > /tmp/drq_bench.py <<'PY'
import sys, types, time
sys.path.insert(0, '.')
import numpy as np, torch
from oracle import drq_oracle as D
from rlrep_b200.pixel import DrQv2
class Box:
    def __init__(self, shape): self.shape = shape
C,A,bn,H,B = 9,4,50,1024,256
args = types.SimpleNamespace(tau=0.01, update_every=1, critic_loss="mse", stddev_schedule="linear(1.0,0.1,500000)", stddev_clip=0.3, bn_dim=bn, actor_hidden_dim=H, critic_hidden_dim=H, encoder_lr=1e-4, actor_lr=1e-4, critic_lr=1e-4)
init = D.init_state(C, A, bn, H, seed=0)
agent = DrQv2(Box((C,84,84)), Box((A,)), args, precision="tf32"); agent.load_state_dict(init)
b = tuple(D.synthetic_pixel_batch(B, C, 84, A, seed=0))
for _ in range(3): agent.train_step(iter([b]), 0)
torch.cuda.synchronize(); t0=time.perf_counter()
n=20
for _ in range(n): agent.train_step(iter([b]), 0)
torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/n
print(f"DrQv2 B=256 tf32 end-to-end (H2D of 2x[256,9,84,84] uint8 inside): {dt*1e3:.2f} ms/update = {1/dt:.1f} updates/s; {agent.gpu_launches_last_update} launches")
oracle = D.OracleDrQv2(A, init, update_every=1)
torch.set_num_threads(__import__('os').cpu_count())
bb = D.synthetic_pixel_batch(B, C, 84, A, seed=0)
oracle.train_step(bb, 0); t0=time.perf_counter(); oracle.train_step(bb, 0); oracle.train_step(bb, 0)
print(f"reference arithmetic on {torch.get_num_threads()} host threads: {(time.perf_counter()-t0)/2*1e3:.0f} ms/update")
PY
python /tmp/drq_bench.py 2>&1 | tail -3
