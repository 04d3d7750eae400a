mkdir -p gpurun_out/r02
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02/bench_default_n2.json 2> gpurun_out/r02/bench_default_n2.err; tail -5 gpurun_out/r02/bench_default_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02/bench_default_n2.json"))
print("default N=2:", round(d["value"], 1), "upd/s", round(d["ms_per_step"], 4), "ms; e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches_per_step"], d["config"]["parallelism"])
print("   sharded", json.dumps(d.get("sharded"))[:900]); print("   parity", d.get("sharded_parity"))
PY
timeout 600 python -m pytest tests/test_gpu_sharded.py -q -s 2>&1 | tail -6
