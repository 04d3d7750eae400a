mkdir -p gpurun_out/r02
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02/bench_default_n8.json 2> gpurun_out/r02/bench_default_n8.err; tail -3 gpurun_out/r02/bench_default_n8.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02/bench_default_n8.json"))
print("default N=8:", round(d["value"], 1), "upd/s", round(d["ms_per_step"], 4), "ms; e2e", round(d["e2e"]["value"], 1))
print("   sharded", json.dumps(d.get("sharded"))[:1100]); print("   parity", d.get("sharded_parity"))
PY
for sl in 1 2; do
RLREP_DP_SLICES=$sl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2964$sl bench.py --gpus 8 --workload ctrlsac_b16384_sharded --steps 10 --warmup 3 --repeats 3 --no-cpu-baseline --no-alt-precision > gpurun_out/r02/bench_sharded_n8_sl$sl.json 2> gpurun_out/r02/bench_sharded_n8_sl$sl.err; tail -2 gpurun_out/r02/bench_sharded_n8_sl$sl.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02/bench_sharded_n8_sl$sl.json"))
print("sharded N=8 slices=$sl:", round(d["value"], 2), "upd/s", round(d["ms_per_step"], 3), "ms; e2e", round(d["e2e"]["value"], 2))
print("   top", d["top_kernels_us_per_step"][:6])
PY
done
