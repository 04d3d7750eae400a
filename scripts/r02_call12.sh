mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_sharded.py -q -s 2>&1 | tail -6
for sl in 1 4; do
RLREP_DP_SLICES=$sl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$sl bench.py --gpus 2 --workload ctrlsac_b16384_sharded --steps 5 --warmup 3 --repeats 3 --no-cpu-baseline --no-alt-precision > gpurun_out/r02/bench_sharded_n2_sl$sl.json 2> gpurun_out/r02/bench_sharded_n2_sl$sl.err; tail -2 gpurun_out/r02/bench_sharded_n2_sl$sl.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02/bench_sharded_n2_sl$sl.json"))
print("sharded N=2 slices=$sl:", round(d["value"], 2), "upd/s", round(d["ms_per_step"], 3), "ms; e2e", round(d["e2e"]["value"], 2))
print("   top", d["top_kernels_us_per_step"][:6])
PY
done
