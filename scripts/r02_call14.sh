mkdir -p gpurun_out/r02
for p in default 1; do
if [ "$p" = "default" ]; then unset RLREP_TC_PERSIST; else export RLREP_TC_PERSIST=$p; fi
timeout 600 python bench.py --workload ctrlsac_b16384_sharded --steps 5 --warmup 3 --repeats 3 --no-cpu-baseline --no-alt-precision > gpurun_out/r02/bench_sharded_n1_persist_$p.json 2> gpurun_out/r02/bench_sharded_n1_persist_$p.err; tail -2 gpurun_out/r02/bench_sharded_n1_persist_$p.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02/bench_sharded_n1_persist_$p.json"))
print("sharded N=1 persist=$p:", round(d["value"], 2), "upd/s", round(d["ms_per_step"], 3), "ms")
print("   top", d["top_kernels_us_per_step"][:6]); r = d["roofline"]; print("   ", r["kernel"], r["bound"], round(r["frac"], 3))
PY
done
unset RLREP_TC_PERSIST
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "ragged or checkpoint or batch_size" 2>&1 | tail -4
