mkdir -p gpurun_out/r02
timeout 300 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "vlsac or ctrlsac_small" 2>&1 | tail -4
for c in 1 0; do
RLREP_CHAIN=$c timeout 300 python bench.py --workload vlsac_hum_b1024 --steps 30 --warmup 5 --repeats 3 --no-cpu-baseline --no-alt-precision > gpurun_out/r02/bench_vlsac_chain$c.json 2> gpurun_out/r02/bench_vlsac_chain$c.err; tail -2 gpurun_out/r02/bench_vlsac_chain$c.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02/bench_vlsac_chain$c.json"))
print("vlsac chain=$c:", round(d["value"], 1), "upd/s", round(d["ms_per_step"], 4), "ms; e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches_per_step"])
print("   top", d["top_kernels_us_per_step"])
PY
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02/bench_default_n1.json 2> gpurun_out/r02/bench_default_n1.err; tail -5 gpurun_out/r02/bench_default_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02/bench_default_n1.json"))
print("default:", round(d["value"], 1), "upd/s", round(d["ms_per_step"], 4), "ms; e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches_per_step"])
print("   top", d["top_kernels_us_per_step"])
r = d["roofline"]; print("   roofline", r["kernel"], r["bound"], round(r["frac"], 3), "share", round(r["share_of_step"], 3), "step", r["step"])
print("   fp32", d.get("fp32")); print("   cpu", d.get("cpu_baseline")); print("   sharded", json.dumps(d.get("sharded"))[:600]); print("   parity", d.get("sharded_parity"))
PY
for wl in mulvdrq_pixels_b256 drqv2_pixels_b256 ldiffsr_pixels_b256; do
timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --repeats 2 --no-cpu-baseline > gpurun_out/r02/bench_$wl.json 2> gpurun_out/r02/bench_$wl.err; tail -3 gpurun_out/r02/bench_$wl.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02/bench_$wl.json"))
    print("$wl:", round(d["value"], 2), "upd/s", round(d["ms_per_step"], 3), "ms; e2e", round(d["e2e"]["value"], 2), "launches", d["gpu_launches_per_step"], "| fp32:", d.get("fp32"))
    print("   top", d["top_kernels_us_per_step"][:5]); print("   step", d["roofline"]["step"] if d.get("roofline") else None)
except Exception as e:
    print("$wl failed", e)
PY
done
