mkdir -p gpurun_out/r02
timeout 300 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -3
for cfg in "fwd 0 0" "bwd 0 0"; do timeout 120 python tests/gpu_chain_probe.py $cfg 2>&1 | tee -a gpurun_out/r02/chain_probe4.log | cut -c1-380; done
for w in fwd bwd; do for bn in 32 64 128; do for sp in 1 2 4; do
  timeout 60 python tests/gpu_chain_probe.py $w $bn $sp 2>&1 | head -1 | tee -a gpurun_out/r02/chain_probe4.log
done; done; done
timeout 300 python bench.py --steps 50 --warmup 5 --repeats 3 --no-cpu-baseline --no-sharded --no-alt-precision > gpurun_out/r02/bench_chain_v4.json 2> gpurun_out/r02/bench_chain_v4.err; tail -3 gpurun_out/r02/bench_chain_v4.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02/bench_chain_v4.json"))
print("chain v4:", round(d["value"], 1), "upd/s", round(d["ms_per_step"], 4), "ms; e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches_per_step"])
print("   top", d["top_kernels_us_per_step"])
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "ctrlsac" 2>&1 | tail -4
timeout 200 python tests/gpu_timeline.py ctrlsac_hc_b256 > gpurun_out/r02/timeline_chain4.csv 2> /dev/null
