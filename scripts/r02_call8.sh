mkdir -p gpurun_out/r02
timeout 300 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -3
for cfg in "fwd 0 0" "bwd 0 0"; do RLREP_CHAIN_VERBOSE=1 timeout 120 python tests/gpu_chain_probe.py $cfg 2>&1 | tee -a gpurun_out/r02/chain_probe6.log | cut -c1-380; done
for w in fwd bwd; do for bn in 64 128; do for sp in 1 2; do
  timeout 60 python tests/gpu_chain_probe.py $w $bn $sp 2>&1 | head -1 | tee -a gpurun_out/r02/chain_probe6.log
done; done; done
RLREP_CHAIN_FILLERS=0 timeout 60 python tests/gpu_chain_probe.py bwd 0 0 2>&1 | head -1
timeout 300 python bench.py --steps 50 --warmup 5 --repeats 3 --no-cpu-baseline --no-sharded --no-alt-precision > gpurun_out/r02/bench_chain_v6.json 2> gpurun_out/r02/bench_chain_v6.err; tail -3 gpurun_out/r02/bench_chain_v6.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02/bench_chain_v6.json"))
print("chain v6:", round(d["value"], 1), "upd/s", round(d["ms_per_step"], 4), "ms; e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches_per_step"])
print("   top", d["top_kernels_us_per_step"])
PY
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ldiffsr.py -q -x -k "ctrlsac or ldiffsr" 2>&1 | tail -4
timeout 200 python tests/gpu_timeline.py ctrlsac_hc_b256 > gpurun_out/r02/timeline_chain6.csv 2> /dev/null
timeout 400 python bench.py --workload mulvdrq_population --agents-per-gpu 8 --steps 5 --warmup 3 --repeats 2 > gpurun_out/r02/bench_population_n1.json 2> gpurun_out/r02/bench_population_n1.err; tail -3 gpurun_out/r02/bench_population_n1.err; head -c 1500 gpurun_out/r02/bench_population_n1.json
