mkdir -p gpurun_out/r02
timeout 300 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -3
for cfg in "fwd 0 0" "bwd 0 0"; do timeout 120 python tests/gpu_chain_probe.py $cfg 2>&1 | tee -a gpurun_out/r02/chain_probe5.log | cut -c1-380; done
for fa in 0 1; do
RLREP_FUSE_ADAM=$fa timeout 300 python bench.py --steps 50 --warmup 5 --repeats 3 --no-cpu-baseline --no-sharded --no-alt-precision > gpurun_out/r02/bench_chain_v5_fa$fa.json 2> gpurun_out/r02/bench_chain_v5_fa$fa.err; tail -3 gpurun_out/r02/bench_chain_v5_fa$fa.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02/bench_chain_v5_fa$fa.json"))
print("chain v5 fuse_adam=$fa:", round(d["value"], 1), "upd/s", round(d["ms_per_step"], 4), "ms; e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches_per_step"])
print("   top", d["top_kernels_us_per_step"])
PY
done
RLREP_FUSE_ADAM=0 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "ctrlsac" 2>&1 | tail -4
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r02/pytest_all3.log 2>&1; tail -12 gpurun_out/r02/pytest_all3.log
timeout 200 python tests/gpu_timeline.py ctrlsac_hc_b256 > gpurun_out/r02/timeline_chain5.csv 2> /dev/null
