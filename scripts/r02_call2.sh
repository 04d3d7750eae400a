# round 2, GPU call 2: GEMM chain kernel -- unit tests, CTRL-SAC parity through it, bench against the per-GEMM path
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -s > gpurun_out/r02/pytest_chain.log 2>&1; tail -25 gpurun_out/r02/pytest_chain.log
RLREP_CHAIN_VERBOSE=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -s -k "ctrlsac" > gpurun_out/r02/pytest_ctrl_chain.log 2>&1; tail -8 gpurun_out/r02/pytest_ctrl_chain.log
for c in 1 0; do
  RLREP_CHAIN=$c timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r02/bench_chain$c.json 2> gpurun_out/r02/bench_chain$c.err
done
python - <<'PY'
import json
for f in ("chain1", "chain0"):
    try:
        d = json.load(open(f"gpurun_out/r02/bench_{f}.json"))
        print(f, round(d["value"], 1), "upd/s", round(d["ms_per_step"], 4), "ms; e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches_per_step"])
        print("   top", d["top_kernels_us_per_step"])
    except Exception as e:
        print(f, "failed", e)
PY
timeout 200 python tests/gpu_timeline.py ctrlsac_hc_b256 > gpurun_out/r02/timeline_chain.csv 2> gpurun_out/r02/timeline_chain.err
timeout 1200 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_chain.py > gpurun_out/r02/pytest_all2.log 2>&1; tail -15 gpurun_out/r02/pytest_all2.log
