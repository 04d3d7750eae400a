# round 2, GPU call 1: baseline state of main + the two round-1 drafts (latent Diff-SR agent, PDL GEMM chain)
mkdir -p gpurun_out/r02
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_ldiffsr.py > gpurun_out/r02/pytest_main.log 2>&1; tail -3 gpurun_out/r02/pytest_main.log
timeout 400 python -m pytest tests/test_gpu_ldiffsr.py -q -s > gpurun_out/r02/pytest_ldiffsr.log 2>&1; tail -40 gpurun_out/r02/pytest_ldiffsr.log
timeout 300 python bench.py --steps 100 --warmup 5 > gpurun_out/r02/bench_pdl0.json 2> gpurun_out/r02/bench_pdl0.err
RLREP_PDL=1 timeout 300 python bench.py --steps 100 --warmup 5 > gpurun_out/r02/bench_pdl1.json 2> gpurun_out/r02/bench_pdl1.err
python - <<'PY'
import json
for f in ("pdl0", "pdl1"):
    try:
        d = json.load(open(f"gpurun_out/r02/bench_{f}.json"))
        print(f, round(d["value"], 1), "upd/s", round(d["ms_per_step"], 4), "ms; e2e", round(d["e2e"]["value"], 1))
    except Exception as e:
        print(f, "failed", e)
PY
RLREP_PDL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "train_matches or graph_replay or ragged" 2>&1 | tail -5
RLREP_PDL=1 timeout 200 python tests/gpu_timeline.py ctrlsac_hc_b256 > gpurun_out/r02/timeline_pdl1.csv 2> /dev/null
du -sh gpurun_out
