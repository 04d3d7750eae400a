mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ctrlsac or graph_replay or checkpoint or batch_size or population or row_operations" > gpurun_out/r02/pytest_heads.log 2>&1; tail -4 gpurun_out/r02/pytest_heads.log
timeout 300 python bench.py --steps 300 --warmup 30 --no-sharded --no-cpu-baseline --no-alt-precision > gpurun_out/r02/bench_heads.json 2> gpurun_out/r02/bench_heads.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/bench_heads.json').read().strip().splitlines()[-1])
print('heads', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d['gpu_launches_per_step'], d['top_kernels_us_per_step'])
PY
