mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ldiffsr.py -m gpu -q -x -k "diffsr" > gpurun_out/r02/pytest_49.log 2>&1; tail -3 gpurun_out/r02/pytest_49.log
timeout 400 python bench.py --workload diffsrsac_hc_b256 --no-alt-precision --no-cpu-baseline > gpurun_out/r02/bench_diffsr_v2.json 2> gpurun_out/r02/bench_diffsr_v2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/bench_diffsr_v2.json').read().strip().splitlines()[-1])
print('diffsr', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['gpu_launches_per_step'])
print(d['top_kernels_us_per_step'])
PY
