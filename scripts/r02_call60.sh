mkdir -p gpurun_out/r02
RLREP_TC_PERSIST=1 timeout 600 python bench.py --workload mulvdrq_pixels_b256 --steps 10 --warmup 3 --repeats 3 --no-cpu-baseline --no-alt-precision > gpurun_out/r02/bench_mulv_persist1.json 2> gpurun_out/r02/bench_mulv_persist1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/bench_mulv_persist1.json').read().strip().splitlines()[-1])
print('persist=1', round(d['value'],1), round(d['ms_per_step'],4), d['top_kernels_us_per_step'][:4])
PY
