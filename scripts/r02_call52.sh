mkdir -p gpurun_out/r02
timeout 900 python bench.py --workload ldiffsr_pixels_b256 --steps 10 --warmup 3 --repeats 3 --no-alt-precision --no-cpu-baseline > gpurun_out/r02/bench_ldiffsr_v15.json 2> gpurun_out/r02/bench_ldiffsr_v15.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/bench_ldiffsr_v15.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],4), [k for k in d['top_kernels_us_per_step'] if 'out_conv' in k[0]])
PY
