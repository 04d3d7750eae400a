mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_drq.py -m gpu -q -x > gpurun_out/r02/pytest_64.log 2>&1; tail -2 gpurun_out/r02/pytest_64.log
timeout 600 python bench.py --workload drqv2_pixels_b256 --steps 10 --warmup 3 --repeats 3 --no-cpu-baseline --no-alt-precision > gpurun_out/r02/bench_drq_v21.json 2> gpurun_out/r02/bench_drq_v21.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/bench_drq_v21.json').read().strip().splitlines()[-1])
print('drq', round(d['value'],1), round(d['ms_per_step'],4), [k for k in d['top_kernels_us_per_step'] if 'im2col' in k[0]])
PY
