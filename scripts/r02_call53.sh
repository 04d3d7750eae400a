mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests/test_gpu_conv.py tests/test_gpu_drq.py tests/test_gpu_mulv.py tests/test_gpu_ldiffsr.py -m gpu -q -x > gpurun_out/r02/pytest_53.log 2>&1; tail -3 gpurun_out/r02/pytest_53.log
for w in drqv2_pixels_b256 mulvdrq_pixels_b256 ldiffsr_pixels_b256; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --repeats 3 --no-alt-precision --no-cpu-baseline > gpurun_out/r02/bench_${w}_v20.json 2> gpurun_out/r02/bench_${w}_v20.err
done
python - <<'PY'
import json
for w in ('drqv2_pixels_b256','mulvdrq_pixels_b256','ldiffsr_pixels_b256'):
    try:
        d=json.loads(open(f'gpurun_out/r02/bench_{w}_v20.json').read().strip().splitlines()[-1])
        print(w, round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['top_kernels_us_per_step'][:2])
    except Exception as e:
        print(w, 'ERR', e)
PY
