mkdir -p gpurun_out/r02
timeout 1200 python -m pytest tests/test_gpu_drq.py tests/test_gpu_mulv.py tests/test_gpu_ldiffsr.py tests/test_gpu_pixreplay.py -m gpu -q -x > gpurun_out/r02/pytest_pix_iw.log 2>&1; tail -5 gpurun_out/r02/pytest_pix_iw.log
for w in mulvdrq_pixels_b256 ldiffsr_pixels_b256; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --repeats 3 --no-cpu-baseline --no-alt-precision > gpurun_out/r02/bench_${w}_iw.json 2> gpurun_out/r02/bench_${w}_iw.err
done
python - <<'PY'
import json
for f in ('bench_mulvdrq_pixels_b256_iw','bench_ldiffsr_pixels_b256_iw'):
    try:
        d=json.loads(open(f'gpurun_out/r02/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d.get('gpu_launches_per_step'), d['top_kernels_us_per_step'][:8])
    except Exception as e:
        print(f, 'ERR', e)
PY
