mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_ldiffsr.py tests/test_gpu_parity.py -m gpu -q -x -k "ldiffsr or row_operations or ctrlsac_small" > gpurun_out/r02/pytest_39.log 2>&1; tail -3 gpurun_out/r02/pytest_39.log
for w in drqv2_pixels_b256 mulvdrq_pixels_b256 ldiffsr_pixels_b256; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --repeats 3 --no-alt-precision --no-cpu-baseline > gpurun_out/r02/bench_${w}_v9.json 2> gpurun_out/r02/bench_${w}_v9.err
  tail -2 gpurun_out/r02/bench_${w}_v9.err
done
python - <<'PY'
import json
for w in ('drqv2_pixels_b256','mulvdrq_pixels_b256','ldiffsr_pixels_b256'):
    try:
        d=json.loads(open(f'gpurun_out/r02/bench_{w}_v9.json').read().strip().splitlines()[-1])
        print(w, round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'e2e_dev', (d.get('e2e_device_replay') or {}).get('value'), d.get('gpu_launches_per_step'), d['roofline']['step']['frac'])
    except Exception as e:
        print(w, 'ERR', e)
PY
