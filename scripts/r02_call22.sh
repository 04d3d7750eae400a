mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -s > gpurun_out/r02/pytest_conv_iw.log 2>&1; tail -15 gpurun_out/r02/pytest_conv_iw.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ctrlsac_small or select_action" > gpurun_out/r02/pytest_tick.log 2>&1; tail -3 gpurun_out/r02/pytest_tick.log
for v in 1 0; do
  RLREP_CONV_WGRAD_V1=$v timeout 300 python bench.py --workload drqv2_pixels_b256 --steps 20 --warmup 5 --repeats 3 --no-cpu-baseline --no-alt-precision > gpurun_out/r02/bench_drq_iw$v.json 2> gpurun_out/r02/bench_drq_iw$v.err
done
python - <<'PY'
import json
for f in ('bench_drq_iw1','bench_drq_iw0'):
    try:
        d=json.loads(open(f'gpurun_out/r02/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d.get('gpu_launches_per_step'), d['top_kernels_us_per_step'][:6])
    except Exception as e:
        print(f, 'ERR', e)
PY
