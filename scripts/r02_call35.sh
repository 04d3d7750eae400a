mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_drq.py tests/test_gpu_mulv.py -m gpu -q -x > gpurun_out/r02/pytest_groups.log 2>&1; tail -3 gpurun_out/r02/pytest_groups.log
for g in 1 2 4; do
  RLREP_WGRAD_GROUPS=$g timeout 600 python bench.py --workload mulvdrq_pixels_b256 --steps 10 --warmup 3 --repeats 3 --no-cpu-baseline --no-alt-precision > gpurun_out/r02/bench_mulv_groups$g.json 2> gpurun_out/r02/bench_mulv_groups$g.err
done
python - <<'PY'
import json
for g in (1,2,4):
    try:
        d=json.loads(open(f'gpurun_out/r02/bench_mulv_groups{g}.json').read().strip().splitlines()[-1])
        print('groups', g, round(d['value'],1), round(d['ms_per_step'],4), d['top_kernels_us_per_step'][:3])
    except Exception as e:
        print(g, 'ERR', e)
PY
