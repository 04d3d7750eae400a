mkdir -p gpurun_out/r02
for g in 2 4 6 8; do
  RLREP_WGRAD_GROUPS=$g timeout 600 python bench.py --workload mulvdrq_pixels_b256 --steps 10 --warmup 3 --repeats 3 --no-cpu-baseline --no-alt-precision > gpurun_out/r02/bench_mulv_g$g.json 2> gpurun_out/r02/bench_mulv_g$g.err
done
python - <<'PY'
import json
for g in (2,4,6,8):
    try:
        d=json.loads(open(f'gpurun_out/r02/bench_mulv_g{g}.json').read().strip().splitlines()[-1])
        print('groups', g, round(d['value'],1), round(d['ms_per_step'],4), d['top_kernels_us_per_step'][:1])
    except Exception as e:
        print(g, 'ERR', e)
PY
