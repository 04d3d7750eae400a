mkdir -p gpurun_out/r02
for a in 0 1; do
  RLREP_ADAM_CACHE=$a timeout 300 python bench.py --steps 300 --warmup 30 --no-sharded --no-cpu-baseline --no-alt-precision > gpurun_out/r02/bench_adamcache$a.json 2> gpurun_out/r02/bench_adamcache$a.err
done
python - <<'PY'
import json
for a in (0,1):
    try:
        d=json.loads(open(f'gpurun_out/r02/bench_adamcache{a}.json').read().strip().splitlines()[-1])
        print('adam cache', a, round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d['top_kernels_us_per_step'][:4])
    except Exception as e:
        print(a, 'ERR', e)
PY
