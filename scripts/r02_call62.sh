mkdir -p gpurun_out/r02
for c in 148 296; do
  RLREP_GEMM_KGROUP_CTAS=$c timeout 600 python bench.py --workload mulvdrq_pixels_b256 --steps 10 --warmup 3 --repeats 3 --no-cpu-baseline --no-alt-precision > gpurun_out/r02/bench_mulv_kg$c.json 2> gpurun_out/r02/bench_mulv_kg$c.err
done
python - <<'PY'
import json
for c in (148,296):
    d=json.loads(open(f'gpurun_out/r02/bench_mulv_kg{c}.json').read().strip().splitlines()[-1])
    print('cap', c, round(d['value'],1), round(d['ms_per_step'],4), d['top_kernels_us_per_step'][:1])
PY
