mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_chain.py -m gpu -q -x > gpurun_out/r02/pytest_dist.log 2>&1; tail -4 gpurun_out/r02/pytest_dist.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ctrlsac or device_noise or row_operations" > gpurun_out/r02/pytest_dist2.log 2>&1; tail -4 gpurun_out/r02/pytest_dist2.log
for d in 0 1 2; do
  RLREP_CHAIN_DIST=$d timeout 300 python bench.py --steps 300 --warmup 30 --no-sharded --no-cpu-baseline --no-alt-precision > gpurun_out/r02/bench_dist$d.json 2> gpurun_out/r02/bench_dist$d.err
done
RLREP_CHAIN_VERBOSE=1 timeout 120 python bench.py --steps 5 --warmup 3 --repeats 1 --no-sharded --no-cpu-baseline --no-alt-precision 2>&1 | grep -B1 -A24 "^rlrep chain" | head -150 > gpurun_out/r02/chain_dist_plan.txt
python - <<'PY'
import json
for f in ('bench_dist0','bench_dist1','bench_dist2'):
    try:
        d=json.loads(open(f'gpurun_out/r02/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d['gpu_launches_per_step'], d['roofline']['avg_launch_us'])
    except Exception as e:
        print(f, 'ERR', e)
PY
