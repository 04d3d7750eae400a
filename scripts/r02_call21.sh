mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_parity.py -m gpu -q -x -k "chain or ctrlsac or graph_replay or row_operations" > gpurun_out/r02/pytest_bgadam.log 2>&1; tail -6 gpurun_out/r02/pytest_bgadam.log
for a in 0 1; do
  RLREP_CHAIN_ADAM=$a timeout 300 python bench.py --steps 300 --warmup 30 --no-sharded --no-cpu-baseline --no-alt-precision > gpurun_out/r02/bench_bgadam$a.json 2> gpurun_out/r02/bench_bgadam$a.err
done
python - <<'PY'
import json
for f in ('bench_bgadam0','bench_bgadam1'):
    try:
        d=json.loads(open(f'gpurun_out/r02/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d['gpu_launches_per_step'], d['top_kernels_us_per_step'][:4])
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -3 gpurun_out/r02/bench_bgadam1.err
