mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_parity.py -m gpu -q -x -k "chain or ctrlsac or device_noise or cuda_generator" > gpurun_out/r02/pytest_rowops2.log 2>&1; tail -6 gpurun_out/r02/pytest_rowops2.log
for ro in 0 1; do
  RLREP_CHAIN_ROWOPS=$ro timeout 300 python bench.py --steps 300 --warmup 30 --no-sharded --no-cpu-baseline --no-alt-precision > gpurun_out/r02/bench_rowops${ro}b.json 2> gpurun_out/r02/bench_rowops${ro}b.err
done
timeout 300 python bench.py --workload vlsac_hum_b1024 --steps 200 --warmup 20 --no-cpu-baseline --no-alt-precision > gpurun_out/r02/bench_vlsac_dn.json 2> gpurun_out/r02/bench_vlsac_dn.err
python - <<'PY'
import json
for f in ('bench_rowops0b','bench_rowops1b','bench_vlsac_dn'):
    try:
        d=json.loads(open(f'gpurun_out/r02/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d.get('e2e_device_noise',{}).get('value'), d['gpu_launches_per_step'])
    except Exception as e:
        print(f, 'ERR', e)
PY
