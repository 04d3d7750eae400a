mkdir -p gpurun_out/r02
# row operations inside the chain (gather + contrastive head): correctness first, then the headline both ways
timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_gpu_parity.py -m gpu -q -x -k "chain or ctrlsac" > gpurun_out/r02/pytest_rowops.log 2>&1; tail -6 gpurun_out/r02/pytest_rowops.log
for ro in 0 1; do
  RLREP_CHAIN_ROWOPS=$ro timeout 300 python bench.py --steps 300 --warmup 30 --no-sharded --no-cpu-baseline --no-alt-precision > gpurun_out/r02/bench_rowops$ro.json 2> gpurun_out/r02/bench_rowops$ro.err; tail -c 1500 gpurun_out/r02/bench_rowops$ro.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rowops=$ro', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
done
RLREP_CHAIN_VERBOSE=1 timeout 120 python bench.py --steps 5 --warmup 3 --no-sharded --no-cpu-baseline --no-alt-precision 2>&1 | grep -A40 "rlrep chain: 2[0-9] GEMMs" | head -60 > gpurun_out/r02/chain_rowops_plan.txt
