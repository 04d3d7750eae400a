python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gemm or ctrlsac or graph" 2>&1 | tail -3
python tests/gpu_tune_gemm.py 2>&1 | cut -c1-330
for cfg in "1 0.5" "1 1.0" "0 0.5" "0 1.0" "1 0.35"; do
  set -- $cfg
  echo "== shallow=$1 dual_share=$2"
  RLREP_TC_SHALLOW=$1 RLREP_DUAL_SHARE=$2 python bench.py --steps 100 --warmup 5 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'upd/s', round(d['ms_per_step'],4),'ms; e2e', round(d['e2e']['value'],1)); print(d['top_kernels_us_per_step'][:4])"
done
