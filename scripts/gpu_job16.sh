timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gemm or graph or ctrlsac" 2>&1 | tail -4
python tests/gpu_tune_gemm.py 2>&1 | cut -c1-200 | head -11
for push in 1 0; do
echo "== push=$push"
RLREP_TC_PUSH=$push python bench.py --steps 100 --warmup 5 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'upd/s', round(d['ms_per_step'],4),'ms; e2e', round(d['e2e']['value'],1)); print(d['top_kernels_us_per_step'][:3])"
done
