for wgt in 1 2 4; do
echo "== l2_weight=$wgt"
RLREP_L2_WEIGHT=$wgt python bench.py --steps 100 --warmup 5 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'upd/s', round(d['ms_per_step'],4),'ms; e2e', round(d['e2e']['value'],1)); print(d['top_kernels_us_per_step'][:3])"
done
