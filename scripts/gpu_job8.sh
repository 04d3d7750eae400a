python -m pytest tests -m gpu -q 2>&1 | tail -4
for sh in 1 0; do
echo "== shallow=$sh"
RLREP_TC_SHALLOW=$sh python bench.py --steps 100 --warmup 5 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'upd/s', round(d['ms_per_step'],4),'ms; e2e', round(d['e2e']['value'],1)); print(d['top_kernels_us_per_step'][:8])"
done
python tests/gpu_timeline.py ctrlsac_hc_b256 > gpurun_out/timeline_ctrlsac_v2.csv 2> gpurun_out/timeline.err
