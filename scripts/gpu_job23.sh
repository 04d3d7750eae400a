python -m pytest tests -m gpu -q 2>&1 | tail -4
for wl in ctrlsac_hc_b256 sac_hc_b256 vlsac_hum_b1024 spedersac_hc_b256 diffsrsac_hc_b256; do
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --workload $wl | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$wl', round(d['value'],1),'upd/s', round(d['ms_per_step'],4),'ms; e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches_per_step'])"
done
