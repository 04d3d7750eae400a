python -m pytest tests -m gpu -q 2>&1 | tail -6
python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "diffsrsac and tf32" 2>&1 | grep -i "worst" | cut -c1-300
for cfg in "1 0.25" "1 0.5" "0 0.5"; do
  set -- $cfg
  echo "== use_aux=$1 aux_share=$2"
  RLREP_USE_AUX=$1 RLREP_AUX_SHARE=$2 python bench.py --steps 100 --warmup 5 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'upd/s', round(d['ms_per_step'],4),'ms; e2e', round(d['e2e']['value'],1)); print(d['top_kernels_us_per_step'][:8])"
done
for wl in sac_hc_b256 vlsac_hum_b1024 spedersac_hc_b256 diffsrsac_hc_b256; do
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --workload $wl | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$wl', round(d['value'],1),'upd/s', round(d['ms_per_step'],4),'ms; e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches_per_step']); print('  ', d['top_kernels_us_per_step'][:8])"
done
