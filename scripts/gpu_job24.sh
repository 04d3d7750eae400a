for cfg in "0.5 0.25" "0.65 0.25" "0.74 0.25" "0.5 0.5" "0.4 0.25"; do
  set -- $cfg
  echo "== dual_share=$1 aux_share=$2"
  RLREP_DUAL_SHARE=$1 RLREP_AUX_SHARE=$2 python bench.py --steps 100 --warmup 5 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'upd/s', round(d['ms_per_step'],4),'ms')"
done
