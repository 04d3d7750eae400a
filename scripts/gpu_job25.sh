python -m pytest tests/test_gpu_parity.py -m gpu -q -k "ctrlsac or graph or curves" 2>&1 | tail -3
python bench.py --steps 100 --warmup 5 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'upd/s', round(d['ms_per_step'],4),'ms; e2e', round(d['e2e']['value'],1))"
