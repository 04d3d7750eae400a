RLREP_TC_PERSIST=1 timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_conv.py -m gpu -q -x -k "gemm or conv" 2>&1 | tail -3
for p in 1 0; do
echo "== persist=$p"
RLREP_TC_PERSIST=$p python bench.py --workload drqv2_pixels_b256 --steps 30 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'upd/s', round(d['ms_per_step'],3),'ms'); print(d['top_kernels_us_per_step'][:3])"
done
RLREP_TC_PERSIST=1 python tests/gpu_tune_gemm.py 2>&1 | grep "big" | cut -c1-200
