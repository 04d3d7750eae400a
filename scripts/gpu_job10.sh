N=$1
timeout 200 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -s 2>&1 | tail -4
bash scripts/gpu_job9.sh $N
