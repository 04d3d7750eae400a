timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_drq.py -m gpu -q 2>&1 | tail -3
bash scripts/gpu_job20.sh
