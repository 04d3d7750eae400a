timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -s 2>&1 | tail -4
bash scripts/gpu_job9.sh 8 ctrlsac_b16384_sharded 10
bash scripts/gpu_job9.sh 8 ctrlsac_hc_b256 100
