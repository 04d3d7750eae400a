timeout 500 python -m pytest tests/test_gpu_drq.py -m gpu -q -s 2>&1 | grep -o "drqv2 .*\|[0-9]* passed.*\|[0-9]* failed.*\|Error.*" | cut -c1-420
