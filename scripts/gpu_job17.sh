timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -s 2>&1 | grep -o "C=[0-9]* B=[0-9]* [a-z_]* [a-z0-9]*: .*\|[0-9]* passed.*\|[0-9]* failed.*" | cut -c1-600
