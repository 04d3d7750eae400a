python tests/gpu_trace.py
