python tests/gpu_debug_vlsac.py tf32 4 2>&1 | tail -8
timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python tests/gpu_debug_vlsac.py tf32 2 2>&1 | grep -v "^$" | head -60
