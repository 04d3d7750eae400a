mkdir -p gpurun_out/r02
timeout 300 python tests/gpu_pixel_timeline.py mulvdrq_pixels_b256 25 > gpurun_out/r02/mulv_timeline_v11.log 2>&1
grep "gemm_conv_halo" gpurun_out/r02/mulv_timeline_v11.log | cut -c1-50 | head -30
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_mulv.py -m gpu -q -x > gpurun_out/r02/pytest_43.log 2>&1; tail -2 gpurun_out/r02/pytest_43.log
