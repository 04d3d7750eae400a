mkdir -p gpurun_out/r02
timeout 300 python tests/gpu_pixel_timeline.py mulvdrq_pixels_b256 25 > gpurun_out/r02/mulv_timeline_v10.log 2>&1
grep -n "gemm_conv_halo\|pad_grid\|gemm_tf32_persistent" gpurun_out/r02/mulv_timeline_v10.log | head -60
