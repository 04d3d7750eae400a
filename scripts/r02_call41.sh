mkdir -p gpurun_out/r02
timeout 300 python tests/gpu_mulv_profile.py > gpurun_out/r02/mulv_profile_v10.log 2>&1
grep -n "gemm_conv_halo\|pad_grid\|col2im" gpurun_out/r02/mulv_profile_v10.log | head -40
