mkdir -p gpurun_out/r02
timeout 300 python tests/gpu_mulv_profile.py > gpurun_out/r02/mulv_profile_v20.log 2>&1
grep "gemm_tf32 " gpurun_out/r02/mulv_profile_v20.log | sort -k3 -n -r | head -16
