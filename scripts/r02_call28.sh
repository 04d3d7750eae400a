mkdir -p gpurun_out/r02
timeout 300 python tests/gpu_mulv_profile.py > gpurun_out/r02/mulv_profile_v6.log 2>&1
tail -5 gpurun_out/r02/mulv_profile_v6.log
