mkdir -p gpurun_out/r02
RLREP_TC_VERBOSE=1 timeout 300 python tests/gpu_mulv_profile.py > gpurun_out/r02/mulv_profile_plans.log 2> gpurun_out/r02/mulv_plans.err
grep "rlrep tc plan" gpurun_out/r02/mulv_plans.err | sort | uniq -c | sort -k1 -n -r | awk '$0 ~ /N=39200|K=39200/' | head -40
