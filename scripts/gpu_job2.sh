python tests/gpu_null.py
python tests/gpu_tune_gemm.py
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/r01_launches_ctrlsac_b256_v3.csv python tests/gpu_ncu_update.py ctrlsac_hc_b256 3 eager > /dev/null 2>&1
ncu --set full --clock-control none -k regex:adam_polyak -s 6 -c 2 -o gpurun_out/r01_adam_polyak python tests/gpu_ncu_update.py ctrlsac_hc_b256 3 eager > gpurun_out/ncu_adam.log 2>&1; tail -2 gpurun_out/ncu_adam.log
ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 172 -c 3 -o gpurun_out/r01_gemm_tf32 python tests/gpu_ncu_update.py ctrlsac_hc_b256 3 eager > gpurun_out/ncu_gemm.log 2>&1; tail -2 gpurun_out/ncu_gemm.log
ls -la gpurun_out; du -sh gpurun_out
