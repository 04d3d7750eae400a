#!/bin/bash
# Round-1 closing artifacts for the pixel path + a re-run of the headline bench with the final library.
# Condenses the ncu report on the box (gpurun_out/ must stay < 64 MiB).
set -u
mkdir -p gpurun_out/profiles_new
timeout 200 python bench.py > gpurun_out/profiles_new/r01_bench_final_ctrlsac_hc_b256.json 2> gpurun_out/bench_final.err
tail -c 400 gpurun_out/profiles_new/r01_bench_final_ctrlsac_hc_b256.json | head -c 400; echo
timeout 90 ncu --set full --clock-control none -k regex:gemm_tf32_persistent -s 30 -c 3 -o gpurun_out/r01_gemm_persistent_mulv python tests/gpu_mulv_profile.py > /dev/null 2>&1
python scripts/ncu_summary.py mulvdrq_pixels_b256 gpurun_out/r01_gemm_persistent_mulv.ncu-rep
cp profiles/r01_gemm_persistent_mulv_summary.csv profiles/ncu_traffic.json gpurun_out/profiles_new/ 2>/dev/null
rm -f gpurun_out/*.ncu-rep
timeout 60 python tests/gpu_mulv_profile.py > gpurun_out/profiles_new/r01_launches_mulvdrq_b256_final.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
ls -la gpurun_out/profiles_new
