# round 2 profiles: launch list, ncu --set full of the chain kernel and the fused optimiser, CUPTI timeline, bench lines
mkdir -p gpurun_out/r02 gpurun_out/profiles_new
timeout 300 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -2
for cfg in "fwd 0 0" "bwd 0 0"; do timeout 120 python tests/gpu_chain_probe.py $cfg 2>&1 | tee -a gpurun_out/r02/chain_probe7.log | grep -v "^  gemm" | cut -c1-380; done
timeout 300 python bench.py --steps 50 --warmup 5 --repeats 5 --no-cpu-baseline --no-sharded --no-alt-precision > gpurun_out/r02/bench_chain_v7.json 2> gpurun_out/r02/bench_chain_v7.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02/bench_chain_v7.json"))
print("chain v7:", round(d["value"], 1), "upd/s", round(d["ms_per_step"], 4), "ms; e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches_per_step"])
print("   top", d["top_kernels_us_per_step"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/profiles_new/r02_launches_ctrlsac_b256.csv python tests/gpu_ncu_update.py ctrlsac_hc_b256 3 eager > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_chain -s 14 -c 4 -o gpurun_out/r02_gemm_chain python tests/gpu_ncu_update.py ctrlsac_hc_b256 3 eager > gpurun_out/r02/ncu_chain.log 2>&1
ncu --set full --clock-control none -k regex:adam_polyak -s 6 -c 2 -o gpurun_out/r02_adam_polyak python tests/gpu_ncu_update.py ctrlsac_hc_b256 3 eager > /dev/null 2>&1
ncu --set full --clock-control none -k regex:contrastive_head -s 4 -c 1 -o gpurun_out/r02_contrastive_head python tests/gpu_ncu_update.py ctrlsac_hc_b256 3 eager > /dev/null 2>&1
python scripts/ncu_summary.py ctrlsac_hc_b256 gpurun_out/r02_gemm_chain.ncu-rep gpurun_out/r02_adam_polyak.ncu-rep gpurun_out/r02_contrastive_head.ncu-rep
cp profiles/r02_*_summary.csv profiles/ncu_traffic.json gpurun_out/profiles_new/ 2>/dev/null
ncu -i gpurun_out/r02_gemm_chain.ncu-rep --page details --csv 2>/dev/null | head -400 > gpurun_out/profiles_new/r02_gemm_chain_details.csv
rm -f gpurun_out/*.ncu-rep
timeout 200 python tests/gpu_timeline.py ctrlsac_hc_b256 > gpurun_out/profiles_new/r02_timeline_ctrlsac_b256.csv 2> /dev/null
ls -la gpurun_out/profiles_new; du -sh gpurun_out
