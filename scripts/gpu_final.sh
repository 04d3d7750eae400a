# Final round-1 validation on one B200: full GPU test suite, smoke, every bench workload, profiler evidence.
# Large .ncu-rep files are condensed on the box (scripts/ncu_summary.py) and deleted: gpurun_out/ must stay < 64 MiB.
python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for wl in ctrlsac_hc_b256 sac_hc_b256 vlsac_hum_b1024 spedersac_hc_b256 diffsrsac_hc_b256 drqv2_pixels_b256; do
  python bench.py --steps 100 --warmup 5 --workload $wl > gpurun_out/r01_bench_final_$wl.json 2> gpurun_out/bench_final_$wl.err || tail -5 gpurun_out/bench_final_$wl.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r01_bench_final_$wl.json'))
r=d['roofline']
print('$wl', round(d['value'],1),'upd/s', round(d['ms_per_step'],3),'ms; e2e', round(d['e2e']['value'],1), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'],2), 'launches', d['gpu_launches_per_step'], '| top', r['kernel'], r['bound'], round(r['frac'],3), 'step frac', round(r['step']['frac'],3))
PY
done
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r01_bench_final_reference_ctrlsac.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/r01_launches_ctrlsac_b256_final.csv python tests/gpu_ncu_update.py ctrlsac_hc_b256 3 eager > /dev/null 2>&1
ncu --set full --clock-control none -k regex:adam_polyak -s 6 -c 2 -o gpurun_out/r01_adam_polyak_final python tests/gpu_ncu_update.py ctrlsac_hc_b256 3 eager > /dev/null 2>&1
ncu --set full --clock-control none -k regex:gemm_tf32 -s 172 -c 3 -o gpurun_out/r01_gemm_tf32_final python tests/gpu_ncu_update.py ctrlsac_hc_b256 3 eager > /dev/null 2>&1
mkdir -p gpurun_out/profiles_new
python scripts/ncu_summary.py ctrlsac_hc_b256 gpurun_out/r01_adam_polyak_final.ncu-rep gpurun_out/r01_gemm_tf32_final.ncu-rep
cp profiles/r01_adam_polyak_final_summary.csv profiles/r01_gemm_tf32_final_summary.csv profiles/ncu_traffic.json gpurun_out/profiles_new/
rm -f gpurun_out/*.ncu-rep
python tests/gpu_timeline.py ctrlsac_hc_b256 > gpurun_out/r01_timeline_ctrlsac_b256_final.csv 2> /dev/null
du -sh gpurun_out
