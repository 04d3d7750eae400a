set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for wl in ctrlsac_hc_b256 sac_hc_b256 vlsac_hum_b1024 spedersac_hc_b256 diffsrsac_hc_b256; do
  python bench.py --steps 100 --warmup 5 --workload $wl > gpurun_out/bench_s2_$wl.json 2> gpurun_out/bench_s2_$wl.err || tail -5 gpurun_out/bench_s2_$wl.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/bench_s2_$wl.json'))
print('$wl', round(d['value'],1), 'upd/s', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value'],1), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'],2), 'launches', d['gpu_launches_per_step'])
r=d['roofline']; print(' top', r['kernel'], r['bound'], round(r['frac'],3), 'share', round(r['share_of_step'],3), 'step', r['step'])
for k in r['kernels']: print('   ', k['kernel'], k['bound'], round(k['frac'],3), round(k['share_of_step'],3), round(k['avg_launch_us'],1),'us')
"
done
python bench.py --impl reference --steps 5 --warmup 1 | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/r01_launches_ctrlsac_b256_v3.csv python tests/gpu_ncu_update.py ctrlsac_hc_b256 3 eager > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:adam_polyak -s 6 -c 2 -o gpurun_out/r01_adam_polyak python tests/gpu_ncu_update.py ctrlsac_hc_b256 3 eager > gpurun_out/ncu_adam.log 2>&1; tail -2 gpurun_out/ncu_adam.log
ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 160 -c 12 -o gpurun_out/r01_gemm_tf32 python tests/gpu_ncu_update.py ctrlsac_hc_b256 3 eager > gpurun_out/ncu_gemm.log 2>&1; tail -2 gpurun_out/ncu_gemm.log
ls -la gpurun_out
