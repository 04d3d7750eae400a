#!/bin/bash
# Round-2 closing run: full GPU suite, smoke, the driver's bench command, the other workloads' bench lines, ncu evidence of
# the kernels that changed last (chain kernel, halo convolution kernel), timeline, chain probe.  Condenses ncu reports on
# the box (gpurun_out/ must stay < 64 MiB).
set -u
O=gpurun_out/final; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > $O/r02_bench_default_n1.json 2> $O/bench_default.err; tail -c 300 $O/r02_bench_default_n1.json; echo
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference_n1.json 2> $O/bench_reference.err
timeout 400 python bench.py --workload vlsac_hum_b1024 > $O/r02_bench_vlsac_hum_b1024.json 2> $O/bench_vlsac.err
for w in drqv2_pixels_b256 mulvdrq_pixels_b256 ldiffsr_pixels_b256; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 > $O/r02_bench_$w.json 2> $O/bench_$w.err
done
for w in sac_hc_b256 spedersac_hc_b256 diffsrsac_hc_b256; do
  timeout 400 python bench.py --workload $w --no-alt-precision > $O/r02_bench_$w.json 2> $O/bench_$w.err
done
timeout 200 python tests/gpu_timeline.py ctrlsac_hc_b256 > $O/r02_timeline_ctrlsac_b256.csv 2> $O/timeline.err
for w in fwd bwd; do timeout 120 python tests/gpu_chain_probe.py $w; done > $O/r02_chain_probe_final.log 2>&1
timeout 120 python tests/gpu_mulv_profile.py > $O/r02_launches_mulvdrq_b256.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_conv_halo -s 6 -c 3 -o gpurun_out/r02_gemm_conv_halo python tests/gpu_mulv_profile.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_chain -s 14 -c 4 -o gpurun_out/r02_gemm_chain python tests/gpu_ncu_update.py ctrlsac_hc_b256 3 eager > /dev/null 2>&1
python scripts/ncu_summary.py mulvdrq_pixels_b256 gpurun_out/r02_gemm_conv_halo.ncu-rep 2>&1 | tail -2
python scripts/ncu_summary.py ctrlsac_hc_b256 gpurun_out/r02_gemm_chain.ncu-rep 2>&1 | tail -2
cp profiles/r02_gemm_conv_halo_summary.csv profiles/r02_gemm_chain_summary.csv profiles/ncu_traffic.json $O/ 2>/dev/null
rm -f gpurun_out/*.ncu-rep
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/final/r02_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['value'],1), d['unit'], round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), (d.get('e2e_device_noise') or {}).get('value'), 'launches', d.get('gpu_launches_per_step'))
    except Exception as e:
        print(f, 'ERR', e)
PY
