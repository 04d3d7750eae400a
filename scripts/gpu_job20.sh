python bench.py --workload drqv2_pixels_b256 --steps 30 --warmup 3 > gpurun_out/bench_drq.json 2> gpurun_out/bench_drq.err || tail -5 gpurun_out/bench_drq.err
python -c "
import json
d=json.load(open('gpurun_out/bench_drq.json'))
print('drq', round(d['value'],1),'upd/s', round(d['ms_per_step'],3),'ms; e2e', round(d['e2e']['value'],1), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'],2), 'launches', d['gpu_launches_per_step'])
for k in d['top_kernels_us_per_step']: print('  ', k)
r=d['roofline']; print(r['kernel'], r['bound'], round(r['frac'],3), r['step'])"
