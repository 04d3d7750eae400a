N=$1
WL=${2:-ctrlsac_b16384_sharded}
STEPS=${3:-10}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps $STEPS --warmup 3 --workload $WL --no-cpu-baseline > gpurun_out/bench_${WL}_n$N.json 2> gpurun_out/bench_${WL}_n$N.err || tail -20 gpurun_out/bench_${WL}_n$N.err
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_${WL}_n$N.json') if l.startswith('{')][0]
print('$WL N=$N', round(d['value'],2),'upd/s', round(d['ms_per_step'],3),'ms; e2e', round(d['e2e']['value'],2)); print(d['top_kernels_us_per_step'][:10]); r=d['roofline']; print(r['kernel'], r['bound'], round(r['frac'],3), r['step'])"
head -c 200 gpurun_out/bench_${WL}_n$N.json
