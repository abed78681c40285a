# round 2, final call 5 (N GPUs, N from the environment): the default bench line at config 4
mkdir -p gpurun_out
N=${NGPU:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631"
timeout 900 $T bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2z_bench_${N}gpu.json 2> gpurun_out/r2z_bench_${N}gpu.err; echo rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/r2z_bench_${N}gpu.json').read().strip().splitlines()[-1])
s=d['solver']
print('N=${N} steps/s %.3f e2e %s ms/step %.2f solve %.2f other %.2f ms/it %.3f iters %s spmv %.4f ms frac %.3f parity %s asm MKD %.2f G el/s'%(d['value'], (d.get('e2e') or {}).get('value'), d['ms_per_step'], s.get('solve_ms_per_step',0), s.get('other_ms_per_step',0), s['ms_per_iteration'], s['iterations'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], (d.get('parity_vs_1gpu') or {}).get('max_rel_diff'), d['assembly']['MKD']['elements_per_s']/1e9))"
