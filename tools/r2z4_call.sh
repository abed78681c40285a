# round 2, final call 4 (8 GPUs): the default bench line at config 4 (as the driver runs it), then the per-call profile
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631"
timeout 600 $T bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2z_bench_8gpu.json 2> gpurun_out/r2z_bench_8gpu.err; echo rc=$?
TEFEM_PROFILE_ALL=1 timeout 600 $T bench.py --gpus 8 --steps 3 --warmup 2 --e2e-steps 0 --parity-cells 0 > gpurun_out/r2z_bench_8gpu_profile.json 2> gpurun_out/r2z_bench_8gpu_profile.err; echo rc=$?
python -c "
import json
for f in ('r2z_bench_8gpu','r2z_bench_8gpu_profile'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    s=d['solver']
    print(f, 'steps/s %.3f e2e %s ms/step %.2f solve %.2f other %.2f ms/it %.3f iters %s spmv %.4f ms frac %.3f parity %s'%(d['value'], (d.get('e2e') or {}).get('value'), d['ms_per_step'], s.get('solve_ms_per_step',0), s.get('other_ms_per_step',0), s['ms_per_iteration'], s['iterations'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], (d.get('parity_vs_1gpu') or {}).get('max_rel_diff')))"
grep -ah "tefem profile" gpurun_out/r2z_bench_8gpu_profile.err | head -2 | cut -c1-220
