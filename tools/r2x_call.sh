# round 2, call x (8 GPUs, config 4): volatile tagged stores, L2 prefetch, CTA-per-row dense product, r2 gather; coarse cycle as separate launches (replicated levels)
# and as the row-partitioned dataflow kernel
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631"
B="$T bench.py --gpus 8 --steps 5 --warmup 3 --e2e-steps 0 --cpu-budget 0 --same-config-steps 0"
rm -f gpurun_out/r2x_summary.log
run() { tag=$1; shift; timeout 500 $B "$@" > gpurun_out/r2x_$tag.json 2> gpurun_out/r2x_$tag.err; echo "$tag rc=$?" >> gpurun_out/r2x_summary.log; grep -ah "tefem precond" gpurun_out/r2x_$tag.err | head -3 | cut -c1-400 >> gpurun_out/r2x_summary.log; python -c "
import json
d=json.loads(open('gpurun_out/r2x_$tag.json').read().strip().splitlines()[-1])
s=d['solver']
print('   steps/s %.3f ms/step %.2f solve %.2f other %.2f ms/it %.3f iters %s spmv %.4f ms frac %.3f parity %s'%(d['value'], d['ms_per_step'], s.get('solve_ms_per_step',0), s.get('other_ms_per_step',0), s['ms_per_iteration'], s['iterations'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], (d.get('parity_vs_1gpu') or {}).get('max_rel_diff')))" >> gpurun_out/r2x_summary.log 2>&1; }
run launches --ml coarse=launches --parity-cells 0
run persistent
TEFEM_PROFILE_PRECOND=1 run persistent_phases --parity-cells 0 --steps 2 --warmup 1
cat gpurun_out/r2x_summary.log
