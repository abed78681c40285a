# round 2, call n (2 GPUs): parity of the partitioned solve with the persistent coarse cycle, then the bench at 150^3 with
# both coarse modes (per-call profile on stderr)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631"
timeout 600 $T tools/dist_check.py 40 3 > gpurun_out/r2n_dist_check40.log 2>&1; echo rc=$? >> gpurun_out/r2n_dist_check40.log
B="$T bench.py --gpus 2 --cells 150 --steps 3 --warmup 2 --e2e-steps 0 --cpu-budget 0 --same-config-steps 0"
rm -f gpurun_out/r2n_summary.log
run() { tag=$1; shift; TEFEM_PROFILE_ALL=1 timeout 600 $B "$@" > gpurun_out/r2n_$tag.json 2> gpurun_out/r2n_$tag.err; echo "$tag rc=$?" >> gpurun_out/r2n_summary.log; grep -ah "tefem profile" gpurun_out/r2n_$tag.err | head -2 | cut -c1-300 >> gpurun_out/r2n_summary.log; python -c "
import json
d=json.loads(open('gpurun_out/r2n_$tag.json').read().strip().splitlines()[-1])
print('   steps/s %.3f ms/it %.3f iters %s parity %s'%(d['value'], d['solver']['ms_per_iteration'], d['solver']['iterations'], (d.get('parity_vs_1gpu') or {}).get('max_rel_diff')))" >> gpurun_out/r2n_summary.log 2>&1; }
run persistent
run launches --ml coarse=launches --parity-cells 0
grep -a "DIST_CHECK\|rc=" gpurun_out/r2n_dist_check40.log | tail -3; cat gpurun_out/r2n_summary.log
