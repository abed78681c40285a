# round 2, call p (2 GPUs): dataflow coarse cycle with one warp per row; 120^3 on 2 GPUs has the per-rank sizes of config 4
# on 8 GPUs (864 k cells and 1 687 aggregates per rank)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "aggregation or two_gpu" > gpurun_out/r2p_gpu_tests.log 2>&1; echo rc=$? >> gpurun_out/r2p_gpu_tests.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631"
B="$T bench.py --gpus 2 --cells 120 --steps 3 --warmup 2 --e2e-steps 0 --cpu-budget 0 --same-config-steps 0"
rm -f gpurun_out/r2p_summary.log
run() { tag=$1; shift; timeout 600 $B "$@" > gpurun_out/r2p_$tag.json 2> gpurun_out/r2p_$tag.err; echo "$tag rc=$?" >> gpurun_out/r2p_summary.log; grep -ah "tefem profile\|tefem precond" gpurun_out/r2p_$tag.err | head -4 | cut -c1-400 >> gpurun_out/r2p_summary.log; python -c "
import json
d=json.loads(open('gpurun_out/r2p_$tag.json').read().strip().splitlines()[-1])
print('   steps/s %.3f ms/it %.3f iters %s parity %s'%(d['value'], d['solver']['ms_per_iteration'], d['solver']['iterations'], (d.get('parity_vs_1gpu') or {}).get('max_rel_diff')))" >> gpurun_out/r2p_summary.log 2>&1; }
export TEFEM_PROFILE_ALL=1
run persistent
run launches --ml coarse=launches --parity-cells 0
export TEFEM_PROFILE_PRECOND=1
run persistent_phases --parity-cells 0
run launches_phases --ml coarse=launches --parity-cells 0
tail -3 gpurun_out/r2p_gpu_tests.log; cat gpurun_out/r2p_summary.log
