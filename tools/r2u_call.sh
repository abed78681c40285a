# round 2, call u (2 GPUs): volatile tagged stores, L2 prefetch of the own operator rows, CTA-per-row dense product
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "aggregation or two_gpu" > gpurun_out/r2u_gpu_tests.log 2>&1; echo rc=$? >> gpurun_out/r2u_gpu_tests.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631"
timeout 600 $T tools/dist_check.py 40 3 > gpurun_out/r2u_dist_check40.log 2>&1; echo rc=$? >> gpurun_out/r2u_dist_check40.log
B="$T bench.py --gpus 2 --cells 120 --steps 3 --warmup 2 --e2e-steps 0 --cpu-budget 0 --same-config-steps 0 --parity-cells 0"
rm -f gpurun_out/r2u_summary.log
run() { tag=$1; shift; timeout 600 $B "$@" > gpurun_out/r2u_$tag.json 2> gpurun_out/r2u_$tag.err; echo "$tag rc=$?" >> gpurun_out/r2u_summary.log; grep -ah "tefem profile\|tefem precond" gpurun_out/r2u_$tag.err | head -3 | cut -c1-400 >> gpurun_out/r2u_summary.log; python -c "
import json
d=json.loads(open('gpurun_out/r2u_$tag.json').read().strip().splitlines()[-1])
print('   steps/s %.3f ms/it %.3f iters %s'%(d['value'], d['solver']['ms_per_iteration'], d['solver']['iterations']))" >> gpurun_out/r2u_summary.log 2>&1; }
export TEFEM_PROFILE_ALL=1
run persistent
export TEFEM_PROFILE_PRECOND=1
run persistent_phases
tail -2 gpurun_out/r2u_gpu_tests.log; grep -a "DIST_CHECK\|rc=" gpurun_out/r2u_dist_check40.log | tail -3; cat gpurun_out/r2u_summary.log
