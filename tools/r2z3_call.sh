# round 2, final call 3 (2 GPUs; tagged partial sums pushed to the owner, no start flags, two alternating areas): parity of the partitioned solves (bench path and public API, incl. the aggregation
# preconditioner with the persistent partitioned cycle), bench at 120^3 with per-call profile
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "aggregation or two_gpu" > gpurun_out/r2z3_gpu_tests.log 2>&1; echo rc=$? >> gpurun_out/r2z3_gpu_tests.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631"
timeout 600 $T tools/dist_check.py 40 3 > gpurun_out/r2z3_dist_check40.log 2>&1; echo rc=$? >> gpurun_out/r2z3_dist_check40.log
timeout 900 $T tools/dist_api_check.py > gpurun_out/r2z3_dist_api_check.log 2>&1; echo rc=$? >> gpurun_out/r2z3_dist_api_check.log
B="$T bench.py --gpus 2 --cells 120 --steps 3 --warmup 2 --e2e-steps 0 --cpu-budget 0 --same-config-steps 0 --parity-cells 0"
TEFEM_PROFILE_ALL=1 timeout 600 $B > gpurun_out/r2z3_persistent.json 2> gpurun_out/r2z3_persistent.err; echo rc=$?
TEFEM_PROFILE_PRECOND=1 timeout 600 $B --steps 2 --warmup 1 > gpurun_out/r2z3_phases.json 2> gpurun_out/r2z3_phases.err; echo rc=$?
tail -2 gpurun_out/r2z3_gpu_tests.log; grep -a "DIST_CHECK\|rc=" gpurun_out/r2z3_dist_check40.log | tail -2; grep -a "rel diff\|iterations\|DIST_API\|rc=" gpurun_out/r2z3_dist_api_check.log | tail -14
grep -a "tefem profile" gpurun_out/r2z3_persistent.err | head -1 | cut -c1-200
grep -a "tefem precond" gpurun_out/r2z3_phases.err | head -5 | cut -c1-300
