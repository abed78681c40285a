# round 2, call y (2 GPUs): batched loads in the level-1 rows of the dataflow coarse cycle: correctness + time line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "aggregation or two_gpu" > gpurun_out/r2y_gpu_tests.log 2>&1; echo rc=$? >> gpurun_out/r2y_gpu_tests.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631"
timeout 600 $T tools/dist_check.py 40 3 > gpurun_out/r2y_dist_check40.log 2>&1; echo rc=$? >> gpurun_out/r2y_dist_check40.log
B="$T bench.py --gpus 2 --cells 120 --steps 2 --warmup 1 --e2e-steps 0 --cpu-budget 0 --same-config-steps 0 --parity-cells 0"
TEFEM_PROFILE_ALL=1 timeout 600 $B > gpurun_out/r2y_persistent.json 2> gpurun_out/r2y_persistent.err; echo rc=$?
TEFEM_PROFILE_PRECOND=1 timeout 600 $B > gpurun_out/r2y_phases.json 2> gpurun_out/r2y_phases.err; echo rc=$?
tail -2 gpurun_out/r2y_gpu_tests.log; grep -a "DIST_CHECK\|rc=" gpurun_out/r2y_dist_check40.log | tail -3
grep -a "tefem profile" gpurun_out/r2y_persistent.err | head -1 | cut -c1-200
grep -a "tefem precond" gpurun_out/r2y_phases.err | head -5 | cut -c1-300
