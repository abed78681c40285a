mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655"
timeout 600 $TR tools/dist_api_check.py > gpurun_out/r2f_dist_api_check.log 2>&1; echo rc=$? >> gpurun_out/r2f_dist_api_check.log
timeout 300 $TR tools/dist_check.py 40 3 > gpurun_out/r2f_dist_check40.log 2>&1; echo rc=$? >> gpurun_out/r2f_dist_check40.log
TEFEM_PROFILE_ALL=1 timeout 600 $TR bench.py --gpus 2 --cells 150 --steps 3 --warmup 2 --e2e-steps 2 > gpurun_out/r2f_bench150_2gpu.json 2> gpurun_out/r2f_bench150_2gpu.err; echo rc=$? >> gpurun_out/r2f_bench150_2gpu.err
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2f_dist_api_check.log | tail -25; tail -4 gpurun_out/r2f_dist_check40.log; grep -h "tefem profile\|rc=\|Error\|error" gpurun_out/r2f_bench150_2gpu.err | tail -5
