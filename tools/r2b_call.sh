mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655"
( echo "== K1 dynamic chunks (default)"; timeout 300 python tools/asm_ab.py 100 512,620,8,1,1;
  echo "== K1 tile-synchronous, 256 threads"; TEFEM_ASM_THREADS=256 timeout 300 python tools/asm_ab.py 100 512,620,8,1,1 ) > gpurun_out/r2b_k1_dyn.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_gpu_tests.log 2>&1; echo rc=$? >> gpurun_out/r2b_gpu_tests.log
timeout 300 $TR tools/dist_check.py 40 3 > gpurun_out/r2b_dist_check40.log 2>&1; echo rc=$? >> gpurun_out/r2b_dist_check40.log
TEFEM_PROFILE_ALL=1 timeout 600 python bench.py --cells 150 --steps 3 --warmup 2 --e2e-steps 0 --cpu-budget 0 > gpurun_out/r2b_bench150_1gpu.json 2> gpurun_out/r2b_bench150_1gpu.err; echo rc=$? >> gpurun_out/r2b_bench150_1gpu.err
TEFEM_PROFILE_ALL=1 timeout 600 $TR bench.py --gpus 2 --cells 150 --steps 3 --warmup 2 --e2e-steps 0 --cpu-budget 0 > gpurun_out/r2b_bench150_2gpu.json 2> gpurun_out/r2b_bench150_2gpu.err; echo rc=$? >> gpurun_out/r2b_bench150_2gpu.err
cat gpurun_out/r2b_k1_dyn.log; tail -4 gpurun_out/r2b_gpu_tests.log; tail -8 gpurun_out/r2b_dist_check40.log; grep -h "tefem profile\|rc=" gpurun_out/r2b_bench150_*.err | tail -8
