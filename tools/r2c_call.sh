mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2c_gpu_tests.log 2>&1; echo rc=$? >> gpurun_out/r2c_gpu_tests.log
( echo "== K1 tile-synchronous, 256 threads (default)"; timeout 300 python tools/asm_ab.py 100 512,620,8,1,1;
  echo "== K1 dynamic chunks, 3 TMA stages"; TEFEM_ASM_THREADS=0 timeout 300 python tools/asm_ab.py 100 512,620,8,1,1 ) > gpurun_out/r2c_k1_dyn3.log 2>&1
B="python bench.py --cells 150 --steps 1 --warmup 1 --e2e-steps 0 --cpu-budget 0"
timeout 600 $B > gpurun_out/r2c_bench150_plain.json 2> gpurun_out/r2c_bench150_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmv_kernel|precond_fused|reduce_scalars|update_|planes_spmv|newmark|init_residual|coarse|restrict|prolong|dense_matvec|block_apply|load_axpy" -c 900 --csv --log-file gpurun_out/r2c_launches_bench150.csv $B > gpurun_out/r2c_ncu.log 2>&1
tail -5 gpurun_out/r2c_gpu_tests.log; cat gpurun_out/r2c_k1_dyn3.log; tail -2 gpurun_out/r2c_ncu.log
