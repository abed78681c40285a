mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655"
B="bench.py --cells 150 --steps 3 --warmup 2 --e2e-steps 0 --cpu-budget 0"
for F in 1 0; do
  TEFEM_PRECOND_FUSED=$F TEFEM_PROFILE_ALL=1 timeout 600 python $B > gpurun_out/r2d_bench150_1gpu_fused$F.json 2> gpurun_out/r2d_bench150_1gpu_fused$F.err; echo rc=$? >> gpurun_out/r2d_bench150_1gpu_fused$F.err
done
timeout 300 $TR tools/dist_check.py 40 3 > gpurun_out/r2d_dist_check40.log 2>&1; echo rc=$? >> gpurun_out/r2d_dist_check40.log
for F in 1 0; do
  TEFEM_PRECOND_FUSED=$F TEFEM_PROFILE_ALL=1 timeout 600 $TR $B --gpus 2 > gpurun_out/r2d_bench150_2gpu_fused$F.json 2> gpurun_out/r2d_bench150_2gpu_fused$F.err; echo rc=$? >> gpurun_out/r2d_bench150_2gpu_fused$F.err
done
grep -h "tefem profile\|rc=" gpurun_out/r2d_bench150_*.err; tail -8 gpurun_out/r2d_dist_check40.log
