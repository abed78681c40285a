mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --e2e-steps 0 --cpu-budget 0 --same-config-steps 0"
$B > gpurun_out/r2i_plain.json 2> gpurun_out/r2i_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"coarse|restrict|prolong|dense_matvec|reduce_scalars|update_|push|init_residual|planes_spmv|block_apply|newmark|load_axpy" -s 300 -c 600 --csv --log-file gpurun_out/r2i_launches_bench190.csv $B > gpurun_out/r2i_ncu1.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"coarse_spmv_kernel" -s 40 -c 2 -o gpurun_out/r2i_coarse_spmv $B > gpurun_out/r2i_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"spmv_kernel<" -s 30 -c 2 -o gpurun_out/r2i_spmv $B > gpurun_out/r2i_ncu3.log 2>&1
tail -2 gpurun_out/r2i_ncu1.log | cut -c1-200
