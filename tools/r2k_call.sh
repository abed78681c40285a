mkdir -p gpurun_out
python tools/run_configs.py all > gpurun_out/r2_configs_3_5.jsonl 2> gpurun_out/r2k_configs.err; echo rc=$? >> gpurun_out/r2k_configs.err
timeout 900 python bench.py > gpurun_out/r2k_bench_default_1gpu.json 2> gpurun_out/r2k_bench_default_1gpu.err; echo rc=$? >> gpurun_out/r2k_bench_default_1gpu.err
B="python bench.py --steps 1 --warmup 1 --e2e-steps 0 --cpu-budget 0 --same-config-steps 0"
$B > gpurun_out/r2k_plain.json 2> gpurun_out/r2k_plain.err && \
timeout 900 ncu --set full --clock-control none -k regex:'^spmv_kernel$' -s 30 -c 4 -o gpurun_out/r2k_spmv $B > gpurun_out/r2k_ncu.log 2>&1
tail -3 gpurun_out/r2k_ncu.log | cut -c1-200; cat gpurun_out/r2_configs_3_5.jsonl | cut -c1-400; tail -2 gpurun_out/r2k_bench_default_1gpu.err
