mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_gpu_tests.log 2>&1; echo rc=$? >> gpurun_out/r2g_gpu_tests.log
TEFEM_PROFILE_ALL=1 timeout 900 python bench.py > gpurun_out/r2g_bench_default_1gpu.json 2> gpurun_out/r2g_bench_default_1gpu.err; echo rc=$? >> gpurun_out/r2g_bench_default_1gpu.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2g_bench_reference.json 2> gpurun_out/r2g_bench_reference.err; echo rc=$? >> gpurun_out/r2g_bench_reference.err
tail -4 gpurun_out/r2g_gpu_tests.log; grep -h "tefem profile\|rc=" gpurun_out/r2g_bench_default_1gpu.err | tail -3; tail -c 600 gpurun_out/r2g_bench_reference.json
