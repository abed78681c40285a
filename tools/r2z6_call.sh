# round 2, final call 6 (1 GPU): the opt-in MIRROR layout of K1: all GPU tests (incl. mirror vs direct planes), then the A/B at 100^3
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2z6_gpu_tests.log 2>&1; echo rc=$? >> gpurun_out/r2z6_gpu_tests.log
timeout 300 python tools/asm_ab.py 100 512,620,8,1,1 512,620,8,1,1,0,1 480,620,8,1,1,0,1 > gpurun_out/r2z6_asm_ab_mirror.log 2>&1; echo rc=$? >> gpurun_out/r2z6_asm_ab_mirror.log
tail -3 gpurun_out/r2z6_gpu_tests.log; cat gpurun_out/r2z6_asm_ab_mirror.log | cut -c1-400
