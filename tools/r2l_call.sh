mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2l_gpu_tests.log 2>&1; echo rc=$? >> gpurun_out/r2l_gpu_tests.log
B="python bench.py --steps 3 --warmup 2 --e2e-steps 0 --cpu-budget 0 --same-config-steps 0"
rm -f gpurun_out/r2l_summary.log
run() { tag=$1; shift; timeout 600 $B "$@" > gpurun_out/r2l_$tag.json 2> gpurun_out/r2l_$tag.err; echo "$tag rc=$?" >> gpurun_out/r2l_summary.log; python -c "
import json
d=json.loads(open('gpurun_out/r2l_$tag.json').read().strip().splitlines()[-1])
print('   steps/s %.3f ms/it %.3f solve/it %.3f iters %s'%(d['value'], d['solver']['ms_per_iteration'], d['solver']['solve_ms_per_iteration'], d['solver']['iterations']))" >> gpurun_out/r2l_summary.log 2>&1; }
run overlapped33
run classic33 --ml cycle=classic
run overlapped32 --ml n_post=2
run overlapped33_fp32 --ml level1_dtype=fp32
grep -a "mat_K stored\|passed\|failed\|rc=" gpurun_out/r2l_gpu_tests.log | tail -4; cat gpurun_out/r2l_summary.log
