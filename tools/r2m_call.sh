# round 2, call m: the persistent row-partitioned coarse cycle on ONE GPU (world = 1): GPU tests, then the bench with
# both coarse modes (per-call profile lines on stderr)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_gpu_tests.log 2>&1; echo rc=$? >> gpurun_out/r2m_gpu_tests.log
B="python bench.py --steps 3 --warmup 2 --e2e-steps 0 --cpu-budget 0 --same-config-steps 0"
rm -f gpurun_out/r2m_summary.log
run() { tag=$1; shift; timeout 600 $B "$@" > gpurun_out/r2m_$tag.json 2> gpurun_out/r2m_$tag.err; echo "$tag rc=$?" >> gpurun_out/r2m_summary.log; grep -ah "tefem profile\|tefem precond" gpurun_out/r2m_$tag.err >> gpurun_out/r2m_summary.log; python -c "
import json
d=json.loads(open('gpurun_out/r2m_$tag.json').read().strip().splitlines()[-1])
print('   steps/s %.3f ms/it %.3f solve/it %.3f iters %s'%(d['value'], d['solver']['ms_per_iteration'], d['solver']['solve_ms_per_iteration'], d['solver']['iterations']))" >> gpurun_out/r2m_summary.log 2>&1; }
run launches --ml coarse=launches
run persistent --ml coarse=persistent
tail -3 gpurun_out/r2m_gpu_tests.log; cat gpurun_out/r2m_summary.log
