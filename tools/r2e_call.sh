mkdir -p gpurun_out
B="python bench.py --cells 150 --steps 2 --warmup 1 --e2e-steps 0 --cpu-budget 0"
run() { tag=$1; shift; env "$@" TEFEM_PROFILE_ALL=1 timeout 600 $B > gpurun_out/r2e_$tag.json 2> gpurun_out/r2e_$tag.err; echo "$tag rc=$?" >> gpurun_out/r2e_summary.log; grep -h "tefem profile" gpurun_out/r2e_$tag.err >> gpurun_out/r2e_summary.log; python -c "
import json,sys
d=json.loads(open('gpurun_out/r2e_$tag.json').read().strip().splitlines()[-1])
print('   steps/s %.3f ms/it %.3f iters %s'%(d['value'], d['solver']['ms_per_iteration'], d['solver']['iterations']))" >> gpurun_out/r2e_summary.log 2>&1; }
rm -f gpurun_out/r2e_summary.log
run base A=1
run fp16 TEFEM_ML_ROUND=fp16
run bf16 TEFEM_ML_ROUND=bf16
run sw22 TEFEM_ML=8,3,1.0,0.7,2,2
run sw11 TEFEM_ML=8,3,1.0,0.7,1,1
run sw22fp16 TEFEM_ML=8,3,1.0,0.7,2,2 TEFEM_ML_ROUND=fp16
$B > gpurun_out/r2e_plain.json 2> gpurun_out/r2e_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmv_kernel|precond_fused|reduce_scalars|update_|coarse|restrict|prolong|dense_matvec" -s 400 -c 400 --csv --log-file gpurun_out/r2e_launches_bench150.csv $B > gpurun_out/r2e_ncu.log 2>&1
cat gpurun_out/r2e_summary.log
