# round 2, final call 1 (1 GPU): all GPU tests, the default bench line, launch list of the same command under ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2z_gpu_tests.log 2>&1; echo rc=$? >> gpurun_out/r2z_gpu_tests.log
python bench.py > gpurun_out/r2z_bench_default_1gpu.json 2> gpurun_out/r2z_bench_default_1gpu.err; echo "bench rc=$?"
python bench.py --steps 1 --warmup 1 --e2e-steps 0 --cpu-budget 0 --same-config-steps 0 > gpurun_out/r2z_plain.json 2> gpurun_out/r2z_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmv|reduce_scalars|update_|coarse|restrict|prolong|dense_matvec|assemble|planes|newmark|load_axpy|form_system|block_" -s 300 -c 300 --csv --log-file gpurun_out/r2z_launches_bench190.csv python bench.py --steps 1 --warmup 1 --e2e-steps 0 --cpu-budget 0 --same-config-steps 0 > gpurun_out/r2z_ncu.log 2>&1
echo "ncu rc=$?"
grep -a "passed\|failed\|rc=" gpurun_out/r2z_gpu_tests.log | tail -3
python -c "
import json
d=json.loads(open('gpurun_out/r2z_bench_default_1gpu.json').read().strip().splitlines()[-1])
s=d['solver']
print('steps/s %.3f e2e %.3f ms/it %.3f iters %s spmv frac %.3f asm MKD %.3f K %.3f'%(d['value'], d['e2e']['value'], s['ms_per_iteration'], s['iterations'], d['roofline']['frac'], d['assembly']['MKD']['frac'], d['assembly']['K']['frac']))
print(d.get('same_config',{}).get('ratio_measured'))"
