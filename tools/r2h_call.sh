mkdir -p gpurun_out
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655"
TEFEM_PROFILE_ALL=1 timeout 600 $TR8 bench.py --gpus 8 --steps 5 --warmup 3 --e2e-steps 2 > gpurun_out/r2h_bench_8gpu.json 2> gpurun_out/r2h_bench_8gpu.err; echo rc=$? >> gpurun_out/r2h_bench_8gpu.err
timeout 600 $TR8 bench.py --gpus 8 --steps 5 --warmup 3 --e2e-steps 0 --parity-cells 0 --halo nccl > gpurun_out/r2h_bench_8gpu_ncclhalo.json 2> gpurun_out/r2h_bench_8gpu_ncclhalo.err; echo rc=$? >> gpurun_out/r2h_bench_8gpu_ncclhalo.err
grep -h "tefem profile\|rc=\|Error" gpurun_out/r2h_bench_8gpu.err | tail -12; grep -h "rc=\|Error" gpurun_out/r2h_bench_8gpu_ncclhalo.err | tail -3
python - <<'PY'
import json
for f in ('r2h_bench_8gpu','r2h_bench_8gpu_ncclhalo'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'steps/s %.3f ms/step %.1f iters %s ms/it %.3f roof %.3f share %.3f e2e %s parity %s prepare %.1f'%(d['value'],d['ms_per_step'],d['solver']['iterations'],d['solver']['ms_per_iteration'],d['roofline']['frac'],d['roofline']['share_of_step'],(d['e2e'] or {}).get('value'),(d.get('parity_vs_1gpu') or {}).get('max_rel_diff'),d['prepare_s']))
    except Exception as e: print(f,'ERR',e)
PY
