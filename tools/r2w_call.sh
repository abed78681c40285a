# round 2, call w (2 GPUs): time line of ALL CTAs of the dataflow coarse cycle (TEFEM_PROFILE_PRECOND=1)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631"
B="$T bench.py --gpus 2 --cells 120 --steps 2 --warmup 1 --e2e-steps 0 --cpu-budget 0 --same-config-steps 0 --parity-cells 0"
TEFEM_PROFILE_PRECOND=1 timeout 600 $B > gpurun_out/r2w_phases.json 2> gpurun_out/r2w_phases.err; echo rc=$?
grep -a "tefem precond" gpurun_out/r2w_phases.err | cut -c1-300
