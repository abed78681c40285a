"""BASELINE config 5: assembly-only throughput sweep (K1) on n^3 Kuhn boxes, 1 GPU or row-partitioned over the ranks of a
torchrun job (x-slabs, interface elements duplicated: no communication).  One JSON line per (n, operators) on rank 0.

    python tools/asm_sweep.py 55 79 119 171 255
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29671 tools/asm_sweep.py 119 171 255
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch.distributed as dist
    cells_list = [int(a) for a in sys.argv[1:]] or [55, 79, 119, 171, 255]
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    peak, _ = bench.peaks()

    def reduce(v, op):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
        return float(t.item())

    for cells in cells_list:
        t0 = time.perf_counter()
        mesh, model, solver = bench.build_case(cells, rank, world)
        dm = mesh.device_mesh()
        S = dm.schedule
        t_setup = reduce(time.perf_counter() - t0, "MAX")
        E_total = 6 * cells ** 3
        for ops, planes in (("MKD", 29), ("K", 13)):
            out = dm.assemble(ops, bench.ALPHA_M, bench.ALPHA_K)
            for _ in range(3):
                dm.assemble(ops, bench.ALPHA_M, bench.ALPHA_K, out=out, check=False)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                dm.assemble(ops, bench.ALPHA_M, bench.ALPHA_K, out=out, check=False)
            e1.record()
            torch.cuda.synchronize()
            ms = reduce(e0.elapsed_time(e1) / 5, "MAX")
            alg_local = 8 * planes * S.n_blocks + 20 * dm.n_elems + 24 * dm.n_nodes + 64 * dm.n_elems
            frac = reduce(alg_local / ms / 1e6 / peak, "MIN")                                                # slowest rank
            alg = reduce(float(alg_local), "SUM")
            if rank == 0:
                print(json.dumps({"config": 5, "cells": cells, "n_tets": E_total, "n_gpus": world, "ops": ops, "ms": ms,
                                  "elements_per_s": E_total / (ms * 1e-3), "algorithmic_bytes_all_ranks": alg,
                                  "frac_of_measured_hbm_per_gpu_min": frac, "setup_s": t_setup,
                                  "duplicated_interface_elements": reduce(float(dm.n_elems), "SUM") - E_total if world > 1 else 0}),
                      flush=True)
            else:
                reduce(float(dm.n_elems), "SUM")
            del out
        dm.check_jacobian()
        del mesh, model, solver, dm
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
