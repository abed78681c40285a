"""BASELINE.json config 5 on N GPUs (torchrun): assembly-only sweep over cubes of n^3 Kuhn cells, x-slabs of node planes per rank
(every rank generates and assembles only the cells that touch its planes: no communication in assembly).  One JSON line per
size on rank 0: aggregate elements/s (all elements of the mesh / the slowest rank's device time) and fraction of the HBM
roofline, K only and M + D + K.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29641 \
        tools/run_configs_dist.py 79 171 255
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from thermoelasticity_fem_b200 import LinearThermoElastic  # noqa: E402
from thermoelasticity_fem_b200.device import DeviceMesh  # noqa: E402
from thermoelasticity_fem_b200.partition import localize, slab_ranges  # noqa: E402
from thermoelasticity_fem_b200.synthetic import kuhn_node_coords, kuhn_slab_tets  # noqa: E402

PEAK = 6554.9
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    m = LinearThermoElastic(rho=2300, Y=64e9, nu=0.1, k=1., c=1., alpha=4e-6, T0=20.)
    for n in [int(v) for v in sys.argv[1:]] or [79, 171, 255]:
        plane = (n + 1) * (n + 1)
        bounds = slab_ranges(n + 1, plane, world)
        p0, p1 = int(bounds[rank] // plane), int(bounds[rank + 1] // plane)
        t0 = time.perf_counter()
        conn_g, elem_ids = kuhn_slab_tets(n, n, n, p0 - 1, p1)
        lp = localize(conn_g, elem_ids, bounds, rank)
        coords = kuhn_node_coords(lp.local_nodes, n, n, n, 1.0, 1.0, 1.0)
        dm = DeviceMesh(coords, lp.conn, np.zeros(lp.conn.shape[0], dtype=np.int32), np.array([m.table_row()]), n_rows=lp.n_own)
        torch.cuda.synchronize()
        t_sym = time.perf_counter() - t0
        E_total, N_total = 6 * n ** 3, (n + 1) ** 3
        P_loc, E_loc, N_loc = dm.schedule.n_blocks, int(lp.conn.shape[0]), int(lp.n_local)
        rec = {"config": 5, "n_gpus": world, "cells": n, "n_tets": E_total, "n_dofs": 4 * N_total,
               "symbolic_s_max": None, "tets_assembled_per_rank": None}
        stats = torch.tensor([t_sym, float(E_loc)], dtype=torch.float64, device="cuda")
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        rec["symbolic_s_max"], rec["tets_assembled_per_rank"] = float(stats[0]), int(stats[1])
        for ops, nnz in (("K", 13), ("MKD", 29)):
            out = dm.assemble(ops, 0.2, 0.2)
            for _ in range(3):
                dm.assemble(ops, 0.2, 0.2, out=out, check=False)
            torch.cuda.synchronize()
            dist.barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(10):
                dm.assemble(ops, 0.2, 0.2, out=out, check=False)
            ev1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([ev0.elapsed_time(ev1) / 10], dtype=torch.float64, device="cuda")
            alg = torch.tensor([float(8 * nnz * P_loc + 20 * E_loc + 24 * N_loc + 64 * E_loc)], dtype=torch.float64, device="cuda")
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(alg, op=dist.ReduceOp.SUM)
            ms, alg = float(ms.item()), float(alg.item())
            rec[ops] = {"ms_max_over_ranks": ms, "elements_per_s": E_total / ms * 1e3, "achieved_gbs_per_gpu": alg / ms / 1e6 / world,
                        "frac_of_measured_peak": alg / ms / 1e6 / world / PEAK}
            del out
        if rank == 0:
            print(json.dumps(rec), flush=True)
        del dm
        torch.cuda.empty_cache()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
