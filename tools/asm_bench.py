"""Assembly micro-benchmark (GPU box): times K1 for schedule variants on an n^3 Kuhn box.
usage: python tools/asm_bench.py [cells] ["items,budget,chunk,mirror" ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from thermoelasticity_fem_b200 import LinearThermoElastic  # noqa: E402
from thermoelasticity_fem_b200.device import DeviceMesh  # noqa: E402
from thermoelasticity_fem_b200.symbolic import build_schedule  # noqa: E402
from thermoelasticity_fem_b200.synthetic import kuhn_box  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 100
variants = sys.argv[2:] or ["256,600,6"]
pts, tets, tris, nodes = kuhn_box(cells, cells, cells, 1.0, 1.0, 1.0, n_layers=1, index_dtype=np.int32)
conn = tets[1]
mat = LinearThermoElastic(rho=2300, Y=64e9, nu=0.1, k=1., c=1., alpha=4e-6, T0=20.)
table = np.array([mat.table_row()])
E, N = conn.shape[0], pts.shape[0]
conn_d = torch.from_numpy(conn).cuda()
for v in variants:
    items, budget, chunk = [int(x) for x in v.split(",")][:3]
    window = int(v.split(",")[3]) if len(v.split(",")) > 3 else None
    S = build_schedule(conn_d, N, items_per_cta=items, elem_budget=budget, max_chunk=chunk, item_sort_window=window)
    dm = DeviceMesh(pts, conn, np.zeros(E, dtype=np.int32), table, schedule=S)
    P = S.n_blocks
    line = f"{v:16s} n_cta={S.n_cta} items={S.n_items} max_el={S.max_cta_elems} max_nodes={S.max_cta_nodes} slots={S.max_cta_slots} sched={S.schedule_bytes() / E:.0f} B/elem |"
    for ops, nnz in (("K", 13), ("MKD", 29)):
        out = dm.assemble(ops, 0.2, 0.2)
        for _ in range(3):
            dm.assemble(ops, 0.2, 0.2, out=out, check=False)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(10):
            dm.assemble(ops, 0.2, 0.2, out=out, check=False)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 10
        alg = 8 * nnz * P + 20 * E + 24 * N + 64 * E
        line += f" {ops}: {ms:.3f} ms {E / ms / 1e6:.2f} Gel/s {alg / ms / 1e6 / 6554.9 * 100:.1f}% |"
        del out
    print(line, flush=True)
    del dm, S
