"""Quick GPU check of the thread-count variants of K1 in ONE process: planes bitwise equal to the default, rough timing.
usage: python tools/asm_threads_check.py [cells]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from thermoelasticity_fem_b200 import LinearThermoElastic  # noqa: E402
from thermoelasticity_fem_b200.device import DeviceMesh  # noqa: E402
from thermoelasticity_fem_b200.symbolic import build_schedule  # noqa: E402
from thermoelasticity_fem_b200.synthetic import kuhn_box  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 60
pts, tets, tris, nodes = kuhn_box(cells, cells, cells, 1.0, 1.0, 1.0, n_layers=1, index_dtype=np.int32)
conn = tets[1]
table = np.array([LinearThermoElastic(rho=2300, Y=64e9, nu=0.1, k=1., c=1., alpha=4e-6, T0=20.).table_row()])
E = conn.shape[0]
conn_d = torch.from_numpy(conn).cuda()
ref = None
for threads, items in ((256, 512), (128, 256), (64, 128), (128, 384), (64, 192)):
    S = build_schedule(conn_d, pts.shape[0], items_per_cta=items, elem_budget=int(1.21 * items))
    dm = DeviceMesh(pts, conn, np.zeros(E, dtype=np.int32), table, schedule=S)
    dm.asm_threads = threads
    line = f"threads {threads:3d} tile {items:3d}:"
    for ops in ("K", "MKD"):
        out = dm.assemble(ops, 0.2, 0.2)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(5):
            dm.assemble(ops, 0.2, 0.2, out=out, check=False)
        ev1.record()
        torch.cuda.synchronize()
        line += f" {ops} {ev0.elapsed_time(ev1) / 5:.3f} ms"
        if ops == "MKD":
            if ref is None:
                ref = {k: v.clone() for k, v in out.items()}
            else:
                line += " bitwise=" + ("YES" if all(torch.equal(ref[k], out[k]) for k in out) else "NO")
    print(line, flush=True)
