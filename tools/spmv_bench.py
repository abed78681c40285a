"""SpMV micro-benchmark (GPU box): GB/s of the 4x4-block SpMV on an n^3 Kuhn box.
usage: python tools/spmv_bench.py [cells] [threads,ctas_per_sm,unroll]   (launch configuration, tefem_spmv_set_config)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from thermoelasticity_fem_b200.device import DeviceMesh  # noqa: E402
from thermoelasticity_fem_b200.symbolic import build_schedule  # noqa: E402
from thermoelasticity_fem_b200.synthetic import kuhn_box  # noqa: E402
from thermoelasticity_fem_b200 import LinearThermoElastic  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 150
cfg = sys.argv[2] if len(sys.argv) > 2 else "default"
pts, tets, tris, nodes = kuhn_box(cells, cells, cells, 1.0, 1.0, 1.0, n_layers=1, index_dtype=np.int32)
conn = tets[1]
E, N = conn.shape[0], pts.shape[0]
mat = LinearThermoElastic(rho=2300, Y=64e9, nu=0.1, k=1., c=1., alpha=4e-6, T0=20.)
dm = DeviceMesh(pts, conn, np.zeros(E, dtype=np.int32), np.array([mat.table_row()]))
if cfg != "default":
    from thermoelasticity_fem_b200 import _lib
    _lib.check(dm.lib.tefem_spmv_set_config(*[int(v) for v in cfg.split(",")]))
P = dm.schedule.n_blocks
sys_ = torch.randn((P, 16), dtype=torch.float64, device="cuda")
x = torch.randn(4 * N, dtype=torch.float64, device="cuda")
y = torch.zeros_like(x)
for _ in range(5):
    dm.spmv(sys_, x, y)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
reps = 30
for _ in range(reps):
    dm.spmv(sys_, x, y)
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / reps
alg = 8 * 16 * P + 4 * P + 4 * (N + 1) + 16 * 4 * N
print(f"cfg={cfg} cells={cells} P={P} {ms:.4f} ms {alg / ms / 1e6:.0f} GB/s = {alg / ms / 1e6 / 6554.9 * 100:.1f}% of measured peak")
