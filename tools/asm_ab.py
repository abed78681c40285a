"""A/B of K1 item layouts on the GPU box: inline (split lists) vs pieces-first, same mesh, same chunks.
Checks that the planes are BITWISE identical (same chunks => same summation order) and times both.
usage: python tools/asm_ab.py [cells] ["items,budget,chunk,pieces_first[,bank_opt[,pitch]]" ...]   (first variant = reference;
bank_opt = 1: bank-aware numbering of the staged elements, symbolic.optimize_banks; 2 / 3: optimised for the kernels with D /
without D only)
Kernel form (tool-level switch TEFEM_ASM_THREADS -> DeviceMesh.asm_threads): 256 (default) | 128 | 64 = threads per CTA (pair with items = 2 x threads); build variants: TEFEM_LIB=/path/to/lib.so."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from thermoelasticity_fem_b200 import LinearThermoElastic  # noqa: E402
from thermoelasticity_fem_b200.device import DeviceMesh  # noqa: E402
from thermoelasticity_fem_b200.symbolic import build_schedule, optimize_banks  # noqa: E402
from thermoelasticity_fem_b200.synthetic import kuhn_box  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 100
variants = sys.argv[2:] or ["512,620,8,1", "512,620,8,1,1", "512,620,8,1,2", "512,620,8,1,3", "256,310,8,0"]
pts, tets, tris, nodes = kuhn_box(cells, cells, cells, 1.0, 1.0, 1.0, n_layers=1, index_dtype=np.int32)
conn = tets[1]
mat = LinearThermoElastic(rho=2300, Y=64e9, nu=0.1, k=1., c=1., alpha=4e-6, T0=20.)
table = np.array([mat.table_row()])
E, N = conn.shape[0], pts.shape[0]
conn_d = torch.from_numpy(conn).cuda()
ref = {}
for v in variants:
    f = [int(x) for x in v.split(",")]
    items, budget, chunk, pf = f[:4]
    bank = f[4] if len(f) > 4 else 0
    pitch = bool(f[5]) if len(f) > 5 else False
    t_sym = time.perf_counter()
    S = build_schedule(conn_d, N, items_per_cta=items, elem_budget=budget, max_chunk=chunk, pieces_first=bool(pf), bank_opt=False, pitch=pitch)
    t_sym = time.perf_counter() - t_sym
    t_bank = time.perf_counter()
    if bank:
        S = optimize_banks(S, strides={1: 3, 2: 2, 3: 1}[bank])
    t_bank = time.perf_counter() - t_bank
    dm = DeviceMesh(pts, conn, np.zeros(E, dtype=np.int32), table, schedule=S)
    dm.asm_threads = int(os.environ.get("TEFEM_ASM_THREADS", "256"))      # tool-level switch (the library has a setter)
    P = S.n_blocks
    line = (f"{v:16s} n_cta={S.n_cta} items={S.n_items} max_items={S.max_cta_items} max_el={S.max_cta_elems} slots={S.max_cta_slots} "
            f"symbolic {t_sym:.2f}s bank {t_bank:.2f}s |")
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
        key = (ops, chunk)
        if key not in ref:
            ref[key] = {k: t.clone() for k, t in out.items()}
            line += " (ref)"
        else:
            same = all(torch.equal(ref[key][k], out[k]) for k in out)
            worst = max(float(((ref[key][k] - out[k]).abs().max() / ref[key][k].abs().max()).item()) for k in out)
            line += f" bitwise={'YES' if same else 'NO'} maxrel={worst:.1e}"
        del out
    print(line, flush=True)
    del dm, S
