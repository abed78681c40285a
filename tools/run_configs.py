"""BASELINE.json configs 3 and 5 on one B200 (prints one JSON line per case).

config 3: synthetic refined sandwich, 200 x 50 x 33 cells x 6 = 1 980 000 tets (1 394 136 DOF), config-1 physics,
          steady state: assembly + Dirichlet + solve, checked by the per-field TRUE residual of the unscaled system.
config 5: assembly-only sweep over cubes of n^3 cells, 6 n^3 in {1 M .. 100 M} tets: K only and M + D + K.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from thermoelasticity_fem_b200 import LinearThermoElastic, Mesh, Model, LinearSteadyState, layout  # noqa: E402
from thermoelasticity_fem_b200.synthetic import kuhn_box  # noqa: E402

PEAK = 6554.9
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]


def config3():
    t0 = time.perf_counter()
    pts, tets, tris, nodes = kuhn_box(200, 50, 33, 2.0, 0.5, 0.3, index_dtype=np.int32)
    mesh = Mesh().from_arrays(pts, tets, tris, nodes)
    m = LinearThermoElastic(rho=2300, Y=64e9, nu=0.1, k=1., c=750., alpha=4e-6, T0=20.)
    mesh.set_materials({1: m, 2: m, 3: m})
    mesh.make_elements()
    model = Model(mesh, dict_dirichlet_U={4: [('x', 0.), ('y', 0.), ('z', 0.)], 5: [('x', 0.), ('y', 0.), ('z', 0.)]},
                  dict_dirichlet_theta={4: 0., 5: 0.}, dict_surface_forces={7: np.array([0., 0., -1e9])},
                  dict_heat_flux={6: -25.})
    t1 = time.perf_counter()
    solver = LinearSteadyState(model)
    solver.solve()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    # true residual of the UNSCALED reduced system K_ff x_f = F_f, per field
    dm = mesh.device_mesh()
    x = solver.X_device.clone()
    r = torch.zeros_like(x)
    dm.planes_spmv(model._planes["K"], layout.plane_map_K(), x, r)
    b = model._F
    res = (b - r)
    free = ~model._bc()[0].bool()
    idx = torch.arange(x.numel(), device=x.device)
    out = {}
    for name, sel in (("u", free & (idx % 4 != 3)), ("theta", free & (idx % 4 == 3))):
        out["relres_" + name] = float(torch.linalg.norm(res[sel]) / torch.linalg.norm(b[sel]))
    print(json.dumps({"config": 3, "n_tets": mesh.n_elements, "n_dofs": mesh.n_dofs, "mesh_s": t1 - t0, "solve_total_s": t2 - t1,
                      "timings": solver.timings, "bicgstab": solver.solve_info, "max_abs_U": float(np.abs(solver.U).max()),
                      "T_min": float(solver.T.min()), "T_max": float(solver.T.max()), **out}))


def config5(sizes):
    from thermoelasticity_fem_b200.device import DeviceMesh
    m = LinearThermoElastic(rho=2300, Y=64e9, nu=0.1, k=1., c=1., alpha=4e-6, T0=20.)
    for n in sizes:
        pts, tets, tris, nodes = kuhn_box(n, n, n, 1.0, 1.0, 1.0, n_layers=1, index_dtype=np.int32)
        conn = tets[1]
        E, N = conn.shape[0], pts.shape[0]
        t0 = time.perf_counter()
        dm = DeviceMesh(pts, conn, np.zeros(E, dtype=np.int32), np.array([m.table_row()]))
        torch.cuda.synchronize()
        t_sym = time.perf_counter() - t0
        P = dm.schedule.n_blocks
        rec = {"config": 5, "cells": n, "n_tets": E, "n_dofs": 4 * N, "symbolic_s": t_sym}
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
            rec[ops] = {"ms": ms, "elements_per_s": E / ms * 1e3, "achieved_gbs": alg / ms / 1e6, "frac_of_measured_peak": alg / ms / 1e6 / PEAK}
            del out
        print(json.dumps(rec), flush=True)
        del dm
        torch.cuda.empty_cache()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("3", "all"):
        config3()
    if what in ("5", "all"):
        config5([int(v) for v in sys.argv[2:]] or [55, 79, 119, 171, 255])
