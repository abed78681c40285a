"""Small single-GPU cases for compute-sanitizer (racecheck / memcheck), one gpurun call per tool:

    compute-sanitizer --tool racecheck python tools/sanitize_case.py
    compute-sanitizer --tool memcheck  python tools/sanitize_case.py

(1) __graft_entry__.smoke(): three materials, non-zero Dirichlet values, assembly + staged CG solve, checked against the oracle;
(2) K1 on a 32^3 box: 1 057 tiles on <= 296 persistent CTAs, i.e. >= 3 tiles per CTA - both TMA prefetch stages of the double
    buffer, the mbarrier phases and the parked partial sums are all exercised - M + D + K and K-only launches, compared bitwise
    between two launches;
(3) one coupled BiCGSTAB solve with the aggregation preconditioner on a 30 x 10 x 8 box (coarse kernels, fused reductions,
    cooperative path not taken) and one cooperative BiCGSTAB solve without it.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402
from thermoelasticity_fem_b200 import LinearSteadyState, LinearThermoElastic, Mesh, Model  # noqa: E402
from thermoelasticity_fem_b200.device import DeviceMesh  # noqa: E402
from thermoelasticity_fem_b200.synthetic import kuhn_box  # noqa: E402


def main():
    __graft_entry__.smoke()
    cells = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    pts, tets, tris, nodes = kuhn_box(cells, cells, cells, 1.0, 1.0, 1.0, n_layers=1, index_dtype=np.int32)
    conn = tets[1]
    table = np.array([LinearThermoElastic(rho=2300, Y=64e9, nu=0.1, k=1., c=1., alpha=4e-6, T0=20.).table_row()])
    dm = DeviceMesh(pts, conn, np.zeros(conn.shape[0], dtype=np.int32), table)
    print("K1: tiles", dm.schedule.n_cta, "-> tiles per CTA >=", dm.schedule.n_cta // 296)
    for ops in ("MKD", "K"):
        a = {k: v.clone() for k, v in dm.assemble(ops, 0.2, 0.2).items()}
        b = dm.assemble(ops, 0.2, 0.2)
        assert all(torch.equal(a[k], b[k]) for k in a), ops
    pts, tets, tris, nodes = kuhn_box(30, 10, 8)
    mesh = Mesh().from_arrays(pts, tets, tris, nodes)
    mats = {1: (2300, 64e9, 0.1, 1., 750., 4e-6, 20.), 2: (7800, 210e9, 0.3, 45., 460., 1.2e-5, 25.), 3: (2700, 70e9, 0.33, 200., 900., 2.3e-5, 30.)}
    mesh.set_materials({k: LinearThermoElastic(*v) for k, v in mats.items()})
    mesh.make_elements()
    model = Model(mesh, dict_dirichlet_U={4: [('x', 0.), ('y', 0.), ('z', 0.)], 5: [('x', 0.), ('y', 0.), ('z', 0.)]},
                  dict_dirichlet_theta={4: 0., 5: 0.}, dict_surface_forces={7: np.array([0., 0., -1e9])}, dict_heat_flux={6: -25.})
    for kind in ("multilevel", "block_jacobi"):
        s = LinearSteadyState(model, precond=kind, solver="coupled")
        s.solve()
        print(kind, s.solve_info)
    torch.cuda.synchronize()
    print("SANITIZE_CASE DONE")


if __name__ == "__main__":
    main()
