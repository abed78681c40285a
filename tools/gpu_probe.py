"""Development probe (GPU box): solver behaviour and kernel timings on the fixture mesh and Kuhn boxes."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from thermoelasticity_fem_b200 import LinearThermoElastic, Mesh, Model, LinearSteadyState, LinearTransient  # noqa: E402
from thermoelasticity_fem_b200.synthetic import kuhn_box  # noqa: E402


def timed(fn, n=5):
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(n):
        fn()
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / n


def transient_case(mesh, n_steps, n_t=1000, **kw):
    t_end = 1800e0
    vec_t = np.linspace(0., t_end, n_t)
    arr = np.zeros((3, n_t))
    arr[0, :] = np.sin(2 * np.pi * vec_t / 900.) * -1e9
    model = Model(mesh, dict_dirichlet_U={4: [('x', 0.), ('y', 0.), ('z', 0.)], 5: [('y', 0.), ('z', 0.)]},
                  dict_dirichlet_theta={4: 0., 5: 0.}, dict_surface_forces={5: arr}, alpha_M=0.2, alpha_K=0.2)
    N = mesh.n_nodes
    s = LinearTransient(model, np.zeros(3 * N), np.zeros(3 * N), np.zeros(N), np.zeros(N), t_end, n_t, 0.5, 0.25,
                        progress=False, **kw)
    t0 = time.perf_counter()
    s.prepare()
    torch.cuda.synchronize()
    print("  prepare %.3f s" % (time.perf_counter() - t0))
    for i in range(1, n_steps + 1):
        t0 = time.perf_counter()
        info = s.step(i)
        torch.cuda.synchronize()
        print("  step %d: %.3f s  %s" % (i, time.perf_counter() - t0, info))
    return s


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "sandwich"
    mat = LinearThermoElastic(rho=2300, Y=64e9, nu=0.1, k=1., c=1., alpha=4e-6, T0=20.)
    if what == "sandwich":
        from conftest import load_sandwich_blocks
        pts, blocks = load_sandwich_blocks()
        mesh = Mesh()
        mesh.from_blocks(pts, blocks)
    else:
        nx, ny, nz = (int(v) for v in what.split("x"))
        t0 = time.perf_counter()
        lx, ly, lz = (2.0, 0.5, 0.3) if len(sys.argv) < 3 else (1.0, 1.0, 1.0)
        pts, tets, tris, nodes = kuhn_box(nx, ny, nz, lx, ly, lz)
        mesh = Mesh().from_arrays(pts, tets, tris, nodes)
        print("mesh gen %.2f s: %d nodes %d tets" % (time.perf_counter() - t0, pts.shape[0], sum(len(v) for v in tets.values())))
    mesh.set_materials({1: mat, 2: mat, 3: mat})
    mesh.make_elements()
    t0 = time.perf_counter()
    dm = mesh.device_mesh()
    torch.cuda.synchronize()
    S = dm.schedule
    print("symbolic %.2f s: P=%d n_cta=%d max_cta_elems=%d" % (time.perf_counter() - t0, S.n_blocks, S.n_cta, S.max_cta_elems))
    model = Model(mesh, alpha_M=0.2, alpha_K=0.2)
    E, P, N = mesh.n_elements, S.n_blocks, mesh.n_nodes
    for ops, nnz in (("K", 13), ("MKD", 29)):
        out = dm.assemble(ops, 0.2, 0.2)
        ms = timed(lambda: dm.assemble(ops, 0.2, 0.2, out=out, check=False))
        alg = 8 * nnz * P + 20 * E + 24 * N + 64 * E
        print("assemble %-3s: %.3f ms  %.3f Gelem/s  algorithmic %.1f MB -> %.0f GB/s" % (ops, ms, E / ms / 1e6, alg / 1e6, alg / ms / 1e6))
    model._planes.update(dm.assemble("MKD", 0.2, 0.2))
    model._d_alpha = (0.2, 0.2)
    model.create_free_dofs_lists()
    sys_ = dm.form_system(model._planes, 1.0, 0.9, 0.81, 13, model._dir_mask)
    x = torch.randn(4 * N, dtype=torch.float64, device="cuda")
    y = torch.zeros_like(x)
    ms = timed(lambda: dm.spmv(sys_, x, y), 20)
    alg = 8 * 16 * P + 4 * P + 4 * (N + 1) + 16 * 4 * N
    print("spmv: %.3f ms  algorithmic %.1f MB -> %.0f GB/s" % (ms, alg / 1e6, alg / ms / 1e6))
    if what == "sandwich":
        model = Model(mesh, dict_dirichlet_U={4: [('x', 0.), ('y', 0.), ('z', 0.)], 5: [('x', 0.), ('y', 0.), ('z', 0.)]},
                      dict_dirichlet_theta={4: 0., 5: 0.}, dict_surface_forces={7: np.array([0., 0., -1e9])},
                      dict_heat_flux={6: -25.})
        st = LinearSteadyState(model)
        st.solve()
        print("steady:", st.solve_info, st.timings)
    n_steps = int(os.environ.get("PROBE_STEPS", "5"))
    transient_case(mesh, n_steps)


if __name__ == "__main__":
    main()
