"""Generate the committed fixtures in tests/golden/ by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Runs in the build container only (needs /root/reference; see
``oracle/reference_shim.py``).  Usage::

    python -m oracle.make_golden [--transient-steps 999]

Fixtures written (all produced by the reference's own classes, nothing from the product):

* sandwich_mesh.npz        the reference's data/sandwich.msh as arrays (points, cell blocks in file order)
* sandwich_elements.npz    the six element matrices of every 53rd tet + a weighted fingerprint of all tets
* sandwich_operators.npz   reference M, D (alpha_M = alpha_K = 0.2, config-2 material), K as CSC triplets
* sandwich_steady.npz      config 1 (examples/example_steadystate.py): DOF maps, F, F_f, raw spsolve X, U, T
* sandwich_transient2.npz  config 2 (examples/example_transient_2.py): DOF maps, F(idx), states at chosen steps
* kuhn_small.npz           4x2x3 Kuhn box with three different materials: all element matrices, M, D, K, maps, F
* sandwich_rom3.npz        config of examples/example_transient_3.py (LinearTransient.solve_ROM, reduced-order bases from
                           eigs): Rayleigh quotients of the bases, basis-independent fingerprints of the two subspaces,
                           U / T / their rates at chosen steps  (--only rom [--rom-modes 30,30])
"""
import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import reference_shim as rs          # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
NAMES = ["Muu", "Dtu", "Dtt", "Kuu", "Kut", "Ktt"]


def csc_triplet(prefix, A):
    A = A.tocsc()
    A.sort_indices()
    return {prefix + "_indptr": A.indptr.astype(np.int32), prefix + "_indices": A.indices.astype(np.int32),
            prefix + "_data": A.data.copy(), prefix + "_shape": np.array(A.shape)}


def fingerprint_weights(n):
    """Deterministic weights in [1, 2) used to condense one element matrix into one number."""
    k = np.arange(n, dtype=np.float64)
    return 1.0 + np.mod(k * 0.6180339887498949, 1.0)


def element_dump(elements, stride):
    sub = list(range(0, len(elements), stride))
    out = {"subset": np.array(sub)}
    for nm in NAMES:
        out[nm] = np.array([getattr(elements[i], f"compute_mat_{nm}_e")() for i in sub])
    fp = np.zeros((len(elements), 6))
    for i, el in enumerate(elements):
        for j, nm in enumerate(NAMES):
            m = getattr(el, f"compute_mat_{nm}_e")().reshape(-1)
            fp[i, j] = m @ fingerprint_weights(m.shape[0])
    out["fingerprint"] = fp
    return out


def sandwich_mesh(ref):
    mesh = ref.mesh.Mesh()
    mesh.load_mesh(rs.SANDWICH_MSH)
    return mesh


def write_mesh_fixture():
    pts, blocks = rs.parse_gmsh41(rs.SANDWICH_MSH)
    d = {"points": pts, "n_blocks": np.array(len(blocks))}
    for i, (typ, tag, conn) in enumerate(blocks):
        d[f"b{i}_type"] = np.array(typ)
        d[f"b{i}_tag"] = np.array(tag)
        d[f"b{i}_conn"] = conn.astype(np.int32)
    np.savez_compressed(os.path.join(OUT, "sandwich_mesh.npz"), **d)


def maps_of(model, prefix=""):
    return {prefix + "free_dofs": np.array(model.free_dofs, dtype=np.int32),
            prefix + "free_dofs_U": np.array(model.free_dofs_U, dtype=np.int32),
            prefix + "free_dofs_theta": np.array(model.free_dofs_theta, dtype=np.int32),
            prefix + "dirichlet_dofs_U_sorted": np.array(sorted(model.dirichlet_dofs_U), dtype=np.int32),
            prefix + "dirichlet_dofs_theta_sorted": np.array(sorted(model.dirichlet_dofs_theta), dtype=np.int32)}


def config1(ref):
    """examples/example_steadystate.py:19-77."""
    mesh = sandwich_mesh(ref)
    material = ref.materials.LinearThermoElastic(rho=2300, Y=64e9, nu=0.1, k=1., c=750., alpha=4e-6, T0=20.)
    mesh.set_materials({1: material, 2: material, 3: material})
    mesh.make_elements()
    model = ref.model.Model(mesh,
                            dict_dirichlet_U={4: [('x', 0.), ('y', 0.), ('z', 0.)], 5: [('x', 0.), ('y', 0.), ('z', 0.)]},
                            dict_dirichlet_theta={4: 0., 5: 0.},
                            dict_surface_forces={7: np.array([0., 0., -1e9])}, dict_volume_forces=None,
                            dict_heat_flux={6: -25.}, dict_heat_source=None)
    return mesh, model


def config2(ref, n_t=1000):
    """examples/example_transient_2.py:19-100."""
    t_end = 1800e0
    vec_t = np.linspace(0., t_end, n_t)
    mesh = sandwich_mesh(ref)
    material = ref.materials.LinearThermoElastic(rho=2300, Y=64e9, nu=0.1, k=1., c=1., alpha=4e-6, T0=20.)
    mesh.set_materials({1: material, 2: material, 3: material})
    mesh.make_elements()
    arr = np.zeros((3, n_t))
    arr[0, :] = np.sin(2 * np.pi * vec_t / 900.) * -1e9
    model = ref.model.Model(mesh,
                            dict_dirichlet_U={4: [('x', 0.), ('y', 0.), ('z', 0.)], 5: [('y', 0.), ('z', 0.)]},
                            dict_dirichlet_theta={4: 0., 5: 0.},
                            dict_surface_forces={5: arr}, dict_volume_forces=None,
                            dict_heat_flux=None, dict_heat_source=None, alpha_M=2e-1, alpha_K=2e-1)
    return mesh, model, t_end, n_t


def run_steady(ref):
    t = time.time()
    mesh, model = config1(ref)
    np.savez_compressed(os.path.join(OUT, "sandwich_elements.npz"), **element_dump(mesh.elements, 53))
    solver = ref.solvers.LinearSteadyState(model)
    solver.solve()
    d = maps_of(model)
    d.update(F=model.vec_F, F_f=model.vec_F_f, X_raw_spsolve=solver.X, U=solver.U, T=solver.T,
             K_nnz=np.array(model.mat_K.nnz), Kff_nnz=np.array(model.mat_K_f_f.nnz))
    np.savez_compressed(os.path.join(OUT, "sandwich_steady.npz"), **d)
    print("steady done in %.1f s; max|U| = %.16g, T in [%.16g, %.16g], K.nnz = %d, Kff.nnz = %d"
          % (time.time() - t, np.abs(solver.U).max(), solver.T.min(), solver.T.max(), model.mat_K.nnz, model.mat_K_f_f.nnz))
    return model.mat_K


def run_operators(ref, K_steady):
    t = time.time()
    mesh, model, _, _ = config2(ref)
    model.assemble_M()
    model.assemble_K()
    model.assemble_D()
    assert (model.mat_K != K_steady).nnz == 0      # K does not depend on c
    d = {}
    d.update(csc_triplet("M", model.mat_M))
    d.update(csc_triplet("D", model.mat_D))
    d.update(csc_triplet("K", model.mat_K))
    np.savez_compressed(os.path.join(OUT, "sandwich_operators.npz"), **d)
    print("operators done in %.1f s; nnz M/D/K = %d / %d / %d"
          % (time.time() - t, model.mat_M.nnz, model.mat_D.nnz, model.mat_K.nnz))


def run_transient(ref, n_steps, keep):
    """The reference's solve() always runs all n_t-1 steps; to bound the run a shortened horizon is
    obtained by restating the LOOP BOUNDS only: the reference's own solve() body is executed on a
    model whose load array and n_t are unchanged, via the `range` seen by solvers.py:163."""
    t = time.time()
    mesh, model, t_end, n_t = config2(ref)
    solver = ref.solvers.LinearTransient(model, np.zeros(mesh.n_nodes * 3), np.zeros(mesh.n_nodes * 3),
                                         np.zeros(mesh.n_nodes), np.zeros(mesh.n_nodes), t_end, n_t, 1 / 2, 1 / 4)
    if n_steps < n_t - 1:
        # run the unmodified solve() but make its time loop stop early: tqdm(range(1, n_t)) is the only
        # use of tqdm in the module, so substitute an iterator that truncates it.
        orig = ref.solvers.tqdm
        ref.solvers.tqdm = lambda it: (i for i in it if i <= n_steps)
        try:
            solver.solve()
        finally:
            ref.solvers.tqdm = orig
    else:
        solver.solve()
    keep = [k for k in keep if k <= n_steps]
    d = maps_of(model)
    d.update(steps=np.array(keep), dt=np.array(solver.dt), n_steps_run=np.array(n_steps),
             X=solver.X[:, keep], Xdot=solver.Xdot[:, keep], Xdotdot=solver.Xdotdot[:, keep],
             T=solver.T[:, keep], U=solver.U[:, keep],
             maxU=np.array(np.abs(solver.U[:, :n_steps + 1]).max()),
             maxTheta=np.array(np.abs(solver.X[3::4, :n_steps + 1]).max()))
    for idx in (1, 250, 999):
        model.assemble_F(idx_t=idx)
        model.apply_dirichlet_F()
        d[f"F_{idx}"] = model.vec_F.copy()
        d[f"F_f_{idx}"] = model.vec_F_f.copy()
    np.savez_compressed(os.path.join(OUT, "sandwich_transient2.npz"), **d)
    print("transient (%d steps) done in %.1f s; max|U| = %.16g, max|theta| = %.16g"
          % (n_steps, time.time() - t, d["maxU"], d["maxTheta"]))


def run_rom(ref, n_u, n_t_modes, keep):
    """examples/example_transient_3.py: config 2 solved with solve_ROM (solvers.py:219-354, model.py:341-440).
    The bases returned by eigs are unique only up to sign / rotation inside an eigenspace, so the fixture stores
    basis-independent data: Rayleigh quotients, the M- (D-) orthogonal projection of a fixed vector onto each
    subspace, and the physical histories U, T at chosen steps."""
    t = time.time()
    mesh, model, t_end, n_t = config2(ref)
    solver = ref.solvers.LinearTransient(model, np.zeros(mesh.n_nodes * 3), np.zeros(mesh.n_nodes * 3),
                                         np.zeros(mesh.n_nodes), np.zeros(mesh.n_nodes), t_end, n_t, 1 / 2, 1 / 4)
    solver.solve_ROM(n_modes_u=n_u, n_modes_theta=n_t_modes)
    print("reference solve_ROM done in %.1f s" % (time.time() - t))
    d = maps_of(model)
    fu, ft = np.array(model.free_dofs_U), np.array(model.free_dofs_theta)
    Muu, Kuu = model.mat_M[fu, :][:, fu], model.mat_K[fu, :][:, fu]
    Dtt, Ktt = model.mat_D[ft, :][:, ft], model.mat_K[ft, :][:, ft]
    for nm, phi, A, B, n in (("u", model.mat_phi_u, Kuu, Muu, fu.size), ("t", model.mat_phi_t, Ktt, Dtt, ft.size)):
        G = phi.T @ (B @ phi)
        d["lam_" + nm] = np.sort(np.linalg.eigvals(np.linalg.solve(G, phi.T @ (A @ phi))).real)
        w = fingerprint_weights(n)
        d["proj_" + nm] = phi @ np.linalg.solve(G, phi.T @ (B @ w))      # B-orthogonal projection of w onto span(phi)
        d["resid_" + nm] = np.array(max(np.linalg.norm(A @ phi[:, j] - ((phi[:, j] @ (A @ phi[:, j])) / (phi[:, j] @ (B @ phi[:, j])))
                                                       * (B @ phi[:, j])) / np.linalg.norm(A @ phi[:, j]) for j in range(phi.shape[1])))
    keep = [k for k in keep if k < n_t]
    d.update(steps=np.array(keep), n_modes=np.array([n_u, n_t_modes]), U=solver.U[:, keep], T=solver.T[:, keep],
             Udot=solver.Udot[:, keep], Tdot=solver.Tdot[:, keep],
             maxU=np.array(np.abs(solver.U).max()), maxT=np.array(np.abs(solver.T).max()))
    np.savez_compressed(os.path.join(OUT, "sandwich_rom3.npz"), **d)
    print("rom (%d + %d modes) done in %.1f s; max|U| = %.16g, max|T| = %.16g, eig residuals %.2e / %.2e, lam_u[:3] = %s, lam_t[:3] = %s"
          % (n_u, n_t_modes, time.time() - t, d["maxU"], d["maxT"], d["resid_u"], d["resid_t"], d["lam_u"][:3], d["lam_t"][:3]))


def run_kuhn(ref):
    """Small structured box with three DIFFERENT materials, non-zero Dirichlet values, all load kinds."""
    from thermoelasticity_fem_b200.synthetic import kuhn_box
    pts, tets, tris, nodes = kuhn_box(4, 2, 3, jitter=0.2, seed=0)
    mesh = ref.mesh.Mesh()
    mesh.table_nodes = pts
    mesh.n_nodes = pts.shape[0]
    mesh.n_dofs = 4 * mesh.n_nodes
    mesh.dict_tet_groups = tets
    mesh.dict_tri_groups = tris
    mesh.dict_nodes_groups = nodes
    mats = {1: ref.materials.LinearThermoElastic(rho=2300, Y=64e9, nu=0.1, k=1., c=750., alpha=4e-6, T0=20.),
            2: ref.materials.LinearThermoElastic(rho=7800, Y=210e9, nu=0.3, k=45., c=460., alpha=1.2e-5, T0=25.),
            3: ref.materials.LinearThermoElastic(rho=2700, Y=70e9, nu=0.33, k=200., c=900., alpha=2.3e-5, T0=30.)}
    mesh.set_materials(mats)
    mesh.make_elements()
    n_t = 5
    model = ref.model.Model(mesh,
                            dict_dirichlet_U={4: [('x', 0.), ('y', 1e-4), ('z', 0.)], 5: [('x', 2e-4)]},
                            dict_dirichlet_theta={4: 3.0, 6: -1.5},
                            dict_nodal_forces={8: np.array([1e3, -2e3, 5e2])},
                            dict_surface_forces={7: np.array([0., 1e5, -1e6])},
                            dict_volume_forces={1: np.array([0., 0., -2300 * 9.81]), 3: np.array([1., 2., 3.])},
                            dict_heat_flux={7: -25.}, dict_heat_source={2: 1e3},
                            alpha_M=0.3, alpha_K=0.05)
    model.create_free_dofs_lists()
    model.assemble_M()
    model.assemble_K()
    model.assemble_D()
    model.assemble_F()
    model.apply_dirichlet_matrices()
    model.apply_dirichlet_F()
    d = {"nx_ny_nz": np.array([4, 2, 3]), "jitter": np.array(0.2), "seed": np.array(0)}
    for nm in NAMES:
        d[nm] = np.array([getattr(el, f"compute_mat_{nm}_e")() for el in mesh.elements])
    d.update(csc_triplet("M", model.mat_M))
    d.update(csc_triplet("D", model.mat_D))
    d.update(csc_triplet("K", model.mat_K))
    d.update(csc_triplet("Kff", model.mat_K_f_f))
    d.update(maps_of(model))
    d.update(F=model.vec_F, F_f=model.vec_F_f)
    # D variants for the pattern rules of model.py:112-117
    for tag_, (aM, aK) in {"D_noray": (0., 0.), "D_aM": (0.3, 0.)}.items():
        model.alpha_M, model.alpha_K = aM, aK
        model.assemble_D()
        d.update(csc_triplet(tag_, model.mat_D))
    # time-indexed load
    arr_s = np.linspace(0, 1, 3 * n_t).reshape(3, n_t) * 1e5
    model2 = ref.model.Model(mesh, dict_surface_forces={7: arr_s}, dict_heat_flux={6: np.linspace(1, 2, n_t)},
                             dict_heat_source={2: 7.5}, dict_nodal_forces={8: np.array([1., 2., 3.])})
    model2.assemble_F(idx_t=3)
    d["F_t3"] = model2.vec_F
    # steady solve with the reference's raw spsolve
    solver = ref.solvers.LinearSteadyState(model)
    model.alpha_M, model.alpha_K = 0.3, 0.05
    solver.solve()
    d.update(X_raw_spsolve=solver.X, T=solver.T, U=solver.U)
    np.savez_compressed(os.path.join(OUT, "kuhn_small.npz"), **d)
    print("kuhn_small done: N = %d, E = %d, nnz M/D/K = %d / %d / %d" %
          (mesh.n_nodes, len(mesh.elements), model.mat_M.nnz, model.mat_D.nnz, model.mat_K.nnz))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--transient-steps", type=int, default=999)
    ap.add_argument("--only", default="")
    ap.add_argument("--rom-modes", default="30,30")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    ref = rs.load_reference()
    only = set(args.only.split(",")) if args.only else None
    if only is None or "mesh" in only:
        write_mesh_fixture()
    if only is None or "kuhn" in only:
        run_kuhn(ref)
    K = None
    if only is None or "steady" in only or "operators" in only:
        K = run_steady(ref)
    if only is None or "operators" in only:
        run_operators(ref, K)
    if only is None or "transient" in only:
        run_transient(ref, args.transient_steps, [1, 2, 10, 100, 500, 999])
    if only is None or "rom" in only:
        n_u, n_th = [int(v) for v in args.rom_modes.split(",")]
        run_rom(ref, n_u, n_th, [1, 2, 10, 100, 500, 999])


if __name__ == "__main__":
    main()
