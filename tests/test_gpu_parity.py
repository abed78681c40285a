"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the committed reference outputs."""
import numpy as np
import pytest

from conftest import golden, csc_from
import parity

pytestmark = pytest.mark.gpu

NAMES = ["Muu", "Dtu", "Dtt", "Kuu", "Kut", "Ktt"]
MAT1 = (2300, 64e9, 0.1, 1., 750., 4e-6, 20.)
MAT2 = (2300, 64e9, 0.1, 1., 1., 4e-6, 20.)
KUHN_MATS = {1: (2300, 64e9, 0.1, 1., 750., 4e-6, 20.), 2: (7800, 210e9, 0.3, 45., 460., 1.2e-5, 25.),
             3: (2700, 70e9, 0.33, 200., 900., 2.3e-5, 30.)}


def make_mesh(pts, groups, mats):
    from thermoelasticity_fem_b200 import LinearThermoElastic, Mesh
    mesh = Mesh().from_arrays(pts, groups["tet"], groups["tri"], groups["nodes"])
    mesh.set_materials({k: LinearThermoElastic(*v) for k, v in mats.items()})
    mesh.make_elements()
    return mesh


@pytest.fixture(scope="module")
def sandwich_mesh1(sandwich):
    pts, blocks, groups = sandwich
    return make_mesh(pts, groups, {1: MAT1, 2: MAT1, 3: MAT1})


@pytest.fixture(scope="module")
def sandwich_mesh2(sandwich):
    pts, blocks, groups = sandwich
    return make_mesh(pts, groups, {1: MAT2, 2: MAT2, 3: MAT2})


def kuhn_case():
    from thermoelasticity_fem_b200.synthetic import kuhn_box
    pts, tets, tris, nodes = kuhn_box(4, 2, 3, jitter=0.2, seed=0)
    return pts, dict(tet=tets, tri=tris, nodes=nodes)


def test_library_loaded():
    from thermoelasticity_fem_b200 import _lib
    assert "sm_100a" in _lib.version()


def test_quadrature_tables_match_oracle():
    import ctypes
    from thermoelasticity_fem_b200 import _lib
    from oracle import tefem_oracle as orc
    W4 = ctypes.c_double()
    S = (ctypes.c_double * 4)()
    NN = (ctypes.c_double * 16)()
    _lib.check(_lib.load().tefem_quadrature_tables(ctypes.byref(W4), S, NN))
    oW4, oS, oNN = orc.quadrature_tables()
    assert W4.value == oW4 and list(S) == list(oS) and list(NN) == list(oNN.reshape(-1))     # bit-exact


def test_element_matrices_sandwich(sandwich_mesh1):
    g = golden("sandwich_elements.npz")
    sub = g["subset"]
    em = sandwich_mesh1.element_matrices(NAMES, sub)
    for nm in NAMES:
        ref = g[nm]
        scale = np.abs(ref).reshape(len(sub), -1).max(axis=1)[:, None, None]
        assert np.max(np.abs(em[nm] - ref) / scale) < parity.VALUE_TOL, nm
    # Element view API (reference elements.py:181-286)
    el = sandwich_mesh1.elements[int(sub[3])]
    assert np.max(np.abs(el.compute_mat_Kuu_e() - g["Kuu"][3])) < parity.VALUE_TOL * np.abs(g["Kuu"][3]).max()
    assert el.dofs_nums_u[:3] == [4 * el.nodes_nums[0], 4 * el.nodes_nums[0] + 1, 4 * el.nodes_nums[0] + 2]
    assert el.dofs_nums_t == [4 * n + 3 for n in el.nodes_nums]


def test_element_matrices_kuhn_all():
    g = golden("kuhn_small.npz")
    pts, groups = kuhn_case()
    mesh = make_mesh(pts, groups, KUHN_MATS)
    em = mesh.element_matrices(NAMES)
    for nm in NAMES:
        assert np.max(np.abs(em[nm] - g[nm])) <= parity.VALUE_TOL * np.abs(g[nm]).max(), nm


def test_assembly_sandwich_vs_reference(sandwich_mesh2, sandwich):
    from thermoelasticity_fem_b200 import Model
    from oracle import tefem_oracle as orc
    pts, blocks, groups = sandwich
    g = golden("sandwich_operators.npz")
    model = Model(sandwich_mesh2, alpha_M=0.2, alpha_K=0.2)
    model.assemble_M()
    model.assemble_K()
    model.assemble_D()
    m = orc.Material(*MAT2)
    conn, mat_id, rows = orc.element_table(groups["tet"], {1: m, 2: m, 3: m})
    ops = orc.assemble(pts, conn, rows, mat_id, 0.2, 0.2)
    ops["M"] = orc.drop_component_zeros_M(ops["M"])
    for nm in "MDK":
        A = model.touched(nm)
        ref = csc_from(g, nm)
        parity.check_pattern(A, ref, ops[nm])                       # touched pattern bit-exact; reference subset
        assert parity.scaled_error(A, ref) < parity.VALUE_TOL, nm   # values vs the reference's own assembly
        assert parity.scaled_error(A, ops[nm]) < parity.VALUE_TOL, nm
    assert model.touched("K").nnz == 519805 and model.mat_M.nnz == 119955
    # the exported attribute mimics csc_array(dense): int32 sorted indices, zeros dropped
    K = model.mat_K
    assert K.format == "csc" and K.shape == (12476, 12476) and K.has_sorted_indices
    assert abs(K.nnz - 519725) < 200
    # the actual set difference between the two STORED patterns (H1: entries whose floating-point sum is exactly 0.0 are
    # dropped by csc_array(dense) on either side, and which ones cancel exactly depends on the summation order): every
    # entry stored by only one side must be round-off noise on the other
    kmine, kref = parity.keys(K), parity.keys(csc_from(g, "K"))
    only_mine, only_ref = np.setdiff1d(kmine, kref), np.setdiff1d(kref, kmine)
    print(f"mat_K stored patterns: this repo {K.nnz}, reference {kref.size}, only here {only_mine.size}, only in the reference {only_ref.size}")
    Kr = csc_from(g, "K").tocsr()
    Km = K.tocsr()
    scale = np.abs(Kr).max()
    if only_mine.size:
        assert np.abs(np.asarray(Km[only_mine // 12476, only_mine % 12476])).max() <= 1e-12 * scale
    if only_ref.size:
        assert np.abs(np.asarray(Kr[only_ref // 12476, only_ref % 12476])).max() <= 1e-12 * scale
    # fused launch == three launches, bitwise; repeated assembly bitwise reproducible
    model2 = Model(sandwich_mesh2, alpha_M=0.2, alpha_K=0.2)
    model2.assemble_MDK()
    for nm in "MDK":
        assert np.array_equal(model2._planes[nm].cpu().numpy(), model._planes[nm].cpu().numpy()), nm
    model2.assemble_MDK()
    for nm in "MDK":
        assert np.array_equal(model2._planes[nm].cpu().numpy(), model._planes[nm].cpu().numpy()), nm


def test_assembly_kuhn_variants():
    from thermoelasticity_fem_b200 import Model
    g = golden("kuhn_small.npz")
    pts, groups = kuhn_case()
    mesh = make_mesh(pts, groups, KUHN_MATS)
    for tag_, (aM, aK) in {"D": (0.3, 0.05), "D_noray": (0., 0.), "D_aM": (0.3, 0.)}.items():
        model = Model(mesh, alpha_M=aM, alpha_K=aK)
        model.assemble_D()
        ref = csc_from(g, tag_)
        A = model.touched("D")
        parity.check_pattern(A, ref)
        assert parity.scaled_error(A, ref) < parity.VALUE_TOL, tag_
    model = Model(mesh, alpha_M=0.3, alpha_K=0.05)
    model.assemble_K()
    model.assemble_M()
    assert parity.scaled_error(model.touched("K"), csc_from(g, "K")) < parity.VALUE_TOL
    assert parity.scaled_error(model.touched("M"), csc_from(g, "M")) < parity.VALUE_TOL
    assert np.array_equal(parity.keys(model.mat_M), parity.keys(csc_from(g, "M")))


def test_dirichlet_maps_loads_lifting_kuhn():
    from thermoelasticity_fem_b200 import Model
    g = golden("kuhn_small.npz")
    pts, groups = kuhn_case()
    mesh = make_mesh(pts, groups, KUHN_MATS)
    model = Model(mesh,
                  dict_dirichlet_U={4: [('x', 0.), ('y', 1e-4), ('z', 0.)], 5: [('x', 2e-4)]},
                  dict_dirichlet_theta={4: 3.0, 6: -1.5},
                  dict_nodal_forces={8: np.array([1e3, -2e3, 5e2])},
                  dict_surface_forces={7: np.array([0., 1e5, -1e6])},
                  dict_volume_forces={1: np.array([0., 0., -2300 * 9.81]), 3: np.array([1., 2., 3.])},
                  dict_heat_flux={7: -25.}, dict_heat_source={2: 1e3}, alpha_M=0.3, alpha_K=0.05)
    model.create_free_dofs_lists()
    for k in ("free_dofs", "free_dofs_U", "free_dofs_theta"):
        assert np.array_equal(getattr(model, k), g[k]), k
    assert np.array_equal(np.sort(model.dirichlet_dofs_U), g["dirichlet_dofs_U_sorted"])
    assert np.array_equal(np.sort(model.dirichlet_dofs_theta), g["dirichlet_dofs_theta_sorted"])
    model.assemble_K()
    model.assemble_F()
    model.apply_dirichlet_matrices()
    model.apply_dirichlet_F()
    assert np.max(np.abs(model.vec_F - g["F"])) <= 1e-13 * np.abs(g["F"]).max()
    assert np.max(np.abs(model.vec_F_f - g["F_f"])) <= 1e-12 * np.abs(g["F_f"]).max()
    Kff = model.mat_K_f_f
    ref = csc_from(g, "Kff")
    assert Kff.shape == ref.shape
    assert np.isin(parity.keys(ref), parity.keys(Kff)).all() or (abs(Kff - ref)).max() < 1e-12 * abs(ref).max()
    assert (abs(Kff - ref)).max() <= 1e-12 * abs(ref).max()
    # time-indexed loads
    n_t = 5
    model2 = Model(mesh, dict_surface_forces={7: np.linspace(0, 1, 3 * n_t).reshape(3, n_t) * 1e5},
                   dict_heat_flux={6: np.linspace(1, 2, n_t)}, dict_heat_source={2: 7.5},
                   dict_nodal_forces={8: np.array([1., 2., 3.])})
    model2.assemble_F(idx_t=3)
    assert np.max(np.abs(model2.vec_F - g["F_t3"])) <= 1e-13 * np.abs(g["F_t3"]).max()
    with pytest.raises(ValueError, match="1 or 2 dimensions"):
        Model(mesh, dict_surface_forces={7: np.zeros((2, 2, 2))}).assemble_F(idx_t=0)
    with pytest.raises(ValueError, match="Unrecognized DOF"):
        Model(mesh, dict_dirichlet_U={4: [('w', 0.)]}).create_free_dofs_lists()


def test_negative_jacobian_raises():
    from thermoelasticity_fem_b200 import Model
    pts, groups = kuhn_case()
    tets = {k: v.copy() for k, v in groups["tet"].items()}
    tets[2][5] = tets[2][5][[0, 2, 1, 3]]                 # flips the orientation of element 48 + 5
    mesh = make_mesh(pts, dict(tet=tets, tri=groups["tri"], nodes=groups["nodes"]), KUHN_MATS)
    with pytest.raises(ValueError, match="Element 53 has negative jacobian."):
        Model(mesh).assemble_K()
    tets[1][7][3] = tets[1][7][2]                          # degenerate element 7 comes first in element order
    mesh = make_mesh(pts, dict(tet=tets, tri=groups["tri"], nodes=groups["nodes"]), KUHN_MATS)
    with pytest.raises(ValueError, match="Element 7 has zero jacobian."):
        Model(mesh).assemble_K()


def test_spmv_and_system_against_scipy(sandwich_mesh2):
    import torch
    from thermoelasticity_fem_b200 import Model
    model = Model(sandwich_mesh2, dict_dirichlet_U={4: [('x', 0.), ('y', 0.), ('z', 0.)], 5: [('y', 0.), ('z', 0.)]},
                  dict_dirichlet_theta={4: 0., 5: 0.}, alpha_M=0.2, alpha_K=0.2)
    model.create_free_dofs_lists()
    model.assemble_MDK()
    dm = sandwich_mesh2.device_mesh()
    n = sandwich_mesh2.n_dofs
    rng = np.random.default_rng(0)
    xh = rng.standard_normal(n)
    x = torch.from_numpy(xh).cuda()
    y = torch.zeros_like(x)
    for nm in "MDK":
        dm.planes_spmv(model._planes[nm], model.plane_map(nm), x, y)
        ref = model.touched(nm) @ xh
        assert np.max(np.abs(y.cpu().numpy() - ref)) <= 1e-13 * np.abs(ref).max(), nm
    dt, gam, bet = 1.8, 0.5, 0.25
    sys = dm.form_system(model._planes, 1.0, gam * dt, bet * dt * dt, 13, model._dir_mask)
    free = model.free_dofs
    Kdyn = (model.touched("M") + gam * dt * model.touched("D") + bet * dt * dt * model.touched("K")).tocsr()
    dm.spmv(sys, x, y)
    mask = model._mask_host
    ref = np.where(mask, xh, 0.0)
    ref[free] = Kdyn[free, :][:, free] @ xh[free]
    assert np.max(np.abs(y.cpu().numpy() - ref)) <= 1e-13 * np.abs(ref).max()
    # block-Jacobi scaling: dinv * A_ii = I
    sys2 = sys.clone()
    dinv = dm.block_jacobi_scale(sys2)
    diag = sys2[dm.schedule.diag_idx.long()].reshape(-1, 4, 4).cpu().numpy()
    assert np.max(np.abs(diag - np.eye(4))) < 1e-9      # u rows ~1e10, theta rows ~1: cond(A_ii) ~ 1e6


def test_steady_state_sandwich(sandwich_mesh1, sandwich):
    from thermoelasticity_fem_b200 import Model, LinearSteadyState
    from oracle import tefem_oracle as orc
    pts, blocks, groups = sandwich
    g = golden("sandwich_steady.npz")
    dU = {4: [('x', 0.), ('y', 0.), ('z', 0.)], 5: [('x', 0.), ('y', 0.), ('z', 0.)]}
    dT = {4: 0., 5: 0.}
    model = Model(sandwich_mesh1, dict_dirichlet_U=dU, dict_dirichlet_theta=dT,
                  dict_surface_forces={7: np.array([0., 0., -1e9])}, dict_volume_forces=None,
                  dict_heat_flux={6: -25.}, dict_heat_source=None)
    solver = LinearSteadyState(model)
    solver.solve()
    # default on the reference's own mesh: the staged solve (K_tt then K_uu, CG); it must reach rtol, not the stall floor
    assert solver.solve_info["method"] == "staged-cg" and not solver.solve_info["floor"], solver.solve_info
    for k in ("free_dofs", "free_dofs_U", "free_dofs_theta"):
        assert np.array_equal(getattr(model, k), g[k]), k
    assert np.max(np.abs(model.vec_F - g["F"])) <= 1e-13 * np.abs(g["F"]).max()
    m = orc.Material(*MAT1)
    X, info = orc.steady_state(pts, groups, {1: m, 2: m, 3: m}, dU, dT, dict(surface={7: np.array([0., 0., -1e9])}, flux={6: -25.}))
    eu, et = parity.field_errors(solver.X, X)
    assert eu < parity.SOLUTION_TOL and et < parity.SOLUTION_TOL, (eu, et)
    ru, rt = orc.field_scaled_residual(info["Kff"], solver.X[model.free_dofs], info["Ff"], model.free_dofs)
    assert ru < 1e-8 and rt < 1e-8
    # informational distance to the reference's raw spsolve (theta limited by SuperLU on this scaling, H2)
    eu_raw, et_raw = parity.field_errors(solver.X, g["X_raw_spsolve"])
    assert eu_raw < 1e-7 and et_raw < 1e-4
    assert abs(np.abs(solver.U).max() - 0.2111117771259024) < 1e-8
    assert abs(solver.T.max() - 46.05154818198166) < 1e-3 and solver.T.min() == 20.0
    assert solver.U.shape == (3 * 3119,) and solver.T.shape == (3119,)


def test_steady_state_kuhn_nonzero_dirichlet():
    import __graft_entry__
    __graft_entry__.smoke()


def test_transient_sandwich_first_steps(sandwich_mesh2, sandwich):
    from thermoelasticity_fem_b200 import Model, LinearTransient
    from oracle import tefem_oracle as orc
    pts, blocks, groups = sandwich
    g = golden("sandwich_transient2.npz") if __import__("os").path.exists(
        __import__("os").path.join(__import__("conftest").GOLDEN, "sandwich_transient2.npz")) else None
    t_end, n_t = 1800e0, 1000
    vec_t = np.linspace(0., t_end, n_t)
    arr = np.zeros((3, n_t))
    arr[0, :] = np.sin(2 * np.pi * vec_t / 900.) * -1e9
    dU = {4: [('x', 0.), ('y', 0.), ('z', 0.)], 5: [('y', 0.), ('z', 0.)]}
    dT = {4: 0., 5: 0.}
    model = Model(sandwich_mesh2, dict_dirichlet_U=dU, dict_dirichlet_theta=dT, dict_surface_forces={5: arr},
                  alpha_M=2e-1, alpha_K=2e-1)
    n_steps = 10
    N = sandwich_mesh2.n_nodes
    solver = LinearTransient(model, np.zeros(3 * N), np.zeros(3 * N), np.zeros(N), np.zeros(N), t_end, n_t, 0.5, 0.25,
                             progress=False)
    solver.prepare()
    states = {}
    for i in range(1, n_steps + 1):
        info = solver.step(i)
        assert not info["floor"] and info["relres"] <= 1e-10, info          # reached rtol, not the stall floor
        states[i] = tuple(t.cpu().numpy().copy() for t in (solver._x, solver._v, solver._a))
    m = orc.Material(*MAT2)
    ref = orc.transient(pts, groups, {1: m, 2: m, 3: m}, dU, dT, dict(surface={5: arr}), 0.2, 0.2, t_end, n_t,
                        n_steps=n_steps, keep=[1, 2, 10])
    free = model.free_dofs
    assert free.shape[0] == 11874
    for i in (1, 2, 10):
        for k, name in enumerate(("x", "v", "a")):
            full = np.zeros(4 * N)
            full[free] = ref["hist"][i][k]
            eu, et = parity.field_errors(states[i][k], full)
            # x (the solution the north star names) to 1e-8; its time derivatives accumulate the per-step
            # solve error of the acceleration (cond * rtol) and are checked to 1e-7
            tol = parity.SOLUTION_TOL if name == "x" else 10 * parity.SOLUTION_TOL
            assert eu < tol and et < 10 * tol, (i, name, eu, et)
    if g is not None:
        # reference's own states (raw spsolve each step): u to ~1e-8, theta to ~1e-5 (H2)
        for col, step in enumerate(g["steps"]):
            if step <= n_steps:
                eu, et = parity.field_errors(states[int(step)][0], g["X"][:, col])
                assert eu < 1e-6 and et < 1e-3, (step, eu, et)


def test_two_gpu_transient_matches_single_gpu():
    """Row-partitioned solve (NCCL halo exchange + packed all-reduces) == single-GPU solve.  Needs >= 2 GPUs."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(root, "tools", "dist_check.py"), "16", "2"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "DIST_CHECK OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_steady_state_synthetic_sandwich_vs_refined_oracle():
    """BASELINE config 3 in miniature: Kuhn sandwich (three different layer materials), config-1 boundary conditions,
    against the oracle's LU + iterative refinement."""
    from thermoelasticity_fem_b200 import Model, LinearSteadyState
    from thermoelasticity_fem_b200.synthetic import kuhn_box
    from oracle import tefem_oracle as orc
    pts, tets, tris, nodes = kuhn_box(24, 6, 6)
    groups = dict(tet=tets, tri=tris, nodes=nodes)
    mesh = make_mesh(pts, groups, KUHN_MATS)
    dU = {4: [('x', 0.), ('y', 0.), ('z', 0.)], 5: [('x', 0.), ('y', 0.), ('z', 0.)]}
    dT = {4: 0., 5: 0.}
    model = Model(mesh, dict_dirichlet_U=dU, dict_dirichlet_theta=dT, dict_surface_forces={7: np.array([0., 0., -1e9])},
                  dict_heat_flux={6: -25.})
    solver = LinearSteadyState(model)
    solver.solve()
    X, info = orc.steady_state(pts, groups, {k: orc.Material(*v) for k, v in KUHN_MATS.items()}, dU, dT,
                               dict(surface={7: np.array([0., 0., -1e9])}, flux={6: -25.}))
    K = model.touched("K")
    parity.check_pattern(K, None, info["K"])
    assert parity.scaled_error(K, info["K"]) < parity.VALUE_TOL
    eu, et = parity.field_errors(solver.X, X)
    assert eu < parity.SOLUTION_TOL and et < parity.SOLUTION_TOL, (eu, et)
    # reference temperature = mean of T0 over the materials touching a node (solvers.py:51-60)
    T0 = orc.reference_temperature(pts.shape[0], tets, {k: orc.Material(*v) for k, v in KUHN_MATS.items()})
    assert np.allclose(solver.T, X[3::4] + T0, rtol=0, atol=1e-6 * np.abs(X[3::4]).max())


def _kuhn_steady_model(nx, ny, nz):
    from thermoelasticity_fem_b200 import Model
    from thermoelasticity_fem_b200.synthetic import kuhn_box
    pts, tets, tris, nodes = kuhn_box(nx, ny, nz)
    groups = dict(tet=tets, tri=tris, nodes=nodes)
    mesh = make_mesh(pts, groups, KUHN_MATS)
    dU = {4: [('x', 0.), ('y', 0.), ('z', 0.)], 5: [('x', 0.), ('y', 0.), ('z', 0.)]}
    dT = {4: 0., 5: 0.}
    loads = dict(surface={7: np.array([0., 0., -1e9])}, flux={6: -25.})
    model = Model(mesh, dict_dirichlet_U=dU, dict_dirichlet_theta=dT, dict_surface_forces=loads["surface"],
                  dict_heat_flux=loads["flux"])
    return pts, groups, mesh, model, dU, dT, loads


def test_aggregation_preconditioner_apply_matches_host_restatement():
    """z = omega r + P1 V(P1^T r) of csrc/tefem_precond.cu against the same operator written with scipy from the arrays
    the preconditioner holds, and the level-1 operator against the Galerkin product P1^T A^ P1 of the fine system for
    the rigid-body prolongation (translations + rotations about the aggregate centres, constant theta)."""
    import scipy.sparse as sp
    import torch
    from thermoelasticity_fem_b200.solvers import _SystemSolve
    from thermoelasticity_fem_b200.multilevel import TwoLevelPreconditioner
    pts, groups, mesh, model, *_ = _kuhn_steady_model(16, 9, 6)
    model.create_free_dofs_lists()
    model.assemble_K()
    model.assemble_F()
    model.apply_dirichlet_matrices()
    ss = _SystemSolve(model, 0., 0., 1., 1e-10, 1000, 16, precond="block_jacobi")
    pc, S = TwoLevelPreconditioner(ss.dm, ss.sys, factor1=2.0, factor2=0.0, dense_limit=10), ss.dm.schedule
    n, n1, n2 = mesh.n_nodes, pc.n1, pc.n2
    assert 1 < n2 < n1 < n
    A = sp.bsr_array((ss.sys.cpu().numpy().reshape(-1, 4, 4), S.colidx.cpu().numpy(), S.rowptr.cpu().numpy()),
                     shape=(4 * n, 4 * n)).tocsr()
    agg = pc.node_agg.cpu().numpy().astype(np.int64)
    off = pc.node_off.double().cpu().numpy()[:, :3]
    # offsets are relative to the aggregate centres (mean of the member nodes)
    cen = np.stack([np.bincount(agg, weights=pts[:, k], minlength=n1) / np.bincount(agg, minlength=n1) for k in range(3)], axis=1)
    assert np.abs(off - (pts - cen[agg])).max() <= 1e-6 * np.abs(pts).max()
    ar = np.arange(n)
    rows, cols, vals = [], [], []
    for c in range(4):
        rows.append(4 * ar + c); cols.append(8 * agg + c); vals.append(np.ones(n))
    dx, dy, dz = off[:, 0], off[:, 1], off[:, 2]
    for (c, k, v) in [(0, 1, dz), (0, 2, -dy), (1, 2, dx), (1, 0, -dz), (2, 0, dy), (2, 1, -dx)]:
        rows.append(4 * ar + c); cols.append(8 * agg + 4 + k); vals.append(v)
    P1 = sp.csr_array((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(4 * n, 8 * n1))

    def bsr(rowptr, colidx, data, shape):
        return sp.bsr_array((data.double().cpu().numpy().reshape(-1, 4, 4), colidx.cpu().numpy(), rowptr.cpu().numpy()),
                            shape=shape).tocsr()

    A1 = bsr(pc.rowptr1, pc.colidx1, pc.sys1, (8 * n1, 8 * n1))
    D1inv = sp.bsr_array((pc.dinv1.cpu().numpy().reshape(-1, 8, 8), np.arange(n1), np.arange(n1 + 1)),
                         shape=(8 * n1, 8 * n1)).tocsr()
    G = (D1inv @ (P1.T @ A @ P1)).tocsr()
    Gd, A1d = G.toarray(), A1.toarray()
    off_diag = ~np.kron(np.eye(n1, dtype=bool), np.ones((8, 8), dtype=bool))
    assert np.abs(Gd - A1d)[off_diag].max() <= 1e-3 * np.abs(Gd).max()          # level 1 is stored in half precision
    blk = np.abs(A1d[~off_diag].reshape(n1, 8, 8) - np.eye(8)).max()
    assert blk <= 1e-3                                                           # scaled diagonal blocks = identity (half precision)
    P2 = bsr(pc.p2_rowptr, pc.p2_colidx, pc.p2_vals, (8 * n1, 8 * n2))
    R2 = bsr(pc.r2_rowptr, pc.r2_colidx, pc.r2_vals, (8 * n2, 8 * n1))
    assert abs(R2 - P2.T).max() == 0.0
    inv2 = pc.inv2.double().cpu().numpy()
    A2 = (P2.T @ A1 @ P2).toarray()
    live = np.abs(A2).sum(axis=1) > 0
    assert np.abs((inv2 @ A2)[np.ix_(live, live)] - np.eye(int(live.sum()))).max() < 1e-3
    rng = np.random.default_rng(5)
    r = rng.standard_normal(4 * n)
    bc = D1inv @ (P1.T @ r)

    def cycle_host(kind):
        xc = pc.omega_c * bc
        if kind == "classic":
            for _ in range(pc.n_pre - 1):
                xc = xc + pc.omega_c * (bc - A1 @ xc)
            xc = xc + P2 @ (inv2 @ (P2.T @ (bc - A1 @ xc)))
        else:      # overlapped: the last pre-smoothing sweep and the level-2 correction share one residual
            for _ in range(pc.n_pre - 2):
                xc = xc + pc.omega_c * (bc - A1 @ xc)
            t = bc - A1 @ xc
            xc = xc + pc.omega_c * t + P2 @ (inv2 @ (P2.T @ t))
        for _ in range(pc.n_post):
            xc = xc + pc.omega_c * (bc - A1 @ xc)
        return pc.omega * r + P1 @ xc

    assert pc.cycle == "overlapped"
    z_host = cycle_host("overlapped")
    z = torch.empty(4 * n, dtype=torch.float64, device="cuda")
    pc.apply(torch.from_numpy(r).cuda(), z)
    assert np.abs(z.cpu().numpy() - z_host).max() <= 1e-11 * np.abs(z_host).max()
    pc3 = TwoLevelPreconditioner(ss.dm, ss.sys, factor1=2.0, factor2=0.0, dense_limit=10, cycle="classic")
    z3 = torch.empty_like(z)
    pc3.apply(torch.from_numpy(r).cuda(), z3)
    zc = cycle_host("classic")
    assert np.abs(z3.cpu().numpy() - zc).max() <= 1e-11 * np.abs(zc).max()
    # the persistent row-partitioned dataflow kernel of the coarse cycle (the multi-GPU default) on one GPU
    for kind, zref, zdev in (("overlapped", z_host, z), ("classic", zc, z3)):
        pcp = TwoLevelPreconditioner(ss.dm, ss.sys, factor1=2.0, factor2=0.0, dense_limit=10, cycle=kind, coarse="persistent")
        assert pcp.coarse == "persistent" and pc.coarse == "launches"
        zp = torch.empty_like(z)
        for _ in range(3):                 # repeated applications: barrier counters and epochs advance
            zp.zero_()
            pcp.apply(torch.from_numpy(r).cuda(), zp)
            torch.cuda.synchronize()
            assert np.abs(zp.cpu().numpy() - zref).max() <= 1e-11 * np.abs(zref).max()
        assert np.abs((zp - zdev).cpu().numpy()).max() <= 1e-11 * np.abs(zref).max()
    # the same with the level-1 operator kept in single precision
    pc2 = TwoLevelPreconditioner(ss.dm, ss.sys, factor1=2.0, factor2=0.0, dense_limit=10, level1_dtype="fp32")
    assert pc.level1_dtype == "fp16" and pc2.level1_dtype == "fp32"
    z2 = torch.empty_like(z)
    pc2.apply(torch.from_numpy(r).cuda(), z2)
    assert np.abs((z2 - z).cpu().numpy()).max() <= 5e-3 * np.abs(z_host).max()      # half-precision storage: ~5e-4 relative


def test_steady_state_multilevel_matches_block_jacobi_and_oracle():
    from thermoelasticity_fem_b200 import LinearSteadyState
    from oracle import tefem_oracle as orc
    pts, groups, mesh, model, dU, dT, loads = _kuhn_steady_model(30, 10, 8)
    sols, its = {}, {}
    for kind in ("block_jacobi", "multilevel"):
        solver = LinearSteadyState(model, precond=kind, solver="coupled")
        solver.solve()
        assert not solver.solve_info["floor"], solver.solve_info
        sols[kind], its[kind] = solver.X.copy(), solver.solve_info["iterations"]
    # staged solve: K_tt theta = F_t, then K_uu u = F_u - K_ut theta, both with CG (tefem_cg)
    solver = LinearSteadyState(model, solver="staged")
    solver.solve()
    info_s = solver.solve_info
    assert info_s["method"] == "staged-cg" and not info_s["floor"], info_s
    sols["staged"] = solver.X.copy()
    X, info = orc.steady_state(pts, groups, {k: orc.Material(*v) for k, v in KUHN_MATS.items()}, dU, dT, loads)
    for kind in sols:
        eu, et = parity.field_errors(sols[kind], X)
        assert eu < parity.SOLUTION_TOL and et < parity.SOLUTION_TOL, (kind, eu, et)
    assert its["multilevel"] < its["block_jacobi"], its
    # CG spends one SpMV per iteration, BiCGSTAB two: the staged solve must need fewer SpMVs than the coupled one
    print("iterations: coupled block-Jacobi BiCGSTAB", its["block_jacobi"], "| staged CG theta + u",
          info_s["stages"]["theta"]["iterations"], "+", info_s["stages"]["u"]["iterations"])
    assert info_s["iterations"] < 2 * its["block_jacobi"], (info_s, its)


def test_transient_sandwich_full_run_vs_reference_states(sandwich_mesh2):
    """BASELINE config 2 end to end (999 steps, public API, full histories) against the states the UNMODIFIED reference
    produced (tests/golden/sandwich_transient2.npz).  The reference re-solves with raw spsolve every step: its own theta
    carries ~5e-7 of error per solve, so the comparison is at 1e-6 (u) / 1e-4 (theta) after 999 steps."""
    from thermoelasticity_fem_b200 import Model, LinearTransient
    g = golden("sandwich_transient2.npz")
    t_end, n_t = 1800e0, 1000
    vec_t = np.linspace(0., t_end, n_t)
    arr = np.zeros((3, n_t))
    arr[0, :] = np.sin(2 * np.pi * vec_t / 900.) * -1e9
    model = Model(sandwich_mesh2, dict_dirichlet_U={4: [('x', 0.), ('y', 0.), ('z', 0.)], 5: [('y', 0.), ('z', 0.)]},
                  dict_dirichlet_theta={4: 0., 5: 0.}, dict_surface_forces={5: arr}, alpha_M=2e-1, alpha_K=2e-1)
    N = sandwich_mesh2.n_nodes
    solver = LinearTransient(model, np.zeros(3 * N), np.zeros(3 * N), np.zeros(N), np.zeros(N), t_end, n_t, 0.5, 0.25,
                             progress=False)
    solver.solve()
    assert not any(i["floor"] for i in solver.solve_info)                    # every one of the 999 solves reached rtol
    assert solver.X.shape == (4 * N, n_t) and solver.U.shape == (3 * N, n_t) and solver.T.shape == (N, n_t)
    assert solver.dt == float(g["dt"])
    for col, step in enumerate(g["steps"]):
        eu, et = parity.field_errors(solver.X[:, int(step)], g["X"][:, col])
        assert eu < 1e-6 and et < 1e-4, (int(step), eu, et)
        ev, _ = parity.field_errors(solver.Xdot[:, int(step)], g["Xdot"][:, col])
        assert ev < 1e-5, (int(step), ev)
        assert np.max(np.abs(solver.T[:, int(step)] - g["T"][:, col])) < 1e-4 * np.abs(g["T"][:, col]).max()
    assert abs(np.abs(solver.U).max() - float(g["maxU"])) < 1e-8 * float(g["maxU"])
    assert abs(np.abs(solver.X[3::4]).max() - float(g["maxTheta"])) < 1e-6 * float(g["maxTheta"])


# ---- measurement knobs of K1 -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("threads,items", [(128, 256), (64, 128)])
def test_assembly_thread_count_variants_are_bitwise_the_default(threads, items):
    """The 128- and 64-thread instantiations of K1 (tefem_assemble_set_threads) must produce the default's planes bit for bit."""
    import torch
    from thermoelasticity_fem_b200 import LinearThermoElastic
    from thermoelasticity_fem_b200.device import DeviceMesh
    from thermoelasticity_fem_b200.symbolic import build_schedule
    from thermoelasticity_fem_b200.synthetic import kuhn_box
    pts, tets, tris, nodes = kuhn_box(24, 20, 18, 1., 1., 1., n_layers=1, index_dtype=np.int32)
    conn = tets[1]
    table = np.array([LinearThermoElastic(rho=2300, Y=64e9, nu=0.1, k=1., c=1., alpha=4e-6, T0=20.).table_row()])
    outs = {}
    for tag, nt, it in (("default", 256, 512), ("variant", threads, items)):
        S = build_schedule(torch.from_numpy(conn).cuda(), pts.shape[0], items_per_cta=it, elem_budget=int(1.21 * it))
        dm = DeviceMesh(pts, conn, np.zeros(conn.shape[0], dtype=np.int32), table, schedule=S)
        dm.asm_threads = nt
        outs[tag] = {k: v.cpu().numpy() for k, v in dm.assemble("MKD", 0.2, 0.2).items()}
    for k in outs["default"]:
        assert np.array_equal(outs["default"][k], outs["variant"][k]), k


def test_clear_full_matrices_keeps_the_reduced_operators_usable():
    """reference sequence apply_dirichlet_matrices(); clear_full_matrices() (model.py:288-302): mat_K reads None, mat_K_f_f and
    every later solve keep working (the planes back both views)."""
    from thermoelasticity_fem_b200 import LinearSteadyState
    pts, groups, mesh, model, dU, dT, loads = _kuhn_steady_model(8, 4, 3)
    model.create_free_dofs_lists()
    model.assemble_K()
    model.assemble_F()
    model.apply_dirichlet_matrices()
    nnz = model.mat_K_f_f.nnz
    model.clear_full_matrices()
    assert model.mat_K is None and model.mat_M is None
    assert model.mat_K_f_f is not None and model.mat_K_f_f.nnz == nnz
    model.apply_dirichlet_F()
    assert model.lift_vector() is not None
    model.assemble_K()                       # re-assembly restores the full view
    assert model.mat_K is not None
