"""Reduced-order path (reference model.py:341-440, solvers.py:219-354).

CPU: the block shift-invert Lanczos iteration of rom.py against dense eigen-decompositions (scipy-backed operators).
GPU: bases / reduced solve on data/sandwich.msh against the committed outputs of the reference's own solve_ROM
(tests/golden/sandwich_rom3.npz, the configuration of examples/example_transient_3.py).  The reference's bases are
unique only up to sign (rotation inside a multiple eigenvalue), so the comparison uses basis-independent quantities:
eigenvalues, the B-orthogonal projection of a fixed vector onto each subspace, and the physical histories U, T."""
import numpy as np
import pytest
import scipy.linalg
import scipy.sparse as sp
import scipy.sparse.linalg as spl
import torch

from conftest import golden
from thermoelasticity_fem_b200.rom import block_lanczos_smallest

MAT2 = (2300, 64e9, 0.1, 1., 1., 4e-6, 20.)


def laplacian3(m, neumann=False):
    T = sp.diags([-np.ones(m - 1), 2 * np.ones(m), -np.ones(m - 1)], [-1, 0, 1]).tolil()
    if neumann:
        T[0, 0] = T[m - 1, m - 1] = 1.0
    T = T.tocsc()
    eye = sp.identity(m, format="csc")
    return (sp.kron(sp.kron(T, eye), eye) + sp.kron(sp.kron(eye, T), eye) + sp.kron(sp.kron(eye, eye), T)).tocsc()


def scipy_ops(A, B, shift):
    lu = spl.splu((A + shift * B).tocsc())
    return (lambda X: torch.from_numpy((B @ X.numpy().T).T.copy()),
            lambda R, X0=None: torch.from_numpy(lu.solve(R.numpy().T.copy()).T.copy()))


@pytest.mark.parametrize("neumann", [False, True])
def test_block_lanczos_matches_dense_eigh(neumann):
    m, k = 9, 14
    A = laplacian3(m, neumann)                     # triple eigenvalues (cube symmetry); singular when neumann
    n = A.shape[0]
    rng = np.random.default_rng(3)
    B = sp.diags(1.0 + rng.random(n)).tocsc()
    shift = 0.05 if neumann else 0.0
    apply_B, solve_A = scipy_ops(A, B, shift)
    x0 = torch.from_numpy(rng.random((6, n)) - 0.5)
    lam, X, info = block_lanczos_smallest(apply_B, solve_A, x0, k, tol=1e-10, shift=shift,
                                          apply_A=(lambda X: torch.from_numpy((A @ X.numpy().T).T.copy())) if not neumann else None)
    assert info["converged"]
    ref, vec = scipy.linalg.eigh(A.toarray(), B.toarray())
    assert np.max(np.abs(lam - ref[:k])) <= 1e-9 * ref[k]
    Xn = X.numpy()
    assert np.allclose(Xn @ (B @ Xn.T), np.eye(k), atol=1e-9)                     # B-orthonormal rows
    # same invariant subspace wherever the cut does not split a cluster
    cut = k
    while cut > 0 and ref[cut] - ref[cut - 1] < 1e-8 * ref[k]:
        cut -= 1
    P_ref = vec[:, :cut] @ (vec[:, :cut].T @ B.toarray())
    lead = Xn[:cut].T
    assert np.max(np.abs(P_ref @ lead - lead)) < 1e-7
    res = np.linalg.norm(A @ Xn.T - (B @ Xn.T) * lam, axis=0) / np.maximum(np.linalg.norm(A @ Xn.T, axis=0), 1e-300)
    assert res[1 if neumann else 0:].max() < 1e-7


def test_block_lanczos_needs_enough_vectors():
    A = laplacian3(4)
    B = sp.identity(A.shape[0], format="csc")
    apply_B, solve_A = scipy_ops(A, B, 0.0)
    x0 = torch.from_numpy(np.random.default_rng(0).random((2, A.shape[0])) - 0.5)
    with pytest.raises(RuntimeError):
        block_lanczos_smallest(apply_B, solve_A, x0, 10, max_vectors=4)


def test_reduced_newmark_loop_against_the_reference_recurrence():
    """solvers.reduced_newmark (host loop of solve_ROM) against the recurrence written exactly as reference solvers.py:268-291."""
    from thermoelasticity_fem_b200.solvers import reduced_newmark
    rng = np.random.default_rng(11)
    nq, n_t, dt, gamma, beta = 7, 40, 0.37, 0.5, 0.25
    B = rng.standard_normal((nq, nq))
    Mr, Kr = B @ B.T + nq * np.eye(nq), rng.standard_normal((nq, nq)) + 5 * nq * np.eye(nq)
    Dr = 0.2 * Mr + 0.1 * Kr + rng.standard_normal((nq, nq))              # non-symmetric like the coupled Drom
    F = rng.standard_normal((nq, n_t))
    q0, v0 = rng.standard_normal(nq), rng.standard_normal(nq)
    q, qd, qdd = reduced_newmark(Mr, Dr, Kr, lambda i: F[:, i], q0, v0, dt, gamma, beta, n_t)
    Kdyn = Mr + gamma * dt * Dr + beta * dt ** 2 * Kr
    prev_q, prev_qdot, prev_qdotdot = q0.copy(), v0.copy(), np.zeros(nq)
    for i in range(1, n_t):
        rhs = (F[:, i] - Dr @ (prev_qdot + (1 - gamma) * dt * prev_qdotdot)
               - Kr @ (prev_q + dt * prev_qdot + 0.5 * (dt ** 2) * (1 - 2 * beta) * prev_qdotdot))
        new_qdotdot = np.linalg.solve(Kdyn, rhs)
        new_qdot = prev_qdot + (1 - gamma) * dt * prev_qdotdot + gamma * dt * new_qdotdot
        new_q = prev_q + dt * prev_qdot + 0.5 * (dt ** 2) * ((1 - 2 * beta) * prev_qdotdot + 2 * beta * new_qdotdot)
        for got, want in ((q[:, i], new_q), (qd[:, i], new_qdot), (qdd[:, i], new_qdotdot)):
            assert np.abs(got - want).max() <= 1e-11 * max(np.abs(want).max(), 1.0)
        prev_q, prev_qdot, prev_qdotdot = new_q, new_qdot, new_qdotdot
    assert np.array_equal(q[:, 0], q0) and np.array_equal(qd[:, 0], v0) and not qdd[:, 0].any()


def test_oracle_reduced_order_restatement_matches_reference_fixture(sandwich):
    """Pins oracle.reduced_order_transient (numpy / scipy restatement of model.py:341-440 + solvers.py:219-354) to the
    outputs of the reference's own solve_ROM: eigenvalues, subspaces, histories."""
    from oracle import tefem_oracle as orc
    g = golden("sandwich_rom3.npz")
    pts, blocks, groups = sandwich
    m = orc.Material(*MAT2)
    t_end, n_t = 1800e0, 1000
    vec_t = np.linspace(0., t_end, n_t)
    arr = np.zeros((3, n_t))
    arr[0, :] = np.sin(2 * np.pi * vec_t / 900.) * -1e9
    n_u, n_th = [int(v) for v in g["n_modes"]]
    steps = [int(v) for v in g["steps"]]
    r = orc.reduced_order_transient(pts, groups, {1: m, 2: m, 3: m}, {4: [('x', 0.), ('y', 0.), ('z', 0.)], 5: [('y', 0.), ('z', 0.)]},
                                    {4: 0., 5: 0.}, dict(surface={5: arr}), 0.2, 0.2, t_end, n_t, n_u, n_th, keep=steps)
    assert np.array_equal(r["maps"]["free_dofs_U"], g["free_dofs_U"]) and np.array_equal(r["maps"]["free_dofs_theta"], g["free_dofs_theta"])
    assert np.max(np.abs(r["lam_u"] / g["lam_u"] - 1)) < 1e-9 and np.max(np.abs(r["lam_t"] / g["lam_t"] - 1)) < 1e-9
    for nm, phi, op, free in (("u", r["phi_u"], "M", g["free_dofs_U"]), ("t", r["phi_t"], "D", g["free_dofs_theta"])):
        B = r["ops"][op].tocsr()[free, :].tocsc()[:, free]
        w = fingerprint_weights(free.size)
        proj = phi @ np.linalg.solve(phi.T @ (B @ phi), phi.T @ (B @ w))
        assert np.linalg.norm(proj - g["proj_" + nm]) <= 1e-8 * np.linalg.norm(g["proj_" + nm]), nm
    N = pts.shape[0]
    for X, U_ref, T_ref, shift in ((r["X"], g["U"], g["T"], 20.0), (r["Xdot"], g["Udot"], g["Tdot"], 0.0)):
        U = np.zeros((3 * N, len(steps)))
        for c in range(3):
            U[c::3] = X[c::4]
        T = X[3::4] + shift                                                         # reference temperature: T0 = 20 everywhere
        assert np.abs(U - U_ref).max() <= 1e-8 * np.abs(U_ref).max()
        assert np.abs(T - T_ref).max() <= 1e-8 * np.abs(T_ref - shift).max()


# ---- GPU ---------------------------------------------------------------------------------------------------------------

def fingerprint_weights(n):                         # = oracle/make_golden.py
    k = np.arange(n, dtype=np.float64)
    return 1.0 + np.mod(k * 0.6180339887498949, 1.0)


def rom_model(sandwich):
    from thermoelasticity_fem_b200 import LinearThermoElastic, Mesh, Model
    pts, blocks, groups = sandwich
    mesh = Mesh().from_arrays(pts, groups["tet"], groups["tri"], groups["nodes"])
    mesh.set_materials({k: LinearThermoElastic(*MAT2) for k in (1, 2, 3)})
    mesh.make_elements()
    t_end, n_t = 1800e0, 1000
    vec_t = np.linspace(0., t_end, n_t)
    arr = np.zeros((3, n_t))
    arr[0, :] = np.sin(2 * np.pi * vec_t / 900.) * -1e9
    model = Model(mesh, dict_dirichlet_U={4: [('x', 0.), ('y', 0.), ('z', 0.)], 5: [('y', 0.), ('z', 0.)]},
                  dict_dirichlet_theta={4: 0., 5: 0.}, dict_surface_forces={5: arr}, alpha_M=2e-1, alpha_K=2e-1)
    return mesh, model, t_end, n_t


@pytest.mark.gpu
def test_reduced_bases_span_the_reference_subspaces(sandwich):
    g = golden("sandwich_rom3.npz")
    mesh, model, t_end, n_t = rom_model(sandwich)
    model.create_free_dofs_lists()
    model.assemble_MDK()
    n_u, n_th = [int(v) for v in g["n_modes"]]
    model.compute_ROB_u(n_u)
    model.compute_ROB_theta(n_th)
    assert model.rob_info_u["converged"] and model.rob_info_t["converged"]
    print("ROB info", model.rob_info_u, model.rob_info_t)
    assert model.rob_info_u["residual"] < 1e-8 and model.rob_info_t["residual"] < 1e-8     # |K x - lam B x| / |K x|
    assert np.max(np.abs(model.eig_u / g["lam_u"] - 1)) < 1e-9
    assert np.max(np.abs(model.eig_t / g["lam_t"] - 1)) < 1e-9
    assert np.array_equal(model.free_dofs_U, g["free_dofs_U"]) and np.array_equal(model.free_dofs_theta, g["free_dofs_theta"])
    for nm, phi, op, free in (("u", model.mat_phi_u, "M", model.free_dofs_U), ("t", model.mat_phi_t, "D", model.free_dofs_theta)):
        assert phi.shape == (free.size, n_u if nm == "u" else n_th)
        assert np.allclose(np.linalg.norm(phi, axis=0), 1.0, atol=1e-12)           # unit columns like eigs
        B = model.touched(op).tocsr()[free, :].tocsc()[:, free]
        w = fingerprint_weights(free.size)
        proj = phi @ np.linalg.solve(phi.T @ (B @ phi), phi.T @ (B @ w))
        perr = np.linalg.norm(proj - g["proj_" + nm]) / np.linalg.norm(g["proj_" + nm])
        print("subspace fingerprint error", nm, perr, "eigenvalue error",
              np.max(np.abs((model.eig_u if nm == "u" else model.eig_t) / g["lam_" + nm] - 1)))
        assert perr <= 1e-8, nm          # measured on the B200: 1.5e-12 (u), 1.9e-11 (theta)


@pytest.mark.gpu
def test_solve_ROM_matches_reference_histories(sandwich):
    from thermoelasticity_fem_b200 import LinearTransient
    g = golden("sandwich_rom3.npz")
    mesh, model, t_end, n_t = rom_model(sandwich)
    N = mesh.n_nodes
    solver = LinearTransient(model, np.zeros(3 * N), np.zeros(3 * N), np.zeros(N), np.zeros(N), t_end, n_t, 0.5, 0.25,
                             progress=False)
    n_u, n_th = [int(v) for v in g["n_modes"]]
    solver.solve_ROM(n_modes_u=n_u, n_modes_theta=n_th)
    assert solver.U.shape == (3 * N, n_t) and solver.T.shape == (N, n_t) and solver.q.shape == (n_u + n_th, n_t)
    assert abs(np.abs(solver.U).max() / float(g["maxU"]) - 1) < 1e-6
    steps = g["steps"]
    for nm in ("U", "T", "Udot", "Tdot"):
        got, ref = getattr(solver, nm)[:, steps], g[nm]
        ref_shift = 20.0 if nm == "T" else 0.0                                      # T carries the reference temperature
        err = np.abs(got - ref).max(axis=0) / np.maximum(np.abs(ref - ref_shift).max(axis=0), 1e-300)
        print("solve_ROM history error", nm, err)
        assert err.max() < 1e-8, (nm, err)          # measured: 1e-12 (U), 2e-11 (T)
    # the reduced matrices reproduce the pencils' eigenvalues: Krom_uu / Mrom_uu and Krom_tt / Drom_tt
    m = solver.model
    lu = np.sort(scipy.linalg.eigvals(m.mat_Krom[:n_u, :n_u], m.mat_Mrom[:n_u, :n_u]).real)
    lt = np.sort(scipy.linalg.eigvals(m.mat_Krom[n_u:, n_u:], m.mat_Drom[n_u:, n_u:]).real)
    assert np.max(np.abs(lu / g["lam_u"] - 1)) < 1e-6 and np.max(np.abs(lt / g["lam_t"] - 1)) < 1e-6


@pytest.mark.gpu
def test_rom_forces_equal_projected_load_vector(sandwich):
    mesh, model, t_end, n_t = rom_model(sandwich)
    model.create_free_dofs_lists()
    model.assemble_MDK()
    model.compute_ROB_u(5, tol=1e-7)
    model.compute_ROB_theta(4, tol=1e-7)
    for idx in (1, 250):
        model.assemble_F(idx_t=idx)
        model.compute_ROM_forces()                                                   # model.py:395-440 from vec_F
        fast = model.rom_forces(idx)                                                 # cached projected weights
        assert np.max(np.abs(fast - model.vec_From)) <= 1e-12 * np.abs(model.vec_From).max()
        F = model.vec_F
        want = np.r_[model.mat_phi_u.T @ F[model.free_dofs_U], model.mat_phi_t.T @ F[model.free_dofs_theta]]
        assert np.max(np.abs(model.vec_From - want)) <= 1e-12 * np.abs(want).max()


def test_solve_ROM_host_flow_with_stubbed_device_steps(monkeypatch):
    """The host side of LinearTransient.solve_ROM (initial projection, reduced loop, expansion, prescribed values, reference
    temperature, history option) with the device steps replaced by CPU stand-ins: bases = random vectors supported on the free
    DOFs of their field, reduced matrices = their exact projections of random operators."""
    from thermoelasticity_fem_b200 import Mesh, Model, LinearThermoElastic, LinearTransient
    from thermoelasticity_fem_b200.synthetic import kuhn_box
    pts, tets, tris, nodes = kuhn_box(3, 2, 3)
    mesh = Mesh().from_arrays(pts, tets, tris, nodes)
    mats = {1: LinearThermoElastic(2300, 64e9, 0.1, 1., 750., 4e-6, 20.), 2: LinearThermoElastic(7800, 210e9, 0.3, 45., 460., 1.2e-5, 25.),
            3: LinearThermoElastic(2700, 70e9, 0.33, 200., 900., 2.3e-5, 30.)}
    mesh.set_materials(mats)
    mesh.make_elements()
    model = Model(mesh, dict_dirichlet_U={4: [('x', 1e-3), ('y', 0.), ('z', 0.)]}, dict_dirichlet_theta={4: 2.5})
    N, n = mesh.n_nodes, mesh.n_dofs
    n_t, nu, nt = 9, 3, 2
    rng = np.random.default_rng(2)

    def fake_rob(field, k):
        model.create_free_dofs_lists()
        comp = np.arange(n) % 4
        free = ((comp < 3) if field == "u" else (comp == 3)) & ~model._mask_host
        return torch.from_numpy(rng.standard_normal((k, n)) * free)

    def compute_u(k, tol=None):
        model.n_q_u, model._phi_u_dev = k, fake_rob("u", k)

    def compute_t(k, tol=None):
        model.n_q_t, model._phi_t_dev = k, fake_rob("theta", k)

    def matrices():
        nq = model.n_q_u + model.n_q_t
        B = rng.standard_normal((nq, nq))
        model.mat_Mrom, model.mat_Krom = B @ B.T + nq * np.eye(nq), 3.0 * np.eye(nq) + 0.1 * rng.standard_normal((nq, nq))
        model.mat_Drom = 0.2 * model.mat_Mrom

    forces = rng.standard_normal((nu + nt, n_t))
    monkeypatch.setattr(model, "assemble_MDK", lambda: None)
    monkeypatch.setattr(model, "compute_ROB_u", compute_u)
    monkeypatch.setattr(model, "compute_ROB_theta", compute_t)
    monkeypatch.setattr(model, "compute_ROM_matrices", matrices)
    monkeypatch.setattr(model, "rom_forces", lambda i: forces[:, i])
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    U0 = rng.standard_normal(3 * N)
    th0 = rng.standard_normal(N)
    solver = LinearTransient(model, U0, np.zeros(3 * N), th0, np.zeros(N), 2.0, n_t, progress=False)
    solver.solve_ROM(nu, nt)
    Phi = np.vstack([model._phi_u_dev.numpy(), model._phi_t_dev.numpy()])
    X0 = np.zeros(n)
    X0[0::4] = X0[1::4] = X0[2::4] = U0[::3]                                   # reference quirk, solvers.py:237-239
    X0[3::4] = th0
    assert np.allclose(solver.q[:, 0], Phi @ X0, rtol=1e-13, atol=1e-13)
    assert solver.q.shape == (nu + nt, n_t) and solver.X.shape == (n, n_t) and solver.U.shape == (3 * N, n_t) and solver.T.shape == (N, n_t)
    X = Phi.T @ solver.q
    mask = model._mask_host
    X[mask, :] = model._g_fill_host[mask][:, None]
    assert np.allclose(solver.X, X, rtol=1e-12, atol=1e-12)
    assert np.allclose(solver.Xdot, Phi.T @ solver.qdot, rtol=1e-12, atol=1e-12)
    for c in range(3):
        assert np.array_equal(solver.U[c::3], solver.X[c::4])
    assert np.allclose(solver.T, solver.X[3::4] + mesh.reference_temperature()[:, None])
    dir_nodes = np.unique(np.asarray(mesh.dict_tri_groups[4]).reshape(-1))
    assert np.all(solver.U[3 * dir_nodes, :] == 1e-3) and np.allclose(solver.T[dir_nodes, :], 2.5 + mesh.reference_temperature()[dir_nodes][:, None])
    q_full = solver.q.copy()
    solver2 = LinearTransient(model, U0, np.zeros(3 * N), th0, np.zeros(N), 2.0, n_t, progress=False)
    monkeypatch.setattr(model, "compute_ROB_u", lambda k, tol=None: None)      # keep the same bases
    monkeypatch.setattr(model, "compute_ROB_theta", lambda k, tol=None: None)
    monkeypatch.setattr(model, "compute_ROM_matrices", lambda: None)
    solver2.solve_ROM(nu, nt, history='last')
    assert np.allclose(solver2.q, q_full) and solver2.X.shape == (n, 1) and np.allclose(solver2.X[:, 0], solver.X[:, -1])


class _ScipyPlanes:
    """Stand-in for an operator in plane storage (CPU tests of the host logic only): a scipy matrix."""

    def __init__(self, A):
        self.A = A.tocsr()


class _ScipyDeviceMesh:
    """Stand-in for DeviceMesh with the one call the ROM host logic makes: y = beta_y y + alpha Op_masked x, where the plane
    map selects which (row component, column component) entries of the 4x4 node blocks the operator has."""
    device = torch.device("cpu")

    def planes_spmv(self, vals, plane_map, x, y, alpha=1.0, beta_y=0.0):
        n = vals.A.shape[0]
        comp = np.arange(n) % 4
        out = np.zeros(n)
        xv = x.numpy()
        for r in range(4):
            for c in range(4):
                if plane_map[r, c] >= 0:
                    out[comp == r] += (vals.A[comp == r, :][:, comp == c] @ xv[comp == c])
        base = 0.0 if beta_y == 0.0 else beta_y * y.numpy()     # like the kernel: y is not read when beta_y == 0 (may be uninitialised)
        y.copy_(torch.from_numpy(base + alpha * out))
        return y


def test_reduced_matrices_and_forces_host_logic_against_the_oracle(monkeypatch):
    """Model.compute_ROM_matrices / rom_forces / compute_ROM_forces (projection block structure, per-field lifting with NON-ZERO
    prescribed values, cached load projections) with the plane SpMV replaced by scipy, against direct projections of the
    oracle's operators (reference model.py:365-440)."""
    from oracle import tefem_oracle as orc
    from thermoelasticity_fem_b200 import Mesh, Model, LinearThermoElastic, _lib, layout
    from thermoelasticity_fem_b200.synthetic import kuhn_box
    monkeypatch.setattr(_lib, "require_cuda", lambda: torch.device("cpu"))
    monkeypatch.setattr(_lib, "load", lambda: None)
    pts, tets, tris, nodes = kuhn_box(3, 2, 3, jitter=0.15, seed=4)
    props = {1: (2300, 64e9, 0.1, 1., 750., 4e-6, 20.), 2: (7800, 210e9, 0.3, 45., 460., 1.2e-5, 25.), 3: (2700, 70e9, 0.33, 200., 900., 2.3e-5, 30.)}
    mesh = Mesh().from_arrays(pts, tets, tris, nodes)
    mesh.set_materials({k: LinearThermoElastic(*v) for k, v in props.items()})
    mesh.make_elements()
    n_t = 6
    amp = np.linspace(0.0, 1.0, n_t)
    surf = np.outer(np.array([0., 1e5, -1e6]), amp)
    flux = -25.0 * amp
    dU = {4: [('x', 2e-4), ('y', 1e-4), ('z', 0.)], 5: [('x', -1e-4)]}
    dT = {4: 3.0, 6: -1.5}
    aM, aK = 0.3, 0.05
    model = Model(mesh, dict_dirichlet_U=dU, dict_dirichlet_theta=dT, dict_surface_forces={7: surf}, dict_heat_flux={7: flux},
                  alpha_M=aM, alpha_K=aK)
    model.create_free_dofs_lists()
    groups = dict(tet=tets, tri=tris, nodes=nodes)
    omats = {k: orc.Material(*v) for k, v in props.items()}
    conn, mat_id, rows = orc.element_table(tets, omats)
    ops = orc.assemble(pts, conn, rows, mat_id, aM, aK, which="MKD")
    monkeypatch.setattr(type(mesh), "device_mesh", lambda self: _ScipyDeviceMesh())
    model._planes = {k: _ScipyPlanes(ops[k]) for k in "MDK"}
    model._d_alpha = (aM, aK)
    n = mesh.n_dofs
    fu, ft = model.free_dofs_U, model.free_dofs_theta
    rng = np.random.default_rng(9)
    nu, nt = 4, 3
    phi_u, phi_t = rng.standard_normal((fu.size, nu)), rng.standard_normal((ft.size, nt))
    Pu, Pt = np.zeros((nu, n)), np.zeros((nt, n))
    Pu[:, fu], Pt[:, ft] = phi_u.T, phi_t.T
    model.n_q_u, model.n_q_t = nu, nt
    model._phi_u_dev, model._phi_t_dev = torch.from_numpy(Pu), torch.from_numpy(Pt)
    assert np.array_equal(model.mat_phi_u, phi_u) and np.array_equal(model.mat_phi_t, phi_t)

    def sub(A, r, c):
        return A.tocsr()[r, :].tocsc()[:, c]

    model.compute_ROM_matrices()
    M, D, K = ops["M"], ops["D"], ops["K"]
    want_M = np.zeros((nu + nt, nu + nt)); want_D = np.zeros_like(want_M); want_K = np.zeros_like(want_M)
    want_M[:nu, :nu] = phi_u.T @ (sub(M, fu, fu) @ phi_u)
    want_D[:nu, :nu] = phi_u.T @ (sub(D, fu, fu) @ phi_u)
    want_D[nu:, :nu] = phi_t.T @ (sub(D, ft, fu) @ phi_u)
    want_D[nu:, nu:] = phi_t.T @ (sub(D, ft, ft) @ phi_t)
    want_K[:nu, :nu] = phi_u.T @ (sub(K, fu, fu) @ phi_u)
    want_K[:nu, nu:] = phi_u.T @ (sub(K, fu, ft) @ phi_t)
    want_K[nu:, nu:] = phi_t.T @ (sub(K, ft, ft) @ phi_t)
    for got, want in ((model.mat_Mrom, want_M), (model.mat_Drom, want_D), (model.mat_Krom, want_K)):
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    # reduced loads exactly as reference model.py:395-440: per-tag lifting through K[free_U, dir_U] and K[free_theta, dir_theta]
    Kcsr = K.tocsr()
    for idx in (1, 3, 5):
        F = orc.assemble_F(pts, groups, dict(surface={7: surf}, flux={7: flux}), idx_t=idx)
        Fu, Ft = F[fu].copy(), F[ft].copy()
        for tag, lst in dU.items():
            nodes_t = np.unique(np.asarray(tris[tag]).reshape(-1))
            vec = np.zeros(n)
            dofs = []
            for node in nodes_t:
                for letter, val in lst:
                    dofs.append(4 * node + "xyz".index(letter))
                    vec[4 * node + "xyz".index(letter)] = val
            Fu -= Kcsr[fu, :][:, dofs] @ vec[dofs]
        for tag, theta in dT.items():
            nodes_t = np.unique(np.asarray(tris[tag]).reshape(-1))
            dofs = [4 * node + 3 for node in nodes_t]
            Ft -= Kcsr[ft, :][:, dofs] @ np.full(len(dofs), theta)
        want = np.r_[phi_u.T @ Fu, phi_t.T @ Ft]
        got = model.rom_forces(idx)
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max(), idx
        model._F = torch.from_numpy(F)                        # what assemble_F(idx_t) leaves on the device
        model.compute_ROM_forces()
        assert np.abs(model.vec_From - want).max() <= 1e-12 * np.abs(want).max(), idx


@pytest.mark.parametrize("theta_dirichlet", [True, False])
def test_field_eigenproblem_host_logic_with_scipy_solves(monkeypatch, theta_dirichlet):
    """rom.FieldEigenProblem (field masks, field-restricted plane maps, shift for a field without Dirichlet DOFs, Lanczos +
    polish, unit-norm vectors) with the plane SpMV and the Krylov solves replaced by scipy, against dense eigh of the oracle's
    K_uu / M_uu and K_tt / D_tt on the free DOFs (reference model.py:341-363)."""
    from oracle import tefem_oracle as orc
    from thermoelasticity_fem_b200 import Mesh, Model, LinearThermoElastic, _lib, rom, solvers
    from thermoelasticity_fem_b200.synthetic import kuhn_box
    monkeypatch.setattr(_lib, "require_cuda", lambda: torch.device("cpu"))
    pts, tets, tris, nodes = kuhn_box(4, 3, 3, jitter=0.1, seed=7)
    props = {1: (2300, 64e9, 0.1, 1., 750., 4e-6, 20.), 2: (7800, 210e9, 0.3, 45., 460., 1.2e-5, 25.), 3: (2700, 70e9, 0.33, 200., 900., 2.3e-5, 30.)}
    mesh = Mesh().from_arrays(pts, tets, tris, nodes)
    mesh.set_materials({k: LinearThermoElastic(*v) for k, v in props.items()})
    mesh.make_elements()
    model = Model(mesh, dict_dirichlet_U={4: [('x', 0.), ('y', 0.), ('z', 0.)]},
                  dict_dirichlet_theta={4: 0.} if theta_dirichlet else None, alpha_M=0.1, alpha_K=0.0)
    model.create_free_dofs_lists()
    conn, mat_id, rows = orc.element_table(tets, {k: orc.Material(*v) for k, v in props.items()})
    ops = orc.assemble(pts, conn, rows, mat_id, 0.1, 0.0, which="MKD")
    monkeypatch.setattr(type(mesh), "device_mesh", lambda self: _ScipyDeviceMesh())
    model._planes = {k: _ScipyPlanes(ops[k]) for k in "MDK"}
    model._d_alpha = (0.1, 0.0)
    n = mesh.n_dofs
    seen = {}

    class ScipySystemSolve:                                   # stand-in for solvers._SystemSolve (form_system + BiCGSTAB)
        def __init__(self, model_, cM, cD, cK, rtol, max_iter, check_every, precond="auto", dir_mask=None):
            elim = dir_mask.numpy().astype(bool)
            A = (cM * ops["M"] + cD * ops["D"] + cK * ops["K"]).tocsr()
            keep = sp.diags((~elim).astype(float))
            self.elim = elim
            self.lu = spl.splu((keep @ A @ keep + sp.diags(elim.astype(float))).tocsc())
            seen["coef"] = (cM, cD, cK)

        def solve(self, rhs, x):
            b = rhs.numpy().copy()
            b[self.elim] = 0.0
            x.copy_(torch.from_numpy(self.lu.solve(b)))

    monkeypatch.setattr(solvers, "_SystemSolve", ScipySystemSolve)
    for field, k, free, A_name, B_name in (("u", 6, model.free_dofs_U, "K", "M"), ("theta", 5, model.free_dofs_theta, "K", "D")):
        prob = rom.FieldEigenProblem(model, field)
        singular = field == "theta" and not theta_dirichlet
        assert (prob.shift > 0.0) == singular                                       # pure-Neumann theta field: shifted solves
        assert seen["coef"] == ((prob.shift, 0.0, 1.0) if field == "u" else (0.0, prob.shift, 1.0))
        lam, Phi, info = prob.smallest(k, tol=1e-10)
        A = ops[A_name].tocsr()[free, :].tocsc()[:, free].toarray()
        B = ops[B_name].tocsr()[free, :].tocsc()[:, free].toarray()
        ref = scipy.linalg.eigh(0.5 * (A + A.T), 0.5 * (B + B.T), eigvals_only=True)[:k]
        scale = np.abs(ref).max()
        assert np.abs(lam - ref).max() <= 1e-8 * scale, (field, lam, ref)
        if singular:
            assert abs(ref[0]) <= 1e-9 * scale                                       # the constant mode
        assert info["converged"]
        P = Phi.numpy()
        outside = np.ones(n, dtype=bool)
        outside[free] = False
        assert not P[:, outside].any() and np.allclose(np.linalg.norm(P, axis=1), 1.0, atol=1e-12)
        X = P[:, free].T
        res = np.linalg.norm(A @ X - (B @ X) * lam, axis=0) / np.maximum(np.linalg.norm(B @ X, axis=0) * scale, 1e-300)
        assert res.max() < 1e-8
