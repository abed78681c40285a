"""Host-side logic that needs no GPU: Gmsh I/O, mesh container, materials, DOF maps, plane layouts, and that
the C-ABI library loads and exports every symbol include/tefem.h declares (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden
from oracle import tefem_oracle as orc


def test_gmsh_roundtrip_and_mesh_groups(sandwich, tmp_path):
    from thermoelasticity_fem_b200 import gmsh
    from thermoelasticity_fem_b200.mesh import Mesh
    pts, blocks, groups = sandwich
    path = str(tmp_path / "sandwich.msh")
    gmsh.write_msh41(path, pts, blocks)
    p2, b2 = gmsh.read_msh41(path)
    assert np.array_equal(p2, pts) and len(b2) == len(blocks)
    for (t1, g1, c1), (t2, g2, c2) in zip(blocks, b2):
        assert t1 == t2 and g1 == g2 and np.array_equal(np.asarray(c1).reshape(c2.shape), c2)
    mesh = Mesh()
    mesh.load_mesh(path)
    assert mesh.n_nodes == 3119 and mesh.n_dofs == 12476
    assert list(mesh.dict_tet_groups) == [3, 2, 1]                    # tag first-appearance order (mesh.py:51)
    assert [len(v) for v in mesh.dict_tet_groups.values()] == [4513, 4520, 4679]
    assert list(mesh.dict_tri_groups) == [4, 6, 5, 7] and mesh.dict_nodes_groups[8].tolist() == [16]
    from thermoelasticity_fem_b200 import LinearThermoElastic
    m = LinearThermoElastic(rho=2300, Y=64e9, nu=0.1, k=1., c=750., alpha=4e-6, T0=20.)
    mesh.set_materials({1: m, 2: m, 3: m})
    mesh.make_elements()
    assert mesh.n_elements == 13712 and len(mesh.elements) == 13712
    conn, mat_id, table = mesh.element_tables()
    oconn, omid, orows = orc.element_table(groups["tet"], {k: orc.Material(2300, 64e9, 0.1, 1., 750., 4e-6, 20.) for k in (1, 2, 3)})
    assert np.array_equal(conn, oconn) and np.array_equal(mat_id, omid) and np.array_equal(table, orows)
    el = mesh.elements[100]
    assert el.number == 100 and el.dofs_nums_t == [4 * n + 3 for n in conn[100]]
    assert np.array_equal(mesh.reference_temperature(), np.full(3119, 20.0))
    with pytest.raises(KeyError):
        mesh.set_materials({1: m})
        mesh.make_elements()


def test_material_matches_reference_values():
    from thermoelasticity_fem_b200 import LinearThermoElastic, glass_SG773
    m = LinearThermoElastic(rho=2300, Y=64e9, nu=0.1, k=1., c=750., alpha=4e-6, T0=20.)
    o = orc.Material(2300, 64e9, 0.1, 1., 750., 4e-6, 20.)
    assert m.beta == o.beta == 320000.0 and np.array_equal(m.mat_C, o.mat_C) and np.array_equal(m.table_row(), o.row())
    assert glass_SG773.beta == m.beta


def test_dof_maps_match_reference_without_gpu(sandwich):
    """Model.create_free_dofs_lists is pure host logic (device upload is lazy)."""
    from thermoelasticity_fem_b200 import Mesh, Model
    pts, blocks, groups = sandwich
    mesh = Mesh()
    mesh.from_blocks(pts, blocks)
    for name, dU in (("sandwich_steady.npz", {4: [('x', 0.), ('y', 0.), ('z', 0.)], 5: [('x', 0.), ('y', 0.), ('z', 0.)]}),
                     ("sandwich_transient2.npz", {4: [('x', 0.), ('y', 0.), ('z', 0.)], 5: [('y', 0.), ('z', 0.)]})):
        g = golden(name)
        model = Model(mesh, dict_dirichlet_U=dU, dict_dirichlet_theta={4: 0., 5: 0.})
        model.create_free_dofs_lists()
        for k in ("free_dofs", "free_dofs_U", "free_dofs_theta"):
            assert np.array_equal(getattr(model, k), g[k]), (name, k)          # bit-exact DOF maps
        assert np.array_equal(model.dirichlet_dofs_U, g["dirichlet_dofs_U_sorted"])
        assert np.array_equal(model.dirichlet_dofs_theta, g["dirichlet_dofs_theta_sorted"])
    with pytest.raises(ValueError, match="Unrecognized DOF"):
        Model(mesh, dict_dirichlet_U={4: [('q', 0.)]}).create_free_dofs_lists()


def test_plane_maps_cover_exactly_the_touched_entries():
    from thermoelasticity_fem_b200 import layout
    K = layout.plane_map_K()
    assert sorted(K[K >= 0].tolist()) == list(range(13)) and np.all(K[3, :3] == -1)       # no tu block in K
    M = layout.plane_map_M()
    assert sorted(M[M >= 0].tolist()) == [0, 1, 2] and M[3, 3] == -1                       # M has no theta rows
    for (aM, aK), n in (((0., 0.), 4), ((0.2, 0.), 7), ((0., 0.3), 13), ((0.2, 0.2), 13)):
        D = layout.plane_map_D(aM, aK)
        assert layout.n_planes(D) == n and sorted(D[D >= 0].tolist()) == list(range(n)) and np.all(D[:3, 3] == -1)


def test_library_loads_and_exports_every_declared_symbol():
    from thermoelasticity_fem_b200 import _lib
    header = open(os.path.join(ROOT, "include", "tefem.h")).read()
    declared = sorted(set(re.findall(r"\b(tefem_[a-z_0-9]+)\s*\(", header)))
    assert len(declared) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/tefem.h but not exported"
    assert sorted(_lib.declared_symbols()) == declared                 # the ctypes table binds exactly the header
    # host-only entry points work without a GPU
    W4 = ctypes.c_double()
    S = (ctypes.c_double * 4)()
    NN = (ctypes.c_double * 16)()
    assert _lib.load().tefem_quadrature_tables(ctypes.byref(W4), S, NN) == 0
    oW4, oS, oNN = orc.quadrature_tables()
    assert W4.value == oW4 and list(S) == list(oS) and list(NN) == list(oNN.reshape(-1))   # bit-exact constants
    assert "sm_100a" in _lib.version()
    assert _lib.load().tefem_launch_count() == 0


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from thermoelasticity_fem_b200 import Mesh, Model, LinearThermoElastic
    from thermoelasticity_fem_b200.synthetic import kuhn_box
    pts, tets, tris, nodes = kuhn_box(2, 2, 2)
    mesh = Mesh().from_arrays(pts, tets, tris, nodes)
    m = LinearThermoElastic(rho=2300, Y=64e9, nu=0.1, k=1., c=750., alpha=4e-6, T0=20.)
    mesh.set_materials({1: m, 2: m, 3: m})
    mesh.make_elements()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Model(mesh).assemble_K()


def test_reduced_order_api_order_of_calls_and_no_fallback():
    """compute_ROB_* mirror the reference's preconditions (model.py:342,354: DOF lists and assembled matrices first) and,
    like every numerical entry point, have no CPU path."""
    import torch
    from thermoelasticity_fem_b200 import Mesh, Model, LinearThermoElastic, LinearTransient
    from thermoelasticity_fem_b200.synthetic import kuhn_box
    pts, tets, tris, nodes = kuhn_box(2, 2, 2)
    mesh = Mesh().from_arrays(pts, tets, tris, nodes)
    m = LinearThermoElastic(rho=2300, Y=64e9, nu=0.1, k=1., c=750., alpha=4e-6, T0=20.)
    mesh.set_materials({1: m, 2: m, 3: m})
    mesh.make_elements()
    model = Model(mesh, dict_dirichlet_U={4: [('x', 0.), ('y', 0.), ('z', 0.)]}, dict_dirichlet_theta={4: 0.})
    assert model.mat_phi_u is None and model.mat_phi_t is None and model.mat_Mrom is None and model.vec_From is None
    with pytest.raises(RuntimeError, match="create_free_dofs_lists"):
        model.compute_ROB_u(3)
    with pytest.raises(RuntimeError, match="assemble_F"):
        model.compute_ROM_forces()
    if not torch.cuda.is_available():
        model.create_free_dofs_lists()
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            model.compute_ROB_theta(3)
        N = mesh.n_nodes
        solver = LinearTransient(model, np.zeros(3 * N), np.zeros(3 * N), np.zeros(N), np.zeros(N), 1.0, 5)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            solver.solve_ROM(2, 2)
        with pytest.raises(NotImplementedError):
            solver.solve_ROM_nonparametric(2, 2, {}, 1, [], [])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "thermoelasticity_fem_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src and "host_emulation" not in src, fn


def test_preconditioner_hierarchy_host_logic():
    """Aggregation / level-2 set-up of multilevel.py (pure torch + numpy, no device code)."""
    import torch
    from thermoelasticity_fem_b200 import multilevel as ml
    from thermoelasticity_fem_b200.synthetic import kuhn_box
    pts, tets, tris, nodes = kuhn_box(12, 8, 6, 3.0, 2.0, 1.5)
    N = pts.shape[0]
    A = ml.build_aggregation(torch.from_numpy(pts), N, 3.0)
    n1 = A["n1"]
    assert abs(A["h"] - (3.0 * 2.0 * 1.5 / N) ** (1 / 3)) < 1e-12
    agg = A["agg"].numpy()
    assert agg.min() == 0 and agg.max() == n1 - 1 and np.unique(agg).size == n1
    # every aggregate is a box of edge <= H1 (cubes of the anchored grid)
    for a in range(0, n1, max(1, n1 // 12)):
        box = pts[agg == a]
        assert np.all(box.max(axis=0) - box.min(axis=0) <= A["H1"] + 1e-9)
    cen = A["centres"].numpy()
    assert np.abs(cen - np.stack([np.bincount(agg, weights=pts[:, k], minlength=n1) / np.bincount(agg, minlength=n1)
                                  for k in range(3)], axis=1)).max() < 1e-13
    assert np.abs(A["off"].numpy() - (pts - cen[agg])).max() < 1e-13
    # rigid-body transfer between the levels: a rigid motion of the level-2 aggregates is reproduced at the centres
    agg2s, n2s, _ = ml.build_level2(A["uniq"], A["dims1"], 0.0, dense_limit=8)
    P2 = ml.transfer_level2(cen, agg2s, n2s)
    T, W = np.array([0.3, -0.2, 0.5]), np.array([0.1, 0.4, -0.7])
    C2 = np.stack([np.bincount(agg2s, weights=cen[:, k], minlength=n2s) / np.bincount(agg2s, minlength=n2s) for k in range(3)], axis=1)
    x2 = np.zeros((n2s, 8)); x2[:, :3] = T + np.cross(W, C2); x2[:, 3] = 2.0; x2[:, 4:7] = W      # global rigid motion
    x1 = (P2 @ x2.ravel()).reshape(n1, 8)
    assert np.abs(x1[:, :3] - (T + np.cross(W, cen))).max() < 1e-12 and np.abs(x1[:, 4:7] - W).max() == 0
    assert np.all(x1[:, 3] == 2.0) and np.all(x1[:, 7] == 0.0)
    # level 2: identity below the dense limit, else cubes of factor2 level-1 cells with a few hundred nodes at most
    agg2, n2, f2 = ml.build_level2(A["uniq"], A["dims1"], 0.0, dense_limit=10 ** 6)
    assert n2 == n1 and np.array_equal(agg2, np.arange(n1))
    agg2, n2, f2 = ml.build_level2(A["uniq"], A["dims1"], 0.0, dense_limit=8)
    assert f2 >= 3.0 and 1 <= n2 < n1 and agg2.min() == 0 and agg2.max() == n2 - 1 and np.unique(agg2).size == n2
    d1, d2 = int(A["dims1"][1]), int(A["dims1"][2])
    ijk = np.stack([A["uniq"] // (d1 * d2), (A["uniq"] // d2) % d1, A["uniq"] % d2], axis=1)
    for a in range(n2):
        cells = ijk[agg2 == a]
        assert np.all(cells.max(axis=0) - cells.min(axis=0) < f2)


def test_plots_module_exports_the_names_the_reference_examples_import():
    """reference examples import these at module top level (example_steadystate.py:13, example_UQ_1.py): the names must exist."""
    from thermoelasticity_fem_b200.plots import (animate_U_T, plot_U_T, plot_dofU_vs_t, plot_dofU_vs_t_random,  # noqa: F401
                                                  plot_nodeT_vs_t, plot_nodeT_vs_t_random)


def test_persistent_cycle_buffer_sizes_are_consistent():
    """Host-only size functions of the row-partitioned coarse cycle (include/tefem.h): the peer region holds the plain partial
    residual plus TWO areas of tagged vectors (16 bytes per value = 2 doubles) with a slot of partial sums per rank; the local
    buffer holds bc, two areas without the partial slots, the plain r2 and 32 words of counters / stamps."""
    from thermoelasticity_fem_b200 import _lib
    lib = _lib.load()
    al = lambda n: (n + 15) // 16 * 16
    for n1, n2, n_pre, n_post in ((27, 27, 3, 3), (3375, 125, 3, 3), (13824, 512, 3, 3), (13824, 512, 1, 0), (500, 64, 8, 8)):
        a1, a2 = al(8 * n1), al(8 * n2)
        iterates = 2 + (n_pre - 1) + n_post                       # sized for the classic cycle (one iterate more than the overlapped one)
        shared = (iterates + 1) * a1 + 2 * a2                     # iterates, t, then x2 and r2
        peer_area = shared + 8 * a1                               # + one slot of tagged partial sums per rank (<= 8)
        assert lib.tefem_precond_peer_doubles(n1, n2, n_pre, n_post) == a1 + 2 * 2 * peer_area
        assert lib.tefem_precond_cycle_doubles(n1, n2, n_pre, n_post) == 2 * (a1 + 2 * shared) + a2 + 32
        assert lib.tefem_precond_work_doubles(n1, n2) == 4 * a1 + 2 * a2 + 16


def test_vtu_export_round_trip(tmp_path):
    """Binary export for plotting (thermoelasticity_fem_b200/vtk.py): what the reference hands to
    pyvista in plot_U_T / animate_U_T (plots.py:6-75), written as raw-appended VTK XML and read back."""
    from thermoelasticity_fem_b200 import Mesh
    from thermoelasticity_fem_b200 import vtk, plots
    from thermoelasticity_fem_b200.synthetic import kuhn_box
    pts, tets, tris, nodes = kuhn_box(3, 2, 2, jitter=0.1, seed=4)
    mesh = Mesh().from_arrays(pts, tets, tris, nodes)
    n = mesh.n_nodes
    rng = np.random.default_rng(0)
    T, U = rng.standard_normal(n), rng.standard_normal(3 * n)
    f = plots.export_U_T(mesh, T, U, str(tmp_path / "state.vtu"))
    head = open(f, "rb").read(200).decode("ascii", "replace")
    assert head.startswith('<?xml version="1.0"?>') and 'type="UnstructuredGrid"' in head and 'header_type="UInt64"' in head
    back = vtk.read_vtu(f)
    assert np.array_equal(back["points"], mesh.table_nodes) and np.array_equal(back["tets"], mesh.all_tets())
    assert np.array_equal(back["point_data"]["T"], T) and np.array_equal(back["point_data"]["U"], U.reshape(n, 3))
    tags = np.concatenate([np.full(len(t), k, dtype=np.int32) for k, t in mesh.dict_tet_groups.items()])
    assert np.array_equal(back["cell_data"]["tag"], tags)
    # time series: every second frame + the ParaView index
    n_t = 5
    vec_t = np.linspace(0., 2., n_t)
    Th, Uh = rng.standard_normal((n, n_t)), rng.standard_normal((3 * n, n_t))
    files = plots.export_animation(mesh, Th, Uh, vec_t, str(tmp_path / "run"), step=2)
    assert [os.path.basename(p) for p in files] == ["run_0000.vtu", "run_0001.vtu", "run_0002.vtu", "run.pvd"]
    assert np.array_equal(vtk.read_vtu(files[1])["point_data"]["T"], Th[:, 2])
    assert np.array_equal(vtk.read_vtu(files[2])["point_data"]["U"], Uh[:, 4].reshape(n, 3))
    pvd = open(files[-1]).read()
    assert pvd.count("<DataSet") == 3 and 'timestep="1"' in pvd and 'file="run_0002.vtu"' in pvd
    with pytest.raises(ValueError):
        vtk.write_vtu(str(tmp_path / "bad.vtu"), pts, mesh.all_tets(), {"T": T[:-1]})
    with pytest.raises(ValueError):
        plots.export_animation(mesh, Th, Uh[:, :-1], vec_t, str(tmp_path / "bad"))


def test_confidence_bounds_match_the_reference_loops():
    """plots.confidence_bounds against the reference's per-row loops (plots.py:117-125), restated."""
    from thermoelasticity_fem_b200 import plots
    rng = np.random.default_rng(1)
    samples = rng.standard_normal((7, 40))
    for level in (0.95, 0.9, 0.5, 1.0):
        lo, hi = plots.confidence_bounds(samples, level)
        n_rej = int(np.floor((1 - level) * samples.shape[1] / 2))
        for i in range(samples.shape[0]):
            o = np.sort(samples[i, :])
            assert lo[i] == o[n_rej] and hi[i] == o[-(n_rej + 1)]
    # the plotting functions themselves need optional packages: a clear ImportError when those are missing, never at import
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="matplotlib"):
            plots.plot_dofU_vs_t(0, np.arange(3.), np.zeros((2, 3)))
