"""The N > 1 path on the CPU: world_size-2 (and 3) `gloo` processes run the partitioner, assemble their owned rows
with the host emulation of kernel K1, exchange halos with the send / receive lists the CUDA driver uses and
reproduce the single-process operator, SpMV and dot products."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, result_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.join(HERE, "host_emulation"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import build_emu
        from oracle import tefem_oracle as orc
        from thermoelasticity_fem_b200 import layout
        from thermoelasticity_fem_b200.partition import localize, slab_ranges
        from thermoelasticity_fem_b200.symbolic import build_schedule
        from thermoelasticity_fem_b200.synthetic import kuhn_box
        from test_symbolic_and_emulation import emu_assemble, planes_to_csc
        emu = ctypes.CDLL(build_emu.build())
        nx, ny, nz = 7, 3, 2
        pts, tets, tris, nodes = kuhn_box(nx, ny, nz, jitter=0.15, seed=2)
        mats = {1: orc.Material(2300, 64e9, 0.1, 1., 750., 4e-6, 20.), 2: orc.Material(7800, 210e9, 0.3, 45., 460., 1.2e-5, 25.),
                3: orc.Material(2700, 70e9, 0.33, 200., 900., 2.3e-5, 30.)}
        conn, mat_id, rows = orc.element_table(tets, mats)
        N = pts.shape[0]
        bounds = slab_ranges(nx + 1, (ny + 1) * (nz + 1), world)
        lp = localize(conn, np.arange(conn.shape[0]), bounds, rank)
        # local assembly of the owned rows (no communication)
        S = build_schedule(torch.from_numpy(lp.conn), lp.n_local, n_rows=lp.n_own,
                           mat_id=torch.from_numpy(mat_id[lp.elem_ids]))
        vals, bad = emu_assemble(emu, S, pts[lp.local_nodes], lp.conn, mat_id[lp.elem_ids], rows, "MKD", 0.2, 0.1)
        K_loc = planes_to_csc(vals["K"], layout.plane_map_K(), S, 4 * lp.n_local).tocsr()
        # reference: the global operator
        Kg = orc.assemble(pts, conn, rows, mat_id, 0.2, 0.1, which="K")["K"].tocsr()
        lo, hi = bounds[rank], bounds[rank + 1]
        dof_loc = (4 * lp.local_nodes[:, None] + np.arange(4)[None, :]).reshape(-1)
        K_ref = Kg[4 * lo:4 * hi, :][:, dof_loc]
        assert abs(K_loc - K_ref).max() <= 1e-12 * abs(K_ref).max()
        assert (K_loc != 0).nnz <= K_loc.nnz and K_loc.shape == K_ref.shape
        # distributed SpMV: owned x, halo filled by the exchange lists
        xg = np.random.default_rng(5).standard_normal(4 * N)
        x = np.zeros(4 * lp.n_local)
        x[:4 * lp.n_own] = xg[4 * lo:4 * hi]
        reqs, recv_bufs = [], []
        for k, q in enumerate(lp.nbr):
            idx = lp.send_idx[lp.send_ptr[k]:lp.send_ptr[k + 1]]
            send = torch.from_numpy(x.reshape(-1, 4)[idx].copy())
            recv = torch.empty((int(lp.recv_ptr[k + 1] - lp.recv_ptr[k]), 4), dtype=torch.float64)
            reqs.append(dist.isend(send, int(q)))
            reqs.append(dist.irecv(recv, int(q)))
            recv_bufs.append((k, recv))
        for r in reqs:
            r.wait()
        for k, recv in recv_bufs:
            x.reshape(-1, 4)[lp.n_own + lp.recv_ptr[k]: lp.n_own + lp.recv_ptr[k + 1]] = recv.numpy()
        assert np.array_equal(x, xg[dof_loc])                           # halo values are the owners' values
        # peer-memory halo (distributed.halo_destinations): entry k of my send list lands at node dpos[k] of rank dst[k]
        # - emulate the remote stores: every rank publishes (destination rank, position, global node id) and each rank
        # checks that what lands in its halo tail is the node it expects there, and that the tail is covered exactly once
        from thermoelasticity_fem_b200.distributed import halo_destinations
        dst, dpos = halo_destinations(lp)
        sent = [None] * world
        dist.all_gather_object(sent, (dst, dpos, lp.local_nodes[lp.send_idx]))
        landed = np.full(lp.n_local, -1, dtype=np.int64)
        for d_, p_, g_ in sent:
            mine_ = d_ == rank
            assert np.all(landed[p_[mine_]] == -1)
            landed[p_[mine_]] = g_[mine_]
        assert np.all(landed[:lp.n_own] == -1) and np.array_equal(landed[lp.n_own:], lp.local_nodes[lp.n_own:])
        y = K_loc @ x
        y_ref = (Kg @ xg)[4 * lo:4 * hi]
        assert np.max(np.abs(y - y_ref)) <= 1e-13 * np.abs(y_ref).max()
        # interior rows do not touch the halo, boundary rows do
        Kcsr = K_loc.tocsr()
        for i in lp.interior_rows[:: max(1, len(lp.interior_rows) // 20)]:
            cols = Kcsr[4 * i:4 * i + 4].indices
            assert cols.size == 0 or cols.max() < 4 * lp.n_own
        for i in lp.boundary_rows:
            assert Kcsr[4 * i:4 * i + 4].indices.max() >= 4 * lp.n_own
        assert len(lp.interior_rows) + len(lp.boundary_rows) == lp.n_own
        # packed dot products: local partials + all-reduce == global
        pkt = torch.tensor([float(y @ y), float(y @ x[:4 * lp.n_own])], dtype=torch.float64)
        dist.all_reduce(pkt)
        yg = Kg @ xg
        assert abs(pkt[0].item() - yg @ yg) <= 1e-12 * abs(yg @ yg) and abs(pkt[1].item() - yg @ xg) <= 1e-12 * abs(yg @ yg)
        # coarse level of the aggregation preconditioner: aggregates may straddle ranks; the gathered Galerkin
        # operator equals P1^T K P1 of the global operator on every rank
        import scipy.sparse as sp
        from thermoelasticity_fem_b200 import multilevel as ml
        br, ci = S.block_row.numpy().astype(np.int64), S.colidx.numpy().astype(np.int64)
        Kd = K_loc.toarray()
        blocks = np.stack([Kd[4 * r:4 * r + 4, 4 * c:4 * c + 4].reshape(16) for r, c in zip(br, ci)])
        A = ml.build_aggregation(torch.from_numpy(pts[lp.local_nodes]), lp.n_own, 2.0)
        row1, col1, blocks8 = ml.galerkin_level1(A["agg"], A["off"], S.block_row, S.colidx, torch.from_numpy(blocks), A["n1"])
        n1 = A["n1"]
        # the same aggregation computed from the global coordinates
        ptg = torch.from_numpy(pts)
        lo_g, hi_g = ptg.min(dim=0).values, ptg.max(dim=0).values
        ext = hi_g - lo_g
        H1 = 2.0 * float(ext.prod().item() / N) ** (1.0 / 3.0)
        dims1 = torch.ceil(ext / H1 + 1e-9).to(torch.int64).clamp(min=1)
        keys = ml._grid_keys(ptg, lo_g, H1, dims1).numpy()
        uniq_g, agg_g = np.unique(keys, return_inverse=True)
        assert np.array_equal(uniq_g, A["uniq"]) and np.array_equal(agg_g[lp.local_nodes], A["agg"].numpy())
        assert abs(A["H1"] - H1) < 1e-12 and 1 < n1 < N
        cen_g = np.stack([np.bincount(agg_g, weights=pts[:, k], minlength=n1) / np.bincount(agg_g, minlength=n1)
                          for k in range(3)], axis=1)
        assert np.abs(A["centres"].numpy() - cen_g).max() < 1e-13
        off_g = pts - cen_g[agg_g]
        assert np.abs(A["off"].numpy() - off_g[lp.local_nodes]).max() < 1e-13
        # rigid-body prolongation: u_i = t + w x d_i, theta_i = theta (8 coarse values per aggregate)
        ar = np.arange(N)
        rows_, cols_, vals_ = [], [], []
        for c in range(4):
            rows_.append(4 * ar + c); cols_.append(8 * agg_g + c); vals_.append(np.ones(N))
        dx, dy, dz = off_g[:, 0], off_g[:, 1], off_g[:, 2]
        for (c, k, v) in [(0, 1, dz), (0, 2, -dy), (1, 2, dx), (1, 0, -dz), (2, 0, dy), (2, 1, -dx)]:
            rows_.append(4 * ar + c); cols_.append(8 * agg_g + 4 + k); vals_.append(v)
        P1 = sp.csr_array((np.concatenate(vals_), (np.concatenate(rows_), np.concatenate(cols_))), shape=(4 * N, 8 * n1))
        A1_ref = (P1.T @ Kg @ P1).toarray()
        A1 = np.zeros((8 * n1, 8 * n1))
        for r_, c_, b_ in zip(row1.numpy(), col1.numpy(), blocks8.numpy()):
            A1[8 * r_:8 * r_ + 8, 8 * c_:8 * c_ + 8] = b_
        assert np.abs(A1 - A1_ref).max() <= 1e-12 * np.abs(A1_ref).max()
        # scaling with the 8x8 diagonal blocks and the pseudo-node block CSR hold the same operator
        rowptr1, Dinv, scaled = ml.scale_level1(row1, col1, blocks8, n1)
        rp, ci, vv = ml.pseudo_node_csr(rowptr1, col1, scaled)
        A1s = sp.bsr_array((vv.numpy().reshape(-1, 4, 4), ci.numpy(), rp.numpy()), shape=(8 * n1, 8 * n1)).toarray()
        ref = np.zeros_like(A1s)
        for r_, c_, b_ in zip(row1.numpy(), col1.numpy(), scaled.numpy()):
            ref[8 * r_:8 * r_ + 8, 8 * c_:8 * c_ + 8] = b_
        assert np.array_equal(A1s, ref)
        for I in range(n1):
            assert np.abs(A1s[8 * I:8 * I + 8, 8 * I:8 * I + 8] - np.eye(8)).max() < 1e-6       # unscaled K: cond ~ 1e10
        live = np.ones(8 * n1, bool); live[7::8] = False
        Dfull = sp.block_diag([np.linalg.inv(d) for d in Dinv.numpy()]).toarray()
        assert np.abs((Dfull @ A1s - A1)[np.ix_(live, live)]).max() <= 1e-7 * np.abs(A1).max()
        # owned nodes grouped by aggregate
        ptr, order = A["agg_ptr"].numpy(), A["order"].numpy()
        assert ptr[-1] == lp.n_own and np.array_equal(np.sort(order), np.arange(lp.n_own))
        assert np.all(np.repeat(np.arange(n1), np.diff(ptr)) == A["node_agg"].numpy()[order])
        open(os.path.join(result_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_assembly_spmv_and_reductions(world, tmp_path):
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f"ok{r}")) for r in range(world))


def test_dirichlet_and_load_restriction_to_local_nodes():
    """Global boundary data -> local masks / weights (the distributed path never relies on local triangles)."""
    sys.path.insert(0, ROOT)
    from thermoelasticity_fem_b200.distributed import dirichlet_node_data, surface_node_weights
    from thermoelasticity_fem_b200.partition import localize, slab_ranges
    from thermoelasticity_fem_b200.synthetic import kuhn_box, kuhn_node_coords, kuhn_slab_tets
    from oracle import tefem_oracle as orc
    n = 5
    pts, tets, tris, nodes = kuhn_box(n, n, n, 1.0, 1.0, 1.0, n_layers=1)
    dU = {4: [('x', 0.), ('y', 0.5), ('z', 0.)], 5: [('y', 0.), ('z', 0.)]}
    dT = {4: 0., 5: 2.0}
    dofs, fill, lift = dirichlet_node_data({4: tris[4], 5: tris[5]}, dU, dT)
    maps = orc.free_dofs_lists(pts.shape[0], tris, dU, dT)
    assert np.array_equal(dofs, np.sort(np.concatenate([maps["dirichlet_dofs_U"], maps["dirichlet_dofs_theta"]])))
    X = np.zeros(4 * pts.shape[0])
    orc.fill_dirichlet(X, tris, dU, dT)
    assert np.array_equal(X[dofs], fill) and np.array_equal(fill, lift)
    uniq, w = surface_node_weights(lambda ids: kuhn_node_coords(ids, n, n, n, 1., 1., 1.), tris[5])
    F = orc.assemble_F(pts, dict(nodes=nodes, tri=tris, tet=tets), dict(surface={5: np.array([1.0, 0.0, 0.0])}))
    assert np.allclose(F[4 * uniq], w, rtol=1e-14, atol=0)
    # slab generation == global generation restricted to the rank
    world = 3
    bounds = slab_ranges(n + 1, (n + 1) ** 2, world)
    conn = tets[1]
    for rank in range(world):
        p0, p1 = bounds[rank] // (n + 1) ** 2, bounds[rank + 1] // (n + 1) ** 2
        cg, ids = kuhn_slab_tets(n, n, n, p0 - 1, p1)
        a = localize(cg, ids, bounds, rank)
        b = localize(conn, np.arange(conn.shape[0]), bounds, rank)
        assert np.array_equal(a.conn, b.conn) and np.array_equal(a.elem_ids, b.elem_ids)
        assert np.array_equal(a.local_nodes, b.local_nodes) and np.array_equal(a.send_idx, b.send_idx)


def test_longest_interior_range():
    from thermoelasticity_fem_b200.distributed import longest_interior_range
    assert longest_interior_range(np.array([0, 1, 2, 97, 98, 99]), 100) == (3, 97)
    assert longest_interior_range(np.array([], dtype=np.int64), 10) == (0, 10)
    assert longest_interior_range(np.array([5]), 10) == (0, 5)
    lo, hi = longest_interior_range(np.arange(10), 10)
    assert lo == hi
    assert longest_interior_range(np.array([0, 1, 2, 3]), 10) == (4, 10)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_cycle_masks_cover_every_read_of_the_row_partitioned_coarse_cycle(world):
    """Emulation of the data movement of csrc/tefem_precond.cu coarse_cycle_kernel: every rank keeps replicas of the coarse
    vectors that start as NaN; the owner of a row stores it only into the replicas named by the masks of
    multilevel.cycle_masks; every rank then computes its own rows from ITS replicas.  No NaN may be read and the result
    must equal the cycle computed in one piece (overlapped form, 3 + 3 sweeps, level 2, prolongation reads)."""
    import scipy.sparse as sp
    from thermoelasticity_fem_b200.multilevel import cycle_masks, row_owner
    rng = np.random.default_rng(7 + world)
    g = 5                                                    # 5 x 5 x 5 aggregates, 27-point coupling
    n1 = g ** 3
    ijk = np.stack(np.meshgrid(np.arange(g), np.arange(g), np.arange(g), indexing="ij"), axis=-1).reshape(-1, 3)
    near = np.abs(ijk[:, None, :] - ijk[None, :, :]).max(axis=2) <= 1
    pat = sp.csr_array(np.kron(near, np.ones((2, 2))))       # pseudo-node pattern, 2 per aggregate
    rowptr1 = torch.from_numpy(pat.indptr.astype(np.int32))
    colidx1 = torch.from_numpy(pat.indices.astype(np.int32))
    A = sp.bsr_array((0.05 * rng.standard_normal((pat.nnz, 4, 4)), pat.indices, pat.indptr), shape=(8 * n1, 8 * n1)).tocsr()
    A = A + sp.identity(8 * n1, format="csr")
    agg2 = ((ijk[:, 0] // 3) * 2 + ijk[:, 1] // 3) * 2 + ijk[:, 2] // 3
    n2 = int(agg2.max()) + 1
    P2 = sp.csr_array((np.ones(8 * n1), (np.arange(8 * n1), 8 * np.repeat(agg2, 8) + np.tile(np.arange(8), n1))), shape=(8 * n1, 8 * n2))
    inv2 = np.linalg.inv((P2.T @ A @ P2).toarray())
    # fine nodes: slabs along x that do NOT line up with the coarse row partition
    fine_rank_of_agg = [set(int(q) for q in rng.integers(0, world, size=2)) for _ in range(n1)]
    agg_ranks = torch.tensor([sum(1 << q for q in s) for s in fine_rank_of_agg], dtype=torch.uint8)
    x_mask, t_mask = cycle_masks(rowptr1, colidx1, n1, agg2, n2, agg_ranks, world)
    xm, tm = x_mask.numpy(), t_mask.numpy()
    own1, own2 = row_owner(n1, world).numpy(), row_owner(n2, world).numpy()
    wc = 0.7
    bc = rng.standard_normal(8 * n1)
    # one-piece cycle
    x = wc * bc
    x = x + wc * (bc - A @ x)
    t = bc - A @ x
    x = x + wc * t + P2 @ (inv2 @ (P2.T @ t))
    for _ in range(3):
        x = x + wc * (bc - A @ x)
    # partitioned emulation
    rows_of = [np.flatnonzero(np.repeat(own1, 8) == q) for q in range(world)]
    rows2_of = [np.flatnonzero(np.repeat(own2, 8) == q) for q in range(world)]

    def publish(values_by_rank, mask, n):
        reps = [np.full(8 * n, np.nan) for _ in range(world)]
        for q in range(world):
            for dst in range(world):
                sel = rows_of[q] if n == n1 else rows2_of[q]
                if mask is not None:
                    sel = sel[((mask[sel // 8] >> dst) & 1).astype(bool)]
                reps[dst][sel] = values_by_rank[q][sel]
        return reps

    def full(vals_rows):                                      # rank q computed only its rows
        out = [np.zeros(8 * n1) for _ in range(world)]
        for q in range(world):
            out[q][rows_of[q]] = vals_rows[q]
        return out

    def sweep(reps, extra=None):
        outs = []
        for q in range(world):
            r = rows_of[q]
            Ax = A[r] @ np.where(np.isnan(reps[q]), np.nan, reps[q])     # NaN propagates if an unsent value is read
            cols = np.unique(A[r].indices)
            assert not np.isnan(reps[q][cols]).any(), "a level-1 row reads a value that was not sent to its rank"
            outs.append(Ax)
        return outs

    X = publish(full([wc * bc[rows_of[q]] for q in range(world)]), xm, n1)
    Ax = sweep(X)
    X = publish(full([X[q][rows_of[q]] + wc * (bc[rows_of[q]] - Ax[q]) for q in range(world)]), xm, n1)
    Ax = sweep(X)
    T = publish(full([bc[rows_of[q]] - Ax[q] for q in range(world)]), tm, n1)
    r2_rows = []
    for q in range(world):
        r = rows2_of[q]
        cols = np.unique(P2.T.tocsr()[r].indices)
        assert not np.isnan(T[q][cols]).any(), "P2^T reads a residual row that was not sent to the level-2 owner"
        r2_rows.append(P2.T.tocsr()[r] @ T[q])
    r2 = np.zeros(8 * n2)
    for q in range(world):
        r2[rows2_of[q]] = r2_rows[q]                          # r2 and x2 go to every rank
    x2 = inv2 @ r2
    new = []
    for q in range(world):
        r = rows_of[q]
        assert not np.isnan(T[q][r]).any() and not np.isnan(X[q][r]).any()
        new.append(X[q][r] + wc * T[q][r] + P2[r] @ x2)
    X = publish(full(new), xm, n1)
    for _ in range(3):
        Ax = sweep(X)
        X = publish(full([X[q][rows_of[q]] + wc * (bc[rows_of[q]] - Ax[q]) for q in range(world)]), xm, n1)
    for q in range(world):
        r = rows_of[q]
        assert np.abs(X[q][r] - x[r]).max() <= 1e-12 * np.abs(x).max()
        need = np.flatnonzero(np.repeat(np.array([q in s for s in fine_rank_of_agg]), 8))     # prolongation reads
        assert not np.isnan(X[q][need]).any(), "the prolongation reads a coarse value that was not sent to its rank"
        assert np.abs(X[q][need] - x[need]).max() <= 1e-12 * np.abs(x).max()


def test_row_owner_matches_the_kernel_side_formula():
    """multilevel.row_owner (destination masks, Python) and the owner the restriction kernel computes for an aggregate
    (csrc/tefem_precond.cu restrict_kernel: ((i + 1) world - 1) / n, integer division) must agree, and both must describe
    the contiguous ranges [n q / world, n (q + 1) / world) that apply_persistent hands to the cycle kernel."""
    from thermoelasticity_fem_b200.multilevel import row_owner
    for n in (1, 2, 7, 8, 27, 125, 511, 3375, 13824):
        for world in (1, 2, 3, 4, 8):
            if n < world:
                continue
            own = row_owner(n, world).tolist()
            assert own == [((i + 1) * world - 1) // n for i in range(n)]
            for q in range(world):
                lo, hi = n * q // world, n * (q + 1) // world
                assert all(o == q for o in own[lo:hi]) and own.count(q) == hi - lo
