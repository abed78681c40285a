"""scipy prototype of the aggregation preconditioner (RESEARCH / TEST INFRASTRUCTURE, never imported by the package).

Builds the block-Jacobi-scaled Newmark operator of BASELINE config 4's physics on an n^3 Kuhn box with the CPU oracle
and counts right-preconditioned BiCGSTAB iterations for coarse-space / cycle variants, to decide what is worth building
in CUDA (the same way the rigid-body coarse space was chosen, profiles/README.md).

usage: python tests/prototypes/precond_proto.py [cells] [variant ...]
"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import tefem_oracle as orc                                   # noqa: E402
from thermoelasticity_fem_b200.synthetic import kuhn_box                  # noqa: E402

MAT2 = (2300, 64e9, 0.1, 1., 1., 4e-6, 20.)
DU = {4: [('x', 0.), ('y', 0.), ('z', 0.)], 5: [('y', 0.), ('z', 0.)]}
DT = {4: 0., 5: 0.}


def build_system(cells):
    pts, tets, tris, nodes = kuhn_box(cells, cells, cells, 1.0, 1.0, 1.0, n_layers=1)
    groups = dict(tet=tets, tri=tris, nodes=nodes)
    m = orc.Material(*MAT2)
    conn, mat_id, rows = orc.element_table(tets, {1: m})
    ops = orc.assemble(pts, conn, rows, mat_id, 0.2, 0.2, which="MKD")
    n_nodes = pts.shape[0]
    n = 4 * n_nodes
    maps = orc.free_dofs_lists(n_nodes, tris, DU, DT)
    dt = 1800.0 / 999
    A = (ops["M"] + 0.5 * dt * ops["D"] + 0.25 * dt ** 2 * ops["K"]).tocsr()
    free = np.zeros(n, dtype=bool)
    free[maps["free_dofs"]] = True
    Pf = sp.diags(free.astype(float))
    A = (Pf @ A @ Pf + sp.diags((~free).astype(float))).tocsr()          # eliminated DOFs: identity rows / columns
    # 4x4 block-Jacobi scaling (left)
    Ab = A.tobsr(blocksize=(4, 4))
    Ab.sort_indices()
    diag = np.zeros((n_nodes, 4, 4))
    for i in range(n_nodes):
        s, t = Ab.indptr[i], Ab.indptr[i + 1]
        k = s + np.searchsorted(Ab.indices[s:t], i)
        diag[i] = Ab.data[k]
    dinv = np.linalg.inv(diag)
    Dinv = sp.bsr_array((dinv, np.arange(n_nodes), np.arange(n_nodes + 1)), shape=(n, n)).tocsr()
    Ahat = (Dinv @ A).tocsr()
    vec_t = np.linspace(0., 1800.0, 1000)
    arr = np.zeros((3, 1000))
    arr[0, :] = np.sin(2 * np.pi * vec_t / 900.) * -1e9
    F = orc.assemble_F(pts, groups, dict(surface={5: arr}), idx_t=250)
    b = Dinv @ (F * free)
    return pts, Ahat, b, free


def aggregates(pts, cells, f):
    """cubes of f cells anchored at the origin (the product's rule, multilevel.build_aggregation)."""
    h = 1.0 / cells
    ijk = np.minimum(np.floor(pts / (f * h) + 1e-9).astype(np.int64), (cells // f + 1))
    dims = ijk.max(axis=0) + 1
    key = (ijk[:, 0] * dims[1] + ijk[:, 1]) * dims[2] + ijk[:, 2]
    uniq, agg = np.unique(key, return_inverse=True)
    n1 = uniq.size
    cen = np.stack([np.bincount(agg, weights=pts[:, k], minlength=n1) / np.bincount(agg, minlength=n1) for k in range(3)], 1)
    return agg, cen, n1


def prolongation(pts, agg, cen, n1, free, modes="rbm", orthonormalize=True):
    """Tentative prolongation.  rbm: 3 translations, 3 rotations, theta (7 live columns of 8).  lin: + 6 linear strain
    modes of u and 3 gradient modes of theta (16 columns)."""
    n = pts.shape[0]
    d = pts - cen[agg]
    ar = np.arange(n)
    ncol = 8 if modes == "rbm" else 16
    rows, cols, vals = [], [], []

    def add(c, k, v):
        rows.append(4 * ar + c); cols.append(ncol * agg + k); vals.append(v)
    one = np.ones(n)
    for c in range(3):
        add(c, c, one)
    add(3, 3, one)
    dx, dy, dz = d[:, 0], d[:, 1], d[:, 2]
    for (c, k, v) in [(0, 1, dz), (0, 2, -dy), (1, 2, dx), (1, 0, -dz), (2, 0, dy), (2, 1, -dx)]:
        add(c, 4 + k, v)
    if modes == "lin":
        add(0, 8, dx); add(1, 9, dy); add(2, 10, dz)                      # eps_xx, eps_yy, eps_zz
        add(0, 11, dy); add(1, 11, dx)                                    # gamma_xy
        add(1, 12, dz); add(2, 12, dy)                                    # gamma_yz
        add(0, 13, dz); add(2, 13, dx)                                    # gamma_xz
        # theta gradient (3) shares columns 14, 15 and 7
        add(3, 14, dx); add(3, 15, dy); add(3, 7, dz)
    P = sp.csr_array((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(4 * n, ncol * n1))
    P = (sp.diags(free.astype(float)) @ P).tocsr()
    if not orthonormalize:
        return P
    # per aggregate: orthonormal basis of the span of its columns, numerically dependent directions dropped (zero columns);
    # the u and theta parts are treated separately so that the level-2 rule (columns 0-2, 4-6 = t, w; 3 = theta) stays valid
    Pc = P.tocsc()
    out_r, out_c, out_v = [], [], []
    for I in range(n1):
        nodes_I = np.nonzero(agg == I)[0]
        for comps, colset in (((0, 1, 2), [c for c in range(ncol) if c not in (3, 7, 14, 15)]), ((3,), [c for c in (3, 7, 14, 15) if c < ncol])):
            r_idx = (4 * nodes_I[:, None] + np.array(comps)[None, :]).reshape(-1)
            c_idx = ncol * I + np.array(colset)
            B = Pc[:, c_idx][r_idx, :].toarray()
            if modes == "rbm":
                Q, keepc = B, np.ones(len(colset), dtype=bool)              # product rule: no re-basing of the rigid-body modes
            else:
                # keep the rigid-body / constant columns as they are, orthogonalise only the extra modes against them
                base = [k for k, c in enumerate(colset) if c < 8 and c != 7]
                extra = [k for k in range(len(colset)) if k not in base]
                Q = B.copy()
                if extra:
                    Bb = B[:, base]
                    coef = np.linalg.lstsq(Bb, B[:, extra], rcond=None)[0]
                    E = B[:, extra] - Bb @ coef
                    U, sv, _ = np.linalg.svd(E, full_matrices=False)
                    good = sv > 1e-8 * max(sv.max(), 1e-300) if sv.size else np.zeros(0, dtype=bool)
                    Q[:, extra] = 0.0
                    scale = np.abs(Bb).max() if Bb.size else 1.0
                    Q[:, [extra[k] for k in range(len(extra)) if k < good.size and good[k]]] = U[:, good] * scale
            rr, cc = np.nonzero(Q)
            out_r.append(r_idx[rr]); out_c.append(c_idx[cc]); out_v.append(Q[rr, cc])
    return sp.csr_array((np.concatenate(out_v), (np.concatenate(out_r), np.concatenate(out_c))), shape=P.shape)


def level2_transfer(cen, agg2, cen2, n2, modes):
    """Level 1 -> 2 with the same rigid-body rule between the centres: t_I = T_J + W_J x D_I, theta_I = Theta_J, w_I = W_J;
    the extra linear modes (if any) are aggregated piecewise-constant mode by mode."""
    n1 = cen.shape[0]
    ncol = 8 if modes == "rbm" else 16
    D = cen - cen2[agg2]
    ar = np.arange(n1)
    rows, cols, vals = [], [], []

    def add(c, k, v):
        rows.append(ncol * ar + c); cols.append(ncol * agg2 + k); vals.append(v)
    one = np.ones(n1)
    for c in range(ncol):
        add(c, c, one)
    dx, dy, dz = D[:, 0], D[:, 1], D[:, 2]
    for (c, k, v) in [(0, 1, dz), (0, 2, -dy), (1, 2, dx), (1, 0, -dz), (2, 0, dy), (2, 1, -dx)]:
        add(c, 4 + k, v)
    return sp.csr_array((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(ncol * n1, ncol * n2))


def block_diag_inv(A1, bs):
    n1 = A1.shape[0] // bs
    if A1.shape[0] % bs:
        return sp.diags(1.0 / A1.diagonal())
    B = A1.tobsr(blocksize=(bs, bs))
    B.sort_indices()
    blocks = np.zeros((n1, bs, bs))
    for i in range(n1):
        s, t = B.indptr[i], B.indptr[i + 1]
        blocks[i] = B.data[s + np.searchsorted(B.indices[s:t], i)]
    return sp.bsr_array((np.linalg.inv(blocks), np.arange(n1), np.arange(n1 + 1)), shape=A1.shape).tocsr()


class Precond:
    def __init__(self, A, P, kind="exact", sweeps=3, omega_c=0.7, P2=None, smooth_P=0.0, fine_sweeps=0, omega=1.0):
        if smooth_P > 0.0:
            dA = A.diagonal()
            P = (P - smooth_P * sp.diags(1.0 / dA) @ (A @ P)).tocsr()
        self.A, self.P, self.kind, self.sweeps, self.omega_c, self.omega = A, P, kind, sweeps, omega_c, omega
        A1 = (P.T @ A @ P).tocsr()
        dead = A1.diagonal() == 0.0                                       # columns without a free DOF
        A1 = (A1 + sp.diags(dead.astype(float))).tocsr()
        self.nnz1 = A1.nnz
        if kind == "exact":
            self.lu = spla.splu(A1.tocsc())
        else:
            dA1 = A1.diagonal()
            self.D1 = sp.diags(1.0 / np.where(dA1 != 0, dA1, 1.0))      # point Jacobi on the Galerkin operator
            self.A1 = A1
            self.P2 = P2
            if P2 is not None:
                A2 = (P2.T @ A1 @ P2).tocsr()
                A2 = (A2 + sp.diags((A2.diagonal() == 0.0).astype(float))).tocsc()
                self.lu2 = spla.splu(A2)
        self.n_apply = 0

    def coarse(self, rc):
        if self.kind == "exact":
            return self.lu.solve(rc)
        bc = self.D1 @ rc
        x = self.omega_c * bc
        for _ in range(self.sweeps - 1):
            x = x + self.omega_c * (bc - self.D1 @ (self.A1 @ x))
        if self.P2 is not None:
            x = x + self.P2 @ self.lu2.solve(self.P2.T @ (rc - self.A1 @ x))
        for _ in range(self.sweeps):
            x = x + self.omega_c * (bc - self.D1 @ (self.A1 @ x))
        return x

    def __call__(self, r):
        self.n_apply += 1
        return self.omega * r + self.P @ self.coarse(self.P.T @ r)


def bicgstab_iterations(A, b, M, rtol=1e-10, maxiter=3000):
    """Right-preconditioned BiCGSTAB (the product's iteration); returns (iterations, final relative residual)."""
    n = b.size
    x = np.zeros(n)
    r = b.copy()
    rhat = r.copy()
    rho = alpha = omega = 1.0
    v = np.zeros(n)
    p = np.zeros(n)
    bn = np.linalg.norm(b)
    for it in range(1, maxiter + 1):
        rho_new = rhat @ r
        beta = (rho_new / rho) * (alpha / omega)
        rho = rho_new
        p = r + beta * (p - omega * v)
        ph = M(p)
        v = A @ ph
        alpha = rho / (rhat @ v)
        s = r - alpha * v
        sh = M(s)
        t = A @ sh
        omega = (t @ s) / (t @ t)
        x = x + alpha * ph + omega * sh
        r = s - omega * t
        if np.linalg.norm(r) <= rtol * bn:
            return it, np.linalg.norm(b - A @ x) / bn
    return maxiter, np.linalg.norm(b - A @ x) / bn


if __name__ == "__main__":
    cells = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    f = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    t0 = time.time()
    pts, A, b, free = build_system(cells)
    print(f"{cells}^3 box: {A.shape[0]} DOF, nnz {A.nnz}, set-up {time.time() - t0:.1f} s", flush=True)
    it, res = bicgstab_iterations(A, b, lambda r: r)
    print(f"block-Jacobi scaling only: {it} iterations (true residual {res:.1e})", flush=True)
    agg, cen, n1 = aggregates(pts, cells, f)
    agg2, cen2, n2 = aggregates(cen, cells, 3 * f)                        # level 2: 3^3 level-1 aggregates
    variants = [
        ("rbm exact", dict(modes="rbm", kind="exact")),
        ("rbm 3+3 jacobi + level 2 (product)", dict(modes="rbm", kind="cycle")),
        ("lin exact", dict(modes="lin", kind="exact")),
        ("lin 3+3 jacobi + level 2", dict(modes="lin", kind="cycle")),
        ("rbm exact, smoothed P 0.5", dict(modes="rbm", kind="exact", smooth_P=0.5)),
        ("rbm cycle, smoothed P 0.5", dict(modes="rbm", kind="cycle", smooth_P=0.5)),
        ("lin exact, smoothed P 0.5", dict(modes="lin", kind="exact", smooth_P=0.5)),
    ]
    for name, kw in variants:
        t1 = time.time()
        P = prolongation(pts, agg, cen, n1, free, kw["modes"])
        P2 = level2_transfer(cen, agg2, cen2, n2, kw["modes"]) if kw["kind"] == "cycle" else None
        M = Precond(A, P, kind=kw["kind"], P2=P2, smooth_P=kw.get("smooth_P", 0.0))
        it, res = bicgstab_iterations(A, b, M)
        print(f"{name:40s} coarse size {M.P.shape[1]:6d} nnz {M.nnz1:9d}: {it:4d} iterations (true residual {res:.1e}, "
              f"{time.time() - t1:.1f} s)", flush=True)
