"""Host-side setup of the aggregation preconditioner (csrc/tefem_precond.cu), once per system matrix.

Builds, with torch integer / scatter ops on the device (not on the repeatable path):
  * a geometric aggregation of the nodes: cubes of edge factor1 * h on a grid anchored at the GLOBAL bounding box
    (every rank computes the same aggregate for a node from its coordinates; aggregates may straddle ranks);
  * the Galerkin operator of level 1, A1 = P1^T A^ P1, from the block-Jacobi-scaled fine system (each rank sums the
    blocks of its owned rows; the partial operators are gathered and summed: level 1 is replicated on every rank),
    its block-Jacobi scaling, and the dense inverse of the level-2 Galerkin operator (factor2 coarsening of level 1).

The reference has no counterpart (it calls a direct solver, solvers.py:26,172).
"""
import ctypes

import numpy as np
import torch

from . import _lib


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) else None


def _allreduce(t, op):
    dist = _dist()
    if dist is not None:
        dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
    return t


def _gather_objects(obj):
    dist = _dist()
    if dist is None:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def _grid_keys(xyz, lo, H, dims):
    ijk = torch.floor((xyz - lo) / H + 1e-9).to(torch.int64)
    ijk = torch.minimum(ijk.clamp(min=0), (dims - 1).to(ijk.dtype))
    return (ijk[:, 0] * dims[1] + ijk[:, 1]) * dims[2] + ijk[:, 2]


def build_aggregation(xyz, n_own, factor1):
    """Geometric aggregation of the nodes: cubes of edge factor1 * h on a grid anchored at the GLOBAL bounding box
    (h = mean node spacing).  xyz: [n_local, 3] coordinates (owned nodes first).  Returns a dict with the sorted
    global aggregate keys `uniq`, the grid `dims1`, the aggregate of every local node `agg`, and the owned nodes
    grouped by aggregate (`agg_ptr`, `order`).  Collective when torch.distributed is initialised."""
    dev = xyz.device
    f64 = dict(dtype=torch.float64, device=dev)
    own = xyz[:n_own]
    lo = _allreduce(own.min(dim=0).values.clone(), "MIN")
    hi = _allreduce(own.max(dim=0).values.clone(), "MAX")
    n_glob = int(_allreduce(torch.tensor([float(n_own)], **f64), "SUM").item())
    ext = (hi - lo).clamp(min=1e-300)
    h = float(ext.prod().item() / max(n_glob, 1)) ** (1.0 / 3.0)
    H1 = factor1 * h
    dims1 = torch.ceil(ext / H1 + 1e-9).to(torch.int64).clamp(min=1)
    key = _grid_keys(xyz, lo, H1, dims1)                                   # all local nodes (owned + halo)
    uniq_local = torch.unique(key[:n_own]).cpu().numpy()
    uniq = np.unique(np.concatenate(_gather_objects(uniq_local)))          # same on every rank
    n1 = int(uniq.shape[0])
    agg = torch.searchsorted(torch.from_numpy(uniq).to(dev), key)          # aggregate of every local node
    node_agg = agg[:n_own]
    order = torch.sort(node_agg, stable=True)[1]
    agg_ptr = torch.zeros(n1 + 1, dtype=torch.int64, device=dev)
    torch.cumsum(torch.bincount(node_agg, minlength=n1), 0, out=agg_ptr[1:])
    return dict(h=h, H1=H1, dims1=dims1, uniq=uniq, n1=n1, agg=agg, node_agg=node_agg, order=order, agg_ptr=agg_ptr)


def galerkin_level1(agg, block_row, colidx, sys, n1):
    """A1 = P1^T A P1 for the piecewise-constant prolongation of `agg`: the 4x4 blocks of the fine rows of this rank
    are summed per (aggregate, aggregate) pair; the partial operators of all ranks are gathered and summed (level 1
    is replicated).  Returns (rowptr1, col1, diag1, blocks [nnz1, 16]) as int64 / fp64 tensors."""
    dev = sys.device
    f64 = dict(dtype=torch.float64, device=dev)
    k2 = agg[block_row.long()] * n1 + agg[colidx.long()]
    uk, inv = torch.unique(k2, return_inverse=True)
    part = torch.zeros((uk.numel(), 16), **f64).index_add_(0, inv, sys)
    del k2, inv
    if _dist() is not None:
        parts = _gather_objects((uk.cpu().numpy(), part.cpu().numpy()))
        keys = torch.from_numpy(np.concatenate([p[0] for p in parts])).to(dev)
        vals = torch.from_numpy(np.concatenate([p[1] for p in parts])).to(dev)
        uk, inv = torch.unique(keys, return_inverse=True)
        part = torch.zeros((uk.numel(), 16), **f64).index_add_(0, inv, vals)
    row1 = torch.div(uk, n1, rounding_mode='floor')
    col1 = uk - row1 * n1
    rowptr1 = torch.zeros(n1 + 1, dtype=torch.int64, device=dev)
    torch.cumsum(torch.bincount(row1, minlength=n1), 0, out=rowptr1[1:])
    diag1 = torch.nonzero(row1 == col1).reshape(-1)
    assert int(diag1.numel()) == n1, "every aggregate must have a diagonal block"
    return rowptr1, col1, diag1, part.contiguous()


def build_level2(uniq, dims1, factor2, dense_limit=300):
    """Aggregation of the level-1 nodes (grid cells `uniq` of the dims1 grid) into cubes of factor2 cells; identity
    when level 1 is small enough for the dense solve.  factor2 <= 0: chosen so that level 2 has a few hundred nodes."""
    n1 = int(uniq.shape[0])
    if n1 <= dense_limit:
        return np.arange(n1, dtype=np.int64), n1, 1.0
    d1, d2_ = int(dims1[1]), int(dims1[2])
    ijk = np.stack([uniq // (d1 * d2_), (uniq // d2_) % d1, uniq % d2_], axis=1)
    if factor2 <= 0:
        factor2 = max(3.0, float(np.ceil((n1 / 450.0) ** (1.0 / 3.0))))
    c = np.floor(ijk / factor2 + 1e-9).astype(np.int64)
    d2 = c.max(axis=0) + 1
    key2 = (c[:, 0] * d2[1] + c[:, 1]) * d2[2] + c[:, 2]
    _, agg2 = np.unique(key2, return_inverse=True)
    return agg2.astype(np.int64), int(agg2.max()) + 1, factor2


class TwoLevelPreconditioner:
    """Owns the device arrays of one preconditioner and its C handle (attach with KrylovSolver.set_precond)."""

    def __init__(self, dm, sys, comm=None, factor1=6.0, factor2=0.0, omega=1.0, omega_c=0.8, dense_limit=300):
        import os
        if "TEFEM_ML" in os.environ:          # experiments: "factor1,factor2,omega,omega_c"
            factor1, factor2, omega, omega_c = (float(v) for v in os.environ["TEFEM_ML"].split(","))
        lib, dev = dm.lib, dm.device
        S = dm.schedule
        n_own = S.n_rows
        f64 = dict(dtype=torch.float64, device=dev)
        i32 = torch.int32
        p = _lib.ptr
        A = build_aggregation(dm.coords, n_own, factor1)
        n1, uniq, dims1, h, H1 = A["n1"], A["uniq"], A["dims1"], A["h"], A["H1"]
        node_agg, order, agg_ptr = A["node_agg"], A["order"], A["agg_ptr"]
        rowptr1, col1, diag1, part = galerkin_level1(A["agg"], S.block_row, S.colidx, sys, n1)
        self.rowptr1, self.colidx1, self.diag1 = rowptr1.to(i32), col1.to(i32), diag1.to(i32)
        self.sys1 = part
        self.dinv1 = torch.empty((n1, 16), **f64)
        _lib.check(lib.tefem_block_jacobi_scale(p(self.rowptr1), p(self.diag1), n1, p(self.sys1), p(self.dinv1),
                                                _lib.stream_ptr()))
        # ---- level 2: dense inverse of P2^T A1^ P2 (host, a few hundred nodes)
        agg2, n2, factor2 = build_level2(uniq, dims1, factor2, dense_limit)
        self.factor2 = factor2
        import scipy.sparse as sp
        A1 = sp.bsr_array((self.sys1.cpu().numpy().reshape(-1, 4, 4), self.colidx1.cpu().numpy(), self.rowptr1.cpu().numpy()),
                          shape=(4 * n1, 4 * n1)).tocsr()
        rows = (4 * np.arange(n1)[:, None] + np.arange(4)).ravel()
        cols = (4 * agg2[:, None] + np.arange(4)).ravel()
        P2 = sp.csr_array((np.ones(4 * n1), (rows, cols)), shape=(4 * n1, 4 * n2))
        A2 = (P2.T @ A1 @ P2).toarray()
        # preconditioner data of the coarse levels is kept in single precision (half the bytes, L2-resident)
        self.inv2 = torch.from_numpy(np.ascontiguousarray(np.linalg.inv(A2)).astype(np.float32)).to(dev)
        self.sys1 = self.sys1.to(torch.float32).contiguous()
        order2 = np.argsort(agg2, kind="stable")
        agg2_ptr = np.zeros(n2 + 1, dtype=np.int64)
        np.cumsum(np.bincount(agg2, minlength=n2), out=agg2_ptr[1:])
        self.agg_ptr, self.agg_nodes, self.node_agg = agg_ptr.to(i32), order.to(i32), node_agg.to(i32)
        self.agg2_ptr = torch.from_numpy(agg2_ptr.astype(np.int32)).to(dev)
        self.agg2_nodes = torch.from_numpy(order2.astype(np.int32)).to(dev)
        self.node_agg2 = torch.from_numpy(agg2.astype(np.int32)).to(dev)
        self.n1, self.n2, self.h, self.H1 = n1, n2, h, H1
        self.factor1, self.omega, self.omega_c = factor1, omega, omega_c
        if comm is not None:
            comm.setup_peer(4 * n1)               # the restricted residual is all-reduced through peer memory
        self.work = torch.zeros(int(lib.tefem_precond_work_doubles(n1, n2)), **f64)
        self.handle = ctypes.c_void_p()
        self.lib = lib
        _lib.check(lib.tefem_precond_create(
            ctypes.byref(self.handle), n_own, n1, p(self.agg_ptr), p(self.agg_nodes), p(self.node_agg), float(omega),
            p(self.rowptr1), p(self.colidx1), p(self.sys1), p(self.dinv1), float(omega_c), n2, p(self.agg2_ptr),
            p(self.agg2_nodes), p(self.node_agg2), p(self.inv2), p(self.work), comm.handle if comm is not None else None))

    def apply(self, r, z):
        _lib.check(self.lib.tefem_precond_apply(self.handle, _lib.ptr(r), _lib.ptr(z), _lib.stream_ptr()))
        return z

    def close(self):
        if self.handle:
            self.lib.tefem_precond_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
