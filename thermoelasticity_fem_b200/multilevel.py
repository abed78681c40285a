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


class TwoLevelPreconditioner:
    """Owns the device arrays of one preconditioner and its C handle (attach with KrylovSolver.set_precond)."""

    def __init__(self, dm, sys, comm=None, factor1=6.0, factor2=0.0, omega=1.0, omega_c=0.8, dense_limit=300):
        import os
        if "TEFEM_ML" in os.environ:          # experiments: "factor1,factor2,omega,omega_c"
            factor1, factor2, omega, omega_c = (float(v) for v in os.environ["TEFEM_ML"].split(","))
        lib, dev = dm.lib, dm.device
        S = dm.schedule
        n_own, n_loc = S.n_rows, S.n_nodes
        xyz = dm.coords
        f64 = dict(dtype=torch.float64, device=dev)
        # ---- global bounding box, mesh size, aggregate grid
        own = xyz[:n_own]
        lo = _allreduce(own.min(dim=0).values.clone(), "MIN")
        hi = _allreduce(own.max(dim=0).values.clone(), "MAX")
        n_glob = int(_allreduce(torch.tensor([float(n_own)], **f64), "SUM").item())
        ext = (hi - lo).clamp(min=1e-300)
        h = float((ext.prod() / max(n_glob, 1)) ** (1.0 / 3.0))
        H1 = factor1 * h
        dims1 = torch.ceil(ext / H1 + 1e-9).to(torch.int64).clamp(min=1)
        key = _grid_keys(xyz, lo, H1, dims1)                                   # all local nodes (owned + halo)
        uniq_local = torch.unique(key[:n_own]).cpu().numpy()
        uniq = np.unique(np.concatenate(_gather_objects(uniq_local)))          # same on every rank
        uniq_t = torch.from_numpy(uniq).to(dev)
        n1 = int(uniq.shape[0])
        agg = torch.searchsorted(uniq_t, key)                                  # aggregate of every local node
        node_agg = agg[:n_own]
        order = torch.sort(node_agg, stable=True)[1]
        agg_ptr = torch.zeros(n1 + 1, dtype=torch.int64, device=dev)
        torch.cumsum(torch.bincount(node_agg, minlength=n1), 0, out=agg_ptr[1:])
        # ---- level-1 Galerkin operator: sum of the fine 4x4 blocks between aggregates (owned rows of this rank)
        I = agg[S.block_row.long()]
        J = agg[S.colidx.long()]
        k2 = I * n1 + J
        uk, inv = torch.unique(k2, return_inverse=True)
        part = torch.zeros((uk.numel(), 16), **f64).index_add_(0, inv, sys)
        del I, J, k2, inv
        parts = _gather_objects((uk.cpu().numpy(), part.cpu().numpy())) if _dist() is not None else None
        if parts is not None:
            keys = torch.from_numpy(np.concatenate([p[0] for p in parts])).to(dev)
            vals = torch.from_numpy(np.concatenate([p[1] for p in parts])).to(dev)
            uk, inv = torch.unique(keys, return_inverse=True)
            part = torch.zeros((uk.numel(), 16), **f64).index_add_(0, inv, vals)
            del keys, vals, inv
        row1 = torch.div(uk, n1, rounding_mode='floor')
        col1 = uk - row1 * n1
        rowptr1 = torch.zeros(n1 + 1, dtype=torch.int64, device=dev)
        torch.cumsum(torch.bincount(row1, minlength=n1), 0, out=rowptr1[1:])
        diag1 = torch.nonzero(row1 == col1).reshape(-1)
        assert int(diag1.numel()) == n1, "every aggregate must have a diagonal block"
        i32 = torch.int32
        self.rowptr1, self.colidx1, self.diag1 = rowptr1.to(i32), col1.to(i32), diag1.to(i32)
        self.sys1 = part.contiguous()
        self.dinv1 = torch.empty((n1, 16), **f64)
        p = _lib.ptr
        _lib.check(lib.tefem_block_jacobi_scale(p(self.rowptr1), p(self.diag1), n1, p(self.sys1), p(self.dinv1),
                                                _lib.stream_ptr()))
        # ---- level 2: dense inverse of P2^T A1^ P2 (host, a few hundred nodes)
        if n1 <= dense_limit:
            agg2 = np.arange(n1, dtype=np.int64)
            n2 = n1
        else:
            k = uniq
            ijk = np.stack([k // (int(dims1[1]) * int(dims1[2])), (k // int(dims1[2])) % int(dims1[1]), k % int(dims1[2])], axis=1)
            if factor2 <= 0:      # automatic: keep the dense level-2 operator at a few hundred nodes
                factor2 = max(3.0, float(np.ceil((n1 / 450.0) ** (1.0 / 3.0))))
            c = np.floor(ijk / factor2 + 1e-9).astype(np.int64)
            d2 = c.max(axis=0) + 1
            key2 = (c[:, 0] * d2[1] + c[:, 1]) * d2[2] + c[:, 2]
            _, agg2 = np.unique(key2, return_inverse=True)
            n2 = int(agg2.max()) + 1
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
