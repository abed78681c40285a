"""Device-side engine: owns the torch tensors (device memory) and calls the CUDA library.

This is the thin host layer between the reference-compatible classes (mesh / model / solvers) and
the C ABI of ``include/tefem.h``.  Everything here requires a CUDA device; there is no CPU path.
"""
import ctypes

import numpy as np
import torch

from . import _lib, layout
from .symbolic import Schedule, build_schedule


def _i32(t):
    return t.to(torch.int32).contiguous()


class DeviceMesh:
    """Mesh tables + symbolic schedule resident in HBM.

    coords[N,3] f64, conn[E,4] int32 (reference element order, ``mesh.py:51-58``), mat_id[E] int32,
    mat_table[n_mat,8] f64.  ``n_rows`` (< N) restricts assembly to the owned node rows of a
    row-partitioned multi-GPU run (owned nodes are numbered first)."""

    def __init__(self, coords, conn, mat_id, mat_table, n_rows=None, device=None, schedule=None):
        self.device = device if device is not None else _lib.require_cuda()
        self.lib = _lib.load()
        dev = self.device
        self.coords = torch.as_tensor(coords, dtype=torch.float64).to(dev).contiguous()
        conn_t = torch.as_tensor(conn).to(dev)
        self.n_nodes = int(self.coords.shape[0])
        self.n_elems = int(conn_t.shape[0])
        self.n_rows = self.n_nodes if n_rows is None else int(n_rows)
        self.conn = _i32(conn_t)
        self.mat_id = _i32(torch.as_tensor(mat_id).to(dev))
        self.mat_table = torch.as_tensor(mat_table, dtype=torch.float64).to(dev).contiguous()
        self.n_mat = int(self.mat_table.shape[0])
        if schedule is None:
            schedule = build_schedule(conn_t, self.n_nodes, self.n_rows, mat_id=self.mat_id)
        self.schedule: Schedule = schedule.to(dev)
        del conn_t
        # per-tile copy of the tile nodes' coordinates (the byte-packed connectivity and the material of every
        # work item are part of the schedule): every tile segment is one bulk copy for K1
        S = self.schedule
        self.cta_xyz = self.coords[S.cta_nodes.long()].contiguous()
        self._h_max = S.h_max()
        self.bad_elem = torch.full((2,), _lib.INT32_MAX, dtype=torch.int32, device=dev)
        self.asm_threads = 256          # threads per CTA of K1 (64 | 128 | 256: measurement knob, pair with tiles of 2 x threads blocks)

    # ---- K1 -------------------------------------------------------------------------------------
    def alloc_planes(self, which, alpha_M=0., alpha_K=0.):
        P = self.schedule.n_blocks
        n = {"M": 3, "K": 13, "D": layout.n_planes(layout.plane_map_D(alpha_M, alpha_K))}[which]
        return torch.empty((n, P), dtype=torch.float64, device=self.device)

    def assemble(self, ops, alpha_M=0., alpha_K=0., out=None, check=True):
        """ops: string subset of 'MKD'.  Returns dict name -> planes tensor [n_planes, P]."""
        S = self.schedule
        out = {} if out is None else out
        for nm in ops:
            if nm not in out or out[nm] is None:
                out[nm] = self.alloc_planes(nm, alpha_M, alpha_K)
        mask = sum({"M": _lib.OP_M, "K": _lib.OP_K, "D": _lib.OP_D}[nm] for nm in set(ops))
        self.bad_elem.fill_(_lib.INT32_MAX)
        p = _lib.ptr
        _lib.check(self.lib.tefem_assemble_set_threads(int(self.asm_threads)))
        _lib.check(self.lib.tefem_assemble(
            p(self.mat_table), self.n_mat, S.n_cta, p(self._h_max), p(S.cta_hdr), p(S.cta_elems),
            p(S.cta_conn8), p(self.cta_xyz), p(S.item_inc), p(S.item_meta), p(S.split_block), p(S.split_meta),
            p(S.inc), S.n_blocks, mask, float(alpha_M), float(alpha_K),
            p(out.get("M")) if "M" in ops else None, p(out.get("K")) if "K" in ops else None,
            p(out.get("D")) if "D" in ops else None, p(self.bad_elem), _lib.stream_ptr()))
        if check:
            self.check_jacobian()
        return out

    def check_jacobian(self):
        e = ctypes.c_int32(-1)
        _lib.check(self.lib.tefem_check_jacobian(_lib.ptr(self.bad_elem), ctypes.byref(e), _lib.stream_ptr()))

    def element_matrices(self, kinds=("Muu", "Dtu", "Dtt", "Kuu", "Kut", "Ktt"), elem_idx=None):
        """Dense element matrices (parity path / Element.compute_mat_*_e), as device tensors."""
        dev = self.device
        if elem_idx is None:
            n, idx = self.n_elems, None
        else:
            idx = torch.as_tensor(elem_idx, dtype=torch.int64, device=dev).contiguous()
            n = int(idx.numel())
        shapes = {"Muu": (12, 12), "Dtu": (4, 12), "Dtt": (4, 4), "Kuu": (12, 12), "Kut": (12, 4), "Ktt": (4, 4)}
        out = {k: torch.empty((n,) + shapes[k], dtype=torch.float64, device=dev) for k in kinds}
        mask = sum(_lib.KIND[k] for k in kinds)
        self.bad_elem.fill_(_lib.INT32_MAX)
        p = _lib.ptr
        _lib.check(self.lib.tefem_element_matrices(
            p(self.coords), p(self.conn), p(self.mat_id), p(self.mat_table), self.n_mat, p(idx), n, mask,
            p(out.get("Muu")), p(out.get("Dtu")), p(out.get("Dtt")), p(out.get("Kuu")), p(out.get("Kut")),
            p(out.get("Ktt")), p(self.bad_elem), _lib.stream_ptr()))
        self.check_jacobian()
        return out

    # ---- plane-storage operators -------------------------------------------------------------
    def planes_spmv(self, vals, plane_map, x, y, alpha=1.0, beta_y=0.0):
        S = self.schedule
        pm = np.ascontiguousarray(plane_map, dtype=np.int8).reshape(16)
        p = _lib.ptr
        _lib.check(self.lib.tefem_planes_spmv(p(S.rowptr), p(S.colidx), S.n_rows, S.n_blocks, p(vals), p(pm),
                                              p(x), float(alpha), float(beta_y), p(y), _lib.stream_ptr()))
        return y

    def export_csc(self, vals, plane_map, drop_zeros=False):
        """scipy CSC matrix of an operator in plane storage (host copy; for parity / API compatibility)."""
        import scipy.sparse as sp
        S = self.schedule
        br = S.block_row.cpu().numpy().astype(np.int64)
        bc = S.colidx.cpu().numpy().astype(np.int64)
        v = vals.cpu().numpy()
        rows, cols, data = [], [], []
        for r in range(4):
            for c in range(4):
                pl = int(plane_map[r, c])
                if pl >= 0:
                    rows.append(4 * br + r)
                    cols.append(4 * bc + c)
                    data.append(v[pl])
        n_r, n_c = 4 * S.n_rows, 4 * S.n_nodes
        A = sp.csr_array((np.concatenate(data), (np.concatenate(rows), np.concatenate(cols))), shape=(n_r, n_c)).tocsc()
        A.sort_indices()
        if drop_zeros:
            A.eliminate_zeros()          # what csc_array(dense) does in the reference, model.py:92,103,118
        return A

    # ---- K2 / K5 -----------------------------------------------------------------------------
    def form_system(self, planes, cM, cD, cK, d_layout, dir_mask, sys=None):
        S = self.schedule
        if sys is None:
            sys = torch.empty((S.n_blocks, 16), dtype=torch.float64, device=self.device)
        p = _lib.ptr
        _lib.check(self.lib.tefem_form_system(
            p(S.rowptr), p(S.colidx), S.n_rows, S.n_blocks, p(planes.get("M")), float(cM),
            p(planes.get("D")), int(d_layout), float(cD), p(planes.get("K")), float(cK),
            p(dir_mask), p(sys), _lib.stream_ptr()))
        return sys

    def block_jacobi_scale(self, sys, scale=True):
        """Inverse 4x4 diagonal blocks dinv [n_rows, 16]; scale=True also left-scales sys in place (BiCGSTAB path),
        scale=False keeps sys (the symmetric CG path applies dinv itself)."""
        S = self.schedule
        dinv = torch.empty((S.n_rows, 16), dtype=torch.float64, device=self.device)
        singular = torch.empty(1, dtype=torch.int32, device=self.device)        # scratch the library needs: caller-owned
        p = _lib.ptr
        _lib.check(self.lib.tefem_block_jacobi_scale(p(S.rowptr), p(S.diag_idx), S.n_rows, p(sys), p(dinv), p(singular),
                                                     int(bool(scale)), _lib.stream_ptr()))
        return dinv

    def block_jacobi_apply(self, dinv, r, z, dir_mask=None):
        _lib.check(self.lib.tefem_block_jacobi_apply(_lib.ptr(dinv), self.schedule.n_rows, _lib.ptr(r), _lib.ptr(z),
                                                     _lib.ptr(dir_mask), _lib.stream_ptr()))
        return z

    def spmv(self, sys, x, y):
        S = self.schedule
        p = _lib.ptr
        _lib.check(self.lib.tefem_spmv(p(S.rowptr), p(S.colidx), S.n_rows, p(sys), p(x), p(y), _lib.stream_ptr()))
        return y


class KrylovSolver:
    """Handle of the BiCGSTAB driver (K6) with its torch-owned workspace."""

    def __init__(self, dmesh: DeviceMesh, n_halo=0, comm=None):
        self.dmesh = dmesh
        self.lib = dmesh.lib
        S = dmesh.schedule
        n_work = int(self.lib.tefem_solver_work_doubles(S.n_rows, n_halo))
        self.work = torch.zeros(n_work, dtype=torch.float64, device=dmesh.device)
        self.handle = ctypes.c_void_p()
        self.comm = comm
        _lib.check(self.lib.tefem_solver_create(ctypes.byref(self.handle), S.n_rows, n_halo, _lib.ptr(self.work),
                                                comm.handle if comm is not None else None))
        self.n_halo = n_halo
        self._keep = []

    def set_halo(self, nbr, send_ptr, send_idx, recv_ptr, boundary_rows=None, interior_rows=None, send_dst=None,
                 send_dpos=None):
        """Peer-memory halo (communicator with vector slots): send_dst / send_dpos = destination rank and node position
        there of every send entry (int32 device tensors); NCCL halo: the boundary / interior row lists."""
        nbr = np.ascontiguousarray(nbr, dtype=np.int32)
        send_ptr = np.ascontiguousarray(send_ptr, dtype=np.int32)
        recv_ptr = np.ascontiguousarray(recv_ptr, dtype=np.int32)
        self._keep = [send_idx, boundary_rows, interior_rows, send_dst, send_dpos]
        p = _lib.ptr
        _lib.check(self.lib.tefem_solver_set_halo(
            self.handle, self.n_halo, len(nbr), p(nbr), p(send_ptr), p(send_idx), p(send_dst), p(send_dpos), p(recv_ptr),
            p(boundary_rows), int(boundary_rows.numel()) if boundary_rows is not None else 0,
            p(interior_rows), int(interior_rows.numel()) if interior_rows is not None else 0))

    def set_interior_range(self, lo, hi, fused=True):
        """Rows [lo, hi) have no halo column; fused: the SpMV kernel does its own halo exchange (include/tefem.h)."""
        _lib.check(self.lib.tefem_solver_set_interior_range(self.handle, int(lo), int(hi), int(bool(fused))))

    @staticmethod
    def _info(info, method):
        return dict(iterations=int(info[0]), restarts=int(info[1]), relres=float(info[2]), bnorm=float(info[3]),
                    floor=bool(info[4]), stalls=int(info[5]), method=method)

    def bicgstab(self, sys, b, x, rtol=1e-10, max_iter=20000, check_every=16, floor_factor=1.0):
        """Solves sys x = b (both already block-Jacobi scaled); x holds the initial guess and the result.
        Returns dict(iterations, restarts, relres, bnorm, floor, stalls).  Raises ConvergenceError when rtol is not met;
        floor_factor > 1 accepts a solve whose TRUE residual stalled (the eps * cond floor of fp64) within
        floor_factor * rtol and reports it as info['floor'] = True (default 1: never accept above rtol)."""
        S = self.dmesh.schedule
        info = (ctypes.c_double * 6)()
        p = _lib.ptr
        rc = self.lib.tefem_bicgstab(self.handle, p(S.rowptr), p(S.colidx), p(sys), p(b), p(x), float(rtol),
                                     int(max_iter), int(check_every), float(floor_factor), info, _lib.stream_ptr())
        self.last_info = self._info(info, "bicgstab")
        _lib.check(rc)
        return self.last_info

    def cg(self, sys, dinv, b, x, rtol=1e-10, max_iter=20000, check_every=16, floor_factor=1.0):
        """Block-Jacobi-preconditioned CG on the UNSCALED symmetric positive definite system `sys` (dinv from
        block_jacobi_scale(sys, scale=False)); same stopping rule, info and exceptions as bicgstab."""
        S = self.dmesh.schedule
        info = (ctypes.c_double * 6)()
        p = _lib.ptr
        rc = self.lib.tefem_cg(self.handle, p(S.rowptr), p(S.colidx), p(sys), p(dinv), p(b), p(x), float(rtol),
                               int(max_iter), int(check_every), float(floor_factor), info, _lib.stream_ptr())
        self.last_info = self._info(info, "cg")
        _lib.check(rc)
        return self.last_info

    def set_precond(self, pc):
        """Attach a TwoLevelPreconditioner (multilevel.py) as right preconditioner, or detach with None."""
        self._pc = pc
        _lib.check(self.lib.tefem_solver_set_precond(self.handle, pc.handle if pc is not None else None))

    def profile(self, enable=None):
        """SpMV device time accumulated by the driver: returns (ms, count); enable True/False resets."""
        ms, cnt = ctypes.c_double(), ctypes.c_int64()
        flag = -1 if enable is None else int(bool(enable))
        _lib.check(self.lib.tefem_solver_profile(self.handle, flag, ctypes.byref(ms), ctypes.byref(cnt)))
        return ms.value, cnt.value

    def close(self):
        if self.handle:
            self.lib.tefem_solver_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def newmark_predict(lib, x, v, a, dt, gamma, beta, dir_mask, w1, w2):
    p = _lib.ptr
    _lib.check(lib.tefem_newmark_predict(x.numel(), p(x), p(v), p(a), float(dt), float(gamma), float(beta),
                                         p(dir_mask), p(w1), p(w2), _lib.stream_ptr()))


def newmark_correct(lib, x, v, a, a_new, dt, gamma, beta, dir_mask):
    p = _lib.ptr
    _lib.check(lib.tefem_newmark_correct(x.numel(), p(x), p(v), p(a), p(a_new), float(dt), float(gamma), float(beta),
                                         p(dir_mask), _lib.stream_ptr()))


def load_axpy(lib, node, w, coef, F):
    """F[4 node + c] += w * coef[c]; coef: 4 host numbers (kernel arguments) or a device tensor of 4 doubles."""
    if hasattr(coef, "data_ptr"):
        _lib.check(lib.tefem_load_axpy_dev(node.numel(), _lib.ptr(node), _lib.ptr(w), _lib.ptr(coef), _lib.ptr(F), _lib.stream_ptr()))
        return
    c = (ctypes.c_double * 4)(*[float(v) for v in coef])
    _lib.check(lib.tefem_load_axpy(node.numel(), _lib.ptr(node), _lib.ptr(w), c, _lib.ptr(F), _lib.stream_ptr()))


class GuessExtrapolator:
    """Initial guess of the next step's solve from the previous accelerations: a_n, then 2 a_n - a_{n-1}, then
    3 a_n - 3 a_{n-1} + a_{n-2} (the load is smooth in time, so the extrapolation error is O(dt^3) and the Krylov
    solve starts several decades closer to its tolerance).  The reference has no counterpart (direct solve)."""

    def __init__(self, lib, like, order=2):
        self.lib, self.order = lib, order
        self.h1 = torch.zeros_like(like)
        self.h2 = torch.zeros_like(like)
        self.n_hist = 0                   # number of valid older states behind `a`

    def guess(self, a, out):
        p = _lib.ptr
        k = min(self.order, self.n_hist)
        if k <= 0:
            out.copy_(a)
        elif k == 1:
            _lib.check(self.lib.tefem_lincomb3(a.numel(), 2.0, p(a), -1.0, p(self.h1), 0.0, None, p(out), _lib.stream_ptr()))
        else:
            _lib.check(self.lib.tefem_lincomb3(a.numel(), 3.0, p(a), -3.0, p(self.h1), 1.0, p(self.h2), p(out), _lib.stream_ptr()))

    def push(self, a_old):
        """Call with a_n BEFORE the corrector overwrites it with a_{n+1}."""
        self.h1, self.h2 = self.h2, self.h1
        self.h1.copy_(a_old)
        self.n_hist = min(self.n_hist + 1, 2)
