"""Host-side setup of the aggregation preconditioner (csrc/tefem_precond.cu), once per system matrix.

Builds, with torch integer / scatter ops on the device (not on the repeatable path):
  * a geometric aggregation of the nodes: cubes of edge factor1 * h on a grid anchored at the GLOBAL bounding box
    (every rank computes the same aggregate for a node from its coordinates; aggregates may straddle ranks), the
    aggregate centres and the offset of every node from its centre;
  * the Galerkin operator of level 1, A1 = P1^T A^ P1, of the block-Jacobi-scaled fine system for the RIGID-BODY
    prolongation P1 (fine node i of aggregate I: u_i = t_I + w_I x d_i, theta_i = theta_I; 8 coarse values per
    aggregate stored as two 4-DOF pseudo nodes (t, theta), (w, 0)).  Each rank sums the blocks of its owned rows; the
    partial operators are gathered and summed (level 1 is replicated on every rank); A1 is scaled by the inverse of
    its 8x8 diagonal blocks;
  * level 2: aggregates of level-1 aggregates with the same rigid-body transfer between the centres, the transfer
    operators as small block-CSR matrices and the dense inverse of the level-2 Galerkin operator.

The reference has no counterpart (it calls a direct solver, solvers.py:26,172).
"""
import ctypes

import numpy as np
import torch

from . import _lib


def _dist(distributed=True):
    """torch.distributed when this set-up is a collective over the ranks (`distributed`: the caller's solve has a
    communicator - a single-GPU model built inside a multi-rank job must NOT enter collectives)."""
    if not distributed:
        return None
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) else None


def _allreduce(t, op, distributed=True):
    dist = _dist(distributed)
    if dist is not None:
        dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
    return t


def _gather_objects(obj, distributed=True):
    dist = _dist(distributed)
    if dist is None:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def _gather_cat(t, distributed=True):
    """Concatenation over the ranks of a tensor whose first dimension differs per rank (padded all_gather)."""
    dist = _dist(distributed)
    if dist is None:
        return t
    world = dist.get_world_size()
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:s] for o, s in zip(out, sizes)])


def _grid_keys(xyz, lo, H, dims):
    ijk = torch.floor((xyz - lo) / H + 1e-9).to(torch.int64)
    ijk = torch.minimum(ijk.clamp(min=0), (dims - 1).to(ijk.dtype))
    return (ijk[:, 0] * dims[1] + ijk[:, 1]) * dims[2] + ijk[:, 2]


def build_aggregation(xyz, n_own, factor1, distributed=True):
    """Geometric aggregation of the nodes: cubes of edge factor1 * h on a grid anchored at the GLOBAL bounding box
    (h = mean node spacing).  xyz: [n_local, 3] coordinates (owned nodes first).  Returns a dict with the sorted
    global aggregate keys `uniq`, the grid `dims1`, the aggregate of every local node `agg`, the owned nodes grouped
    by aggregate (`agg_ptr`, `order`), the aggregate centres `centres` [n1, 3] (mean of the member nodes over all
    ranks) and the offsets `off` [n_local, 3] of the local nodes from their centre.  Collective over the default
    torch.distributed group when `distributed` (a row-partitioned solve); purely local otherwise."""
    dev = xyz.device
    f64 = dict(dtype=torch.float64, device=dev)
    own = xyz[:n_own]
    lo = _allreduce(own.min(dim=0).values.clone(), "MIN", distributed)
    hi = _allreduce(own.max(dim=0).values.clone(), "MAX", distributed)
    n_glob = int(_allreduce(torch.tensor([float(n_own)], **f64), "SUM", distributed).item())
    ext = (hi - lo).clamp(min=1e-300)
    h = float(ext.prod().item() / max(n_glob, 1)) ** (1.0 / 3.0)
    H1 = factor1 * h
    dims1 = torch.ceil(ext / H1 + 1e-9).to(torch.int64).clamp(min=1)
    key = _grid_keys(xyz, lo, H1, dims1)                                   # all local nodes (owned + halo)
    uniq_local = torch.unique(key[:n_own]).cpu().numpy()
    uniq = np.unique(np.concatenate(_gather_objects(uniq_local, distributed)))          # same on every rank
    n1 = int(uniq.shape[0])
    agg = torch.searchsorted(torch.from_numpy(uniq).to(dev), key)          # aggregate of every local node
    node_agg = agg[:n_own]
    order = torch.sort(node_agg, stable=True)[1]
    agg_ptr = torch.zeros(n1 + 1, dtype=torch.int64, device=dev)
    torch.cumsum(torch.bincount(node_agg, minlength=n1), 0, out=agg_ptr[1:])
    sums = torch.zeros((n1, 4), **f64)
    sums.index_add_(0, node_agg, torch.cat([own, torch.ones((n_own, 1), **f64)], dim=1))
    _allreduce(sums, "SUM", distributed)
    centres = sums[:, :3] / sums[:, 3:4]
    off = xyz - centres[agg]
    return dict(h=h, H1=H1, dims1=dims1, uniq=uniq, n1=n1, agg=agg, node_agg=node_agg, order=order, agg_ptr=agg_ptr,
                centres=centres, off=off)


def rotation_blocks(d):
    """R(d) [m, 4, 4]: u = w x d for the rotation vector w, as the map (w_x, w_y, w_z, 0) -> (u_x, u_y, u_z, theta = 0)."""
    R = torch.zeros((d.shape[0], 4, 4), dtype=d.dtype, device=d.device)
    R[:, 0, 1], R[:, 0, 2] = d[:, 2], -d[:, 1]
    R[:, 1, 0], R[:, 1, 2] = -d[:, 2], d[:, 0]
    R[:, 2, 0], R[:, 2, 1] = d[:, 1], -d[:, 0]
    return R


def galerkin_level1(agg, off, block_row, colidx, sys, n1, chunk=1 << 22, distributed=True):
    """A1 = P1^T A P1 for the rigid-body prolongation P1 = [I4 | R(d_i)] of every node: the 4x4 blocks A_ij of the
    fine rows of this rank contribute the 8x8 block [[A, A R_j], [R_i^T A, R_i^T A R_j]] to the aggregate pair
    (agg i, agg j); the partial operators of all ranks are gathered and summed (level 1 is replicated).
    Returns (row1, col1, blocks [n_pairs, 8, 8] fp64): the aggregate pairs in (row, column) order."""
    dev = sys.device
    f64 = dict(dtype=torch.float64, device=dev)
    br, ci = block_row.long(), colidx.long()
    k2 = agg[br] * n1 + agg[ci]
    uk, inv = torch.unique(k2, return_inverse=True)
    del k2
    part = torch.zeros((uk.numel(), 8, 8), **f64)
    P = int(sys.shape[0])
    for s in range(0, P, chunk):
        e = min(P, s + chunk)
        A = sys[s:e].reshape(-1, 4, 4)
        Ri, Rj = rotation_blocks(off[br[s:e]]), rotation_blocks(off[ci[s:e]])
        B = torch.empty((e - s, 8, 8), **f64)
        B[:, :4, :4] = A
        B[:, :4, 4:] = torch.bmm(A, Rj)
        RtA = torch.bmm(Ri.transpose(1, 2), A)
        B[:, 4:, :4] = RtA
        B[:, 4:, 4:] = torch.bmm(RtA, Rj)
        part.index_add_(0, inv[s:e], B)
        del A, Ri, Rj, B, RtA
    del inv
    if _dist(distributed) is not None:
        keys, vals = _gather_cat(uk, distributed), _gather_cat(part, distributed)
        uk, inv = torch.unique(keys, return_inverse=True)
        part = torch.zeros((uk.numel(), 8, 8), **f64).index_add_(0, inv, vals)
        del keys, vals, inv
    row1 = torch.div(uk, n1, rounding_mode='floor')
    col1 = uk - row1 * n1
    return row1, col1, part


def scale_level1(row1, col1, blocks, n1, eps=1e-8):
    """Block-Jacobi scaling of level 1 with its 8x8 diagonal blocks: returns (rowptr1, D1^-1 [n1, 8, 8], scaled blocks).
    The rotation entries of the diagonal get a relative shift eps (aggregates whose nodes are collinear or single have
    rotation modes without energy; their restricted residual vanishes identically, the shift only keeps D1
    invertible) and the padding DOF a unit diagonal."""
    dev = blocks.device
    rowptr1 = torch.zeros(n1 + 1, dtype=torch.int64, device=dev)
    torch.cumsum(torch.bincount(row1, minlength=n1), 0, out=rowptr1[1:])
    diag1 = torch.nonzero(row1 == col1).reshape(-1)
    assert int(diag1.numel()) == n1, "every aggregate must have a diagonal block"
    D = blocks[diag1].clone()
    scale = D[:, :3, :3].diagonal(dim1=1, dim2=2).abs().mean(dim=1)
    for k in (4, 5, 6):
        D[:, k, k] += eps * scale
    D[:, 7, 7] = 1.0
    Dinv = torch.linalg.inv(D)
    scaled = blocks.clone()
    scaled[diag1] = D                                   # the shifted diagonal is what gets inverted
    scaled = torch.bmm(Dinv[row1], scaled)
    return rowptr1, Dinv, scaled


def pseudo_node_csr(rowptr1, col1, blocks8):
    """4x4-block CSR over the 2 n1 pseudo nodes of an operator given as 8x8 blocks per aggregate pair (pairs sorted by
    (row, column)).  Returns (rowptr [2 n1 + 1], colidx [4 n_pairs], vals [4 n_pairs, 16])."""
    dev = blocks8.device
    n1 = int(rowptr1.numel()) - 1
    cnt = rowptr1[1:] - rowptr1[:-1]
    rp = torch.zeros(2 * n1 + 1, dtype=torch.int64, device=dev)
    torch.cumsum(torch.repeat_interleave(2 * cnt, 2), 0, out=rp[1:])
    npairs = int(col1.numel())
    row_of = torch.repeat_interleave(torch.arange(n1, dtype=torch.int64, device=dev), cnt)
    k_in_row = torch.arange(npairs, dtype=torch.int64, device=dev) - rowptr1[row_of]
    colidx = torch.empty(4 * npairs, dtype=torch.int64, device=dev)
    vals = torch.empty((4 * npairs, 16), dtype=blocks8.dtype, device=dev)
    for a in range(2):
        for b in range(2):
            pos = 4 * rowptr1[row_of] + a * 2 * cnt[row_of] + 2 * k_in_row + b
            colidx[pos] = 2 * col1 + b
            vals[pos] = blocks8[:, 4 * a:4 * a + 4, 4 * b:4 * b + 4].reshape(-1, 16)
    return rp, colidx, vals


def build_level2(uniq, dims1, factor2, dense_limit=200):
    """Aggregation of the level-1 aggregates (grid cells `uniq` of the dims1 grid) into cubes of factor2 cells;
    identity when level 1 is small enough for the dense solve.  factor2 <= 0: chosen so that level 2 has at most a few
    hundred aggregates."""
    n1 = int(uniq.shape[0])
    if n1 <= dense_limit:
        return np.arange(n1, dtype=np.int64), n1, 1.0
    d1, d2_ = int(dims1[1]), int(dims1[2])
    ijk = np.stack([uniq // (d1 * d2_), (uniq // d2_) % d1, uniq % d2_], axis=1)
    if factor2 <= 0:
        factor2 = max(3.0, float(np.ceil((n1 / 250.0) ** (1.0 / 3.0))))
    c = np.floor(ijk / factor2 + 1e-9).astype(np.int64)
    d2 = c.max(axis=0) + 1
    key2 = (c[:, 0] * d2[1] + c[:, 1]) * d2[2] + c[:, 2]
    _, agg2 = np.unique(key2, return_inverse=True)
    return agg2.astype(np.int64), int(agg2.max()) + 1, factor2


def transfer_level2(centres1, agg2, n2):
    """P2 [8 n1, 8 n2] (scipy CSR): rigid-body transfer between the aggregate centres, t_I = T + W x (c_I - C),
    theta_I = Theta, w_I = W."""
    import scipy.sparse as sp
    n1 = centres1.shape[0]
    cnt = np.bincount(agg2, minlength=n2)
    C = np.stack([np.bincount(agg2, weights=centres1[:, k], minlength=n2) / cnt for k in range(3)], axis=1)
    d = centres1 - C[agg2]
    ar = np.arange(n1)
    rows, cols, vals = [], [], []
    for c in range(4):                       # translations and theta
        rows.append(8 * ar + c); cols.append(8 * agg2 + c); vals.append(np.ones(n1))
    dx, dy, dz = d[:, 0], d[:, 1], d[:, 2]
    for (c, k, v) in [(0, 1, dz), (0, 2, -dy), (1, 2, dx), (1, 0, -dz), (2, 0, dy), (2, 1, -dx)]:
        rows.append(8 * ar + c); cols.append(8 * agg2 + 4 + k); vals.append(v)
    for k in range(3):                       # rotations
        rows.append(8 * ar + 4 + k); cols.append(8 * agg2 + 4 + k); vals.append(np.ones(n1))
    return sp.csr_array((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(8 * n1, 8 * n2))


def _bsr_arrays(M, dev):
    """(rowptr int32, colidx int32, vals float32 [nnzb, 16]) of a scipy matrix as 4x4-block CSR, on the device."""
    B = M.tobsr(blocksize=(4, 4))
    B.sort_indices()
    return (torch.from_numpy(B.indptr.astype(np.int32)).to(dev), torch.from_numpy(B.indices.astype(np.int32)).to(dev),
            torch.from_numpy(np.ascontiguousarray(B.data.reshape(-1, 16)).astype(np.float32)).to(dev))


def row_owner(n, world, device=None):
    """Owner rank of each of n rows split as the library does: rank q owns [n q // world, n (q + 1) // world)."""
    bounds = torch.tensor([n * q // world for q in range(world + 1)], dtype=torch.int64, device=device)
    return torch.searchsorted(bounds, torch.arange(n, dtype=torch.int64, device=device), right=True) - 1


def cycle_masks(rowptr1, colidx1, n1, agg2, n2, agg_ranks, world):
    """Destination masks of the row-partitioned coarse cycle (include/tefem.h, tefem_precond_set_cycle_masks), uint8 over the
    aggregates, bit q = rank q reads the row: the level-1 iterates of aggregate i are read by its owner, by the owners of the
    rows that have a block in column i, and by the ranks that own fine nodes of i (prolongation); its residual t by its owner
    and by the owner of its level-2 aggregate.  The same on every rank (built from replicated data)."""
    dev = rowptr1.device
    own1 = row_owner(n1, world, dev)
    cnt = (rowptr1[1:] - rowptr1[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(2 * n1, dtype=torch.int64, device=dev), cnt)
    reader = own1[rows >> 1]                                   # rank that multiplies this block
    cagg = colidx1.long() >> 1
    x_mask = (torch.ones(n1, dtype=torch.int32, device=dev) << own1.to(torch.int32))
    for q in range(world):
        hit = torch.zeros(n1, dtype=torch.bool, device=dev)
        hit[cagg[reader == q]] = True
        x_mask |= hit.to(torch.int32) << q
    if agg_ranks is not None:
        x_mask |= agg_ranks.to(torch.int32)
    own2 = row_owner(n2, world, dev)
    a2 = torch.as_tensor(np.asarray(agg2), dtype=torch.int64, device=dev)
    t_mask = (torch.ones(n1, dtype=torch.int32, device=dev) << own1.to(torch.int32)) | \
             (torch.ones(n1, dtype=torch.int32, device=dev) << own2[a2].to(torch.int32))
    return x_mask.to(torch.uint8).contiguous(), t_mask.to(torch.uint8).contiguous()


class TwoLevelPreconditioner:
    """Owns the device arrays of one preconditioner and its C handle (attach with KrylovSolver.set_precond)."""

    def __init__(self, dm, sys, comm=None, factor1=8.0, factor2=3.0, omega=1.0, omega_c=0.7, n_pre=3, n_post=3,
                 dense_limit=200, level1_dtype="fp16", cycle="overlapped", coarse="auto"):
        """level1_dtype: storage of the scaled level-1 operator for the sweeps - 'fp16' (default; falls back to 'fp32' when
        an entry would overflow half precision) or 'fp32'.  Preconditioner data only.
        cycle: 'overlapped' (default) - the last pre-smoothing sweep and the level-2 correction use one residual (one
        level-1 product less per application) - or 'classic'.
        coarse: how the level-1 / level-2 cycle runs - 'launches' (separate kernels on levels replicated on every rank),
        'persistent' (one cooperative kernel per application; rows of levels 1 and 2 partitioned over the ranks, coarse
        vectors exchanged through peer memory) or 'auto' (persistent with a communicator, launches on one GPU)."""
        if coarse not in ("auto", "launches", "persistent"):
            raise ValueError("coarse must be 'auto', 'launches' or 'persistent'")
        self.coarse = ("persistent" if comm is not None else "launches") if coarse == "auto" else coarse
        self._coarse_auto = coarse == "auto"
        if cycle not in ("overlapped", "classic"):
            raise ValueError("cycle must be 'overlapped' or 'classic'")
        self.cycle = cycle
        import scipy.sparse as sp
        lib, dev = dm.lib, dm.device
        S = dm.schedule
        n_own = S.n_rows
        f64 = dict(dtype=torch.float64, device=dev)
        i32 = torch.int32
        p = _lib.ptr
        distributed = comm is not None
        A = build_aggregation(dm.coords, n_own, factor1, distributed=distributed)
        n1, uniq, dims1, h, H1 = A["n1"], A["uniq"], A["dims1"], A["h"], A["H1"]
        node_agg, order, agg_ptr = A["node_agg"], A["order"], A["agg_ptr"]
        row1, col1, blocks = galerkin_level1(A["agg"], A["off"], S.block_row, S.colidx, sys, n1, distributed=distributed)
        rowptr1, Dinv, scaled = scale_level1(row1, col1, blocks, n1)
        del blocks
        rp, ci, vals = pseudo_node_csr(rowptr1, col1, scaled)
        self.rowptr1, self.colidx1 = rp.to(i32), ci.to(i32)
        self.sys1 = vals.to(torch.float32).contiguous()
        if level1_dtype not in ("fp16", "fp32"):
            raise ValueError("level1_dtype must be 'fp16' or 'fp32'")
        self.sys1_half = None
        if level1_dtype == "fp16" and float(self.sys1.abs().max().item()) < 6.0e4:
            self.sys1_half = self.sys1.to(torch.float16).contiguous()
            # level 2 is built from (and consistent with) the operator the sweeps really use
            self.sys1 = self.sys1_half.to(torch.float32)
        self.level1_dtype = "fp16" if self.sys1_half is not None else "fp32"
        self.dinv1 = Dinv.reshape(n1, 64).contiguous()
        self.node_off = torch.zeros((n_own, 4), dtype=torch.float32, device=dev)
        self.node_off[:, :3] = A["off"][:n_own].to(torch.float32)
        # ---- level 2 (host, a few hundred aggregates): transfer operators and dense inverse of P2^T A1^ P2
        agg2, n2, factor2 = build_level2(uniq, dims1, factor2, dense_limit)
        if self._coarse_auto and self.coarse == "persistent" and 64 * n2 > 96 * 1024:
            self.coarse = "launches"      # the persistent kernel stages P2^T t (8 n2 doubles) in shared memory
        self.factor2 = factor2
        P2 = transfer_level2(A["centres"].cpu().numpy(), agg2, n2)
        A1 = sp.bsr_array((self.sys1.double().cpu().numpy().reshape(-1, 4, 4), self.colidx1.cpu().numpy(),
                           self.rowptr1.cpu().numpy()), shape=(8 * n1, 8 * n1)).tocsr()
        A2 = (P2.T @ A1 @ P2).toarray()
        dead = np.flatnonzero(np.abs(A2).sum(axis=1) == 0.0)          # padding DOFs
        A2[dead, dead] = 1.0
        rot = np.arange(8 * n2).reshape(-1, 8)[:, 4:7].ravel()
        A2[rot, rot] += 1e-8 * np.abs(np.diag(A2)).mean()
        self.inv2 = torch.from_numpy(np.ascontiguousarray(np.linalg.inv(A2)).astype(np.float32)).to(dev)
        self.r2_rowptr, self.r2_colidx, self.r2_vals = _bsr_arrays(P2.T.tocsr(), dev)
        self.p2_rowptr, self.p2_colidx, self.p2_vals = _bsr_arrays(P2, dev)
        assert int(self.r2_rowptr.numel()) == 2 * n2 + 1 and int(self.p2_rowptr.numel()) == 2 * n1 + 1
        self.agg_ptr, self.agg_nodes, self.node_agg = agg_ptr.to(i32), order.to(i32), node_agg.to(i32)
        self.agg2 = agg2
        self.n1, self.n2, self.h, self.H1 = n1, n2, h, H1
        self.factor1, self.omega, self.omega_c, self.n_pre, self.n_post = factor1, omega, omega_c, n_pre, n_post
        self.agg_ranks = None
        if comm is not None:
            # the restricted residual is all-reduced through peer memory; the row-partitioned cycle also keeps its coarse
            # vectors there (one replica per rank)
            comm.setup_peer(int(lib.tefem_precond_peer_doubles(n1, n2, int(n_pre), int(n_post))) if self.coarse == "persistent" else 8 * n1)
            # which ranks own fine nodes of an aggregate: only their partial sums are read over NVLink
            dist = _dist(True)
            mine = ((agg_ptr[1:] - agg_ptr[:-1]) > 0).to(torch.int32)
            if dist is not None:
                parts = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
                dist.all_gather(parts, mine)
            else:
                parts = [mine]
            bits = torch.zeros(n1, dtype=torch.int32, device=dev)
            for r, m in enumerate(parts):
                bits |= m << r
            self.agg_ranks = bits.to(torch.uint8).contiguous()
        self.work = torch.zeros(int(lib.tefem_precond_work_doubles(n1, n2)), **f64)
        self.handle = ctypes.c_void_p()
        self.lib = lib
        _lib.check(lib.tefem_precond_create(
            ctypes.byref(self.handle), n_own, n1, p(self.agg_ptr), p(self.agg_nodes), p(self.node_agg), p(self.node_off),
            float(omega), p(self.rowptr1), p(self.colidx1), p(self.sys1), p(self.dinv1), float(omega_c), int(n_pre),
            int(n_post), n2,
            p(self.r2_rowptr), p(self.r2_colidx), p(self.r2_vals), p(self.p2_rowptr), p(self.p2_colidx), p(self.p2_vals),
            p(self.inv2), p(self.work), comm.handle if comm is not None else None))
        if self.agg_ranks is not None:
            _lib.check(lib.tefem_precond_set_agg_ranks(self.handle, p(self.agg_ranks)))
            # the aggregates with fine nodes on this rank: the restriction runs one CTA per listed aggregate
            self.agg_list = torch.nonzero((self.agg_ptr[1:] - self.agg_ptr[:-1]) > 0).reshape(-1).to(i32).contiguous()
            if 0 < int(self.agg_list.numel()) < n1:
                _lib.check(lib.tefem_precond_set_agg_list(self.handle, p(self.agg_list), int(self.agg_list.numel())))
        if self.sys1_half is not None:
            _lib.check(lib.tefem_precond_set_level1_half(self.handle, p(self.sys1_half)))
        self.cycle = cycle
        _lib.check(lib.tefem_precond_set_cycle(self.handle, int(cycle == "overlapped")))
        self.cycle_work = None
        if self.coarse == "persistent":       # tagged coarse vectors of the persistent cycle (zero = no tag yet)
            self.cycle_work = torch.zeros(int(lib.tefem_precond_cycle_doubles(n1, n2, int(n_pre), int(n_post))), **f64)
        _lib.check(lib.tefem_precond_set_coarse_mode(self.handle, int(self.coarse == "persistent"),
                                                     p(self.cycle_work) if self.cycle_work is not None else None))
        self.x_mask = self.t_mask = None
        if self.coarse == "persistent" and comm is not None:
            self.x_mask, self.t_mask = cycle_masks(self.rowptr1, self.colidx1, n1, agg2, n2, self.agg_ranks, comm.world)
            _lib.check(lib.tefem_precond_set_cycle_masks(self.handle, p(self.x_mask), p(self.t_mask)))

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
