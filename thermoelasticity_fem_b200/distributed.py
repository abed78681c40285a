"""Multi-GPU transient solve: row-partitioned assembly, halo exchange, packed all-reduces.

One process per GPU (torchrun); ``torch.distributed`` provides the rendezvous (it broadcasts the NCCL unique
id), the data path uses the library's own communicator (``tefem_comm_*``: NCCL over NVLink resolved with
dlopen): one grouped send/recv of interface node values before each SpMV - overlapped with the interior
rows - and two packed all-reduces (1 and 8 doubles) per BiCGSTAB iteration.  Assembly needs no
communication (``partition.py``).

The reference has no distributed path; this module is an extension next to the drop-in API, built from the
same kernels.  Constrained DOFs and boundary loads are evaluated from the GLOBAL boundary description (a
halo node can be constrained through a triangle that is not local), then restricted to the local nodes.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib, layout
from .device import DeviceMesh, GuessExtrapolator, KrylovSolver, load_axpy, newmark_correct, newmark_predict
from .materials import LinearThermoElastic
from .partition import LocalProblem, localize, slab_ranges

_MODIF = {'x': 0, 'y': 1, 'z': 2}


def _nccl_path():
    try:
        import nvidia.nccl
        cand = os.path.join(os.path.dirname(nvidia.nccl.__file__), "lib", "libnccl.so.2")
        if os.path.exists(cand):
            return cand
    except Exception:
        pass
    return ""


class Comm:
    """Library communicator over all ranks of the default torch.distributed group."""

    def __init__(self):
        import torch.distributed as dist
        self.lib = _lib.load()
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        path = _nccl_path().encode()
        buf = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            raw = ctypes.create_string_buffer(128)
            _lib.check(self.lib.tefem_comm_unique_id(path, raw))
            buf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        buf = buf.to(dev)
        dist.broadcast(buf, src=0)
        raw = ctypes.create_string_buffer(bytes(buf.cpu().numpy().tobytes()), 128)
        self.handle = ctypes.c_void_p()
        _lib.check(self.lib.tefem_comm_create(ctypes.byref(self.handle), path, raw, self.world, self.rank))

        self.peer_cap = -1
        self.slot_doubles = 0

    def setup_vectors(self, n_local_nodes):
        """Allocate and exchange the peer-mapped vector slots of the solver (include/tefem.h, tefem_comm_vec_alloc):
        once per communicator, BEFORE the KrylovSolver is created.  Every rank passes its own local node count (owned +
        halo); the slots are sized for the largest."""
        import torch.distributed as dist
        need = ((4 * int(n_local_nodes) + 15) // 16) * 16
        t = torch.tensor([need], dtype=torch.int64, device="cuda" if dist.get_backend() == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        need = int(t.item())
        if self.slot_doubles:
            if need > self.slot_doubles:
                raise RuntimeError("peer-memory vector slots already allocated with a smaller size")
            return
        raw = ctypes.create_string_buffer(64)
        _lib.check(self.lib.tefem_comm_vec_alloc(self.handle, int(self.lib.tefem_solver_peer_slots()), need, raw))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(raw.raw))
        _lib.check(self.lib.tefem_comm_vec_open(self.handle, ctypes.create_string_buffer(b"".join(handles), 64 * self.world)))
        dist.barrier()
        self.slot_doubles = need

    def set_timeout(self, seconds):
        _lib.check(self.lib.tefem_comm_set_timeout(self.handle, float(seconds)))

    def setup_peer(self, big_cap_doubles=0):
        """Allocate and exchange the peer-memory region of the in-kernel reductions (include/tefem.h,
        tefem_comm_shm_alloc): once per communicator, before the first solve."""
        import torch.distributed as dist
        if self.peer_cap >= 0:
            if big_cap_doubles > self.peer_cap:
                raise RuntimeError("peer-memory region already allocated with a smaller capacity")
            return
        raw = ctypes.create_string_buffer(64)
        _lib.check(self.lib.tefem_comm_shm_alloc(self.handle, int(big_cap_doubles), raw))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(raw.raw))
        _lib.check(self.lib.tefem_comm_shm_open(self.handle, ctypes.create_string_buffer(b"".join(handles), 64 * self.world)))
        dist.barrier()
        self.peer_cap = int(big_cap_doubles)

    def allreduce_sum_(self, t):
        _lib.check(self.lib.tefem_comm_allreduce_sum(self.handle, _lib.ptr(t), t.numel(), _lib.stream_ptr()))
        return t

    def close(self):
        if self.handle:
            self.lib.tefem_comm_destroy(self.handle)
            self.handle = ctypes.c_void_p()


def dirichlet_node_data(tri_groups_global, dict_dirichlet_U, dict_dirichlet_theta):
    """Global constrained DOFs with fill / lift values as sparse (dof, fill, lift) arrays
    (same rules as Model.create_free_dofs_lists, reference model.py:51-85, 304-339)."""
    fill, lift = {}, {}
    if dict_dirichlet_U is not None:
        for tag, lst in dict_dirichlet_U.items():
            uniq = np.unique(np.asarray(tri_groups_global[tag]).reshape(-1))
            last = {}
            for letter, val in lst:
                if letter not in _MODIF:
                    raise ValueError('Unrecognized DOF.')
                last[_MODIF[letter]] = val
            for letter, _ in lst:
                m = _MODIF[letter]
                for d in 4 * uniq + m:
                    lift[d] = lift.get(d, 0.0) + last[m]
            for m, val in last.items():
                for d in 4 * uniq + m:
                    fill[d] = val
    if dict_dirichlet_theta is not None:
        for tag, theta in dict_dirichlet_theta.items():
            uniq = np.unique(np.asarray(tri_groups_global[tag]).reshape(-1))
            for d in 4 * uniq + 3:
                lift[d] = lift.get(d, 0.0) + theta
                fill[d] = theta
    dofs = np.array(sorted(fill), dtype=np.int64)
    return dofs, np.array([fill[d] for d in dofs]), np.array([lift[d] for d in dofs])


def longest_interior_range(boundary_rows, n_own):
    """[lo, hi): the longest run of owned rows without a halo column (row blocks in node order: everything between the
    interface layers)."""
    b = np.unique(np.asarray(boundary_rows, dtype=np.int64))
    edges = np.concatenate([[-1], b, [n_own]])
    gaps = edges[1:] - edges[:-1] - 1
    k = int(np.argmax(gaps))
    return int(edges[k] + 1), int(edges[k] + 1 + gaps[k])


def halo_destinations(lp: LocalProblem):
    """Per send entry of `lp`: destination rank and node position in THAT rank's local vector (its n_own + the offset of
    this rank's segment in its halo tail).  Collective: every rank publishes (n_own, neighbours, halo offsets)."""
    import torch.distributed as dist
    mine = (int(lp.n_own), [int(q) for q in lp.nbr], [int(v) for v in lp.recv_ptr])
    table = [None] * lp.world
    dist.all_gather_object(table, mine)
    n_send = int(lp.send_ptr[-1])
    dst = np.zeros(n_send, dtype=np.int32)
    dpos = np.zeros(n_send, dtype=np.int32)
    for k, q in enumerate(lp.nbr):
        n_own_q, nbr_q, recv_ptr_q = table[int(q)]
        j = nbr_q.index(lp.rank)                         # my segment in q's halo tail
        s0, s1 = int(lp.send_ptr[k]), int(lp.send_ptr[k + 1])
        if recv_ptr_q[j + 1] - recv_ptr_q[j] != s1 - s0:
            raise RuntimeError(f"halo mismatch between ranks {lp.rank} and {q}")
        dst[s0:s1] = q
        dpos[s0:s1] = n_own_q + recv_ptr_q[j] + np.arange(s1 - s0)
    return dst, dpos


def surface_node_weights(coords_of, tri_global):
    """Lumped nodal weights area/3 with the reference's 'area' 0.5*|X12.X13| (model.py:153), summed per GLOBAL node
    in triangle order.  coords_of(ids) -> [n, 3]."""
    tri = np.asarray(tri_global, dtype=np.int64)
    X1, X2, X3 = coords_of(tri[:, 0]), coords_of(tri[:, 1]), coords_of(tri[:, 2])
    area = 0.5 * np.abs(np.einsum('ij,ij->i', X2 - X1, X3 - X1))
    uniq, inv = np.unique(tri.reshape(-1), return_inverse=True)
    w = np.bincount(inv, weights=np.repeat(area / 3, 3), minlength=uniq.shape[0])
    return uniq, w


class DistributedTransient:
    """Newmark transient solve of one material / surface-load configuration on a row-partitioned mesh.

    lp: LocalProblem; coords_local [n_local, 3]; mat_id_local [E_loc]; materials: list of LinearThermoElastic;
    dirichlet = (global dofs, fill, lift); loads = list of (global node ids, weights, coef_fn(i) -> 4 amplitudes)."""

    def __init__(self, lp: LocalProblem, coords_local, mat_id_local, materials, dirichlet, loads,
                 alpha_M, alpha_K, t_end, n_t, gamma=0.5, beta=0.25, rtol=1e-10, max_iter=50000, check_every=16,
                 comm=None, precond="multilevel", halo="peer", floor_factor=1.0, mat_table=None, precond_options=None):
        """halo: 'peer' (default) - the SpMV is ONE kernel that stores this rank's boundary values straight into the neighbours'
        vectors over NVLink, multiplies the interior rows, waits for the neighbours' flags and multiplies the boundary rows;
        'peer-push' - the same stores from a separate push launch, the SpMV waits before its first row; 'nccl' - packed
        ncclSend / ncclRecv overlapped with the interior rows (both measured alternatives)."""
        assert halo in ("peer", "peer-push", "nccl")
        self.halo = halo
        self.floor_factor = floor_factor
        self.precond_options = precond_options or {}
        self.lp, self.comm = lp, comm
        self.alpha_M, self.alpha_K = alpha_M, alpha_K
        self.gamma, self.beta = gamma, beta
        self.t_end, self.n_t = t_end, n_t
        self.dt = t_end / (n_t - 1)
        self.rtol, self.max_iter, self.check_every, self.precond_kind = rtol, max_iter, check_every, precond
        table = np.asarray(mat_table) if mat_table is not None else np.array([m.table_row() for m in materials])
        self.dm = DeviceMesh(coords_local, lp.conn, mat_id_local, table, n_rows=lp.n_own)
        self.initial_state = None        # optional (x0, v0): local [4 n_local] arrays, free DOFs only are used
        self.n_halo = lp.n_halo
        self.lib = self.dm.lib
        dev = self.dm.device
        n_loc = 4 * lp.n_local
        # constrained DOFs of the LOCAL nodes (owned and halo)
        dofs, fill, lift = dirichlet
        loc = lp.to_local(dofs // 4)
        keep = loc >= 0
        ldof = 4 * loc[keep] + dofs[keep] % 4
        mask = np.zeros(n_loc, dtype=np.uint8)
        mask[ldof] = 1
        g_fill = np.zeros(n_loc)
        g_fill[ldof] = fill[keep]
        g_lift = np.zeros(n_loc)
        g_lift[ldof] = lift[keep]
        self._mask_host = mask.astype(bool)
        self.dir_mask = torch.from_numpy(mask).to(dev)
        self.g_fill = torch.from_numpy(g_fill).to(dev)
        self.g_lift = torch.from_numpy(g_lift).to(dev)
        self.has_lift = bool(np.any(g_lift != 0.0))
        # loads restricted to the OWNED nodes
        self.loads = []
        for nodes_g, w, coef_fn in loads:
            l = lp.to_local(nodes_g)
            own = (l >= 0) & (l < lp.n_own)
            if own.any():
                self.loads.append((torch.from_numpy(l[own].astype(np.int32)).to(dev),
                                   torch.from_numpy(np.asarray(w)[own]).to(dev), coef_fn))
        self.solve_info = []

    def prepare(self, steady=False):
        """Assembly, Dirichlet elimination, block-Jacobi scaling, solver / preconditioner / halo set-up.
        steady=False: the Newmark operator M + gamma dt D + beta dt^2 K (solvers.py:159-161); True: K alone (solvers.py:26)."""
        lp, dm, dev = self.lp, self.dm, self.dm.device
        dt = self.dt
        d_layout = layout.n_planes(layout.plane_map_D(self.alpha_M, self.alpha_K))
        if steady:
            self.planes = dm.assemble("K", self.alpha_M, self.alpha_K)
            self.sys = dm.form_system(self.planes, 0.0, 0.0, 1.0, d_layout, self.dir_mask)
        else:
            self.planes = dm.assemble("MKD", self.alpha_M, self.alpha_K)
            self.sys = dm.form_system(self.planes, 1.0, self.gamma * dt, self.beta * dt ** 2, d_layout, self.dir_mask)
        self.dinv = dm.block_jacobi_scale(self.sys)
        if self.comm is not None and self.halo != "nccl":
            self.comm.setup_vectors(lp.n_local)          # before the solver: it keeps its SpMV inputs in the slots
        self.krylov = KrylovSolver(dm, lp.n_halo, self.comm)
        if self.comm is not None and self.precond_kind != "multilevel":
            self.comm.setup_peer(0)
        self.precond = None
        if self.precond_kind == "multilevel":
            from .multilevel import TwoLevelPreconditioner
            self.precond = TwoLevelPreconditioner(dm, self.sys, comm=self.comm, **self.precond_options)
            self.krylov.set_precond(self.precond)
        send_idx = torch.from_numpy(lp.send_idx.astype(np.int32)).to(dev)
        if self.comm is not None and self.halo != "nccl":
            dst, dpos = halo_destinations(lp)
            self.krylov.set_halo(lp.nbr, lp.send_ptr, send_idx, lp.recv_ptr,
                                 send_dst=torch.from_numpy(dst).to(dev), send_dpos=torch.from_numpy(dpos).to(dev))
            lo, hi = longest_interior_range(lp.boundary_rows, lp.n_own)
            self.krylov.set_interior_range(lo, hi, fused=self.halo == "peer")
        else:
            self.krylov.set_halo(lp.nbr, lp.send_ptr, send_idx, lp.recv_ptr,
                                 torch.from_numpy(lp.boundary_rows.astype(np.int32)).to(dev),
                                 torch.from_numpy(lp.interior_rows.astype(np.int32)).to(dev))
        if self.comm is not None:
            import torch.distributed as dist
            torch.cuda.synchronize()
            dist.barrier()                               # every rank is set up before the first in-kernel wait
        n_loc = 4 * lp.n_local
        z = lambda: torch.zeros(n_loc, dtype=torch.float64, device=dev)
        self._x, self._v, self._a, self._a_new, self._w1, self._w2, self._rhs, self._b = (z() for _ in range(8))
        if self.initial_state is not None:      # solvers.py:155-157: only the free DOFs are advanced
            free = ~self._mask_host
            self._x.copy_(torch.from_numpy(np.asarray(self.initial_state[0]) * free))
            self._v.copy_(torch.from_numpy(np.asarray(self.initial_state[1]) * free))
        self._mapD = layout.plane_map_D(self.alpha_M, self.alpha_K)
        self._mapK = layout.plane_map_K()
        self._guess = GuessExtrapolator(self.lib, self._x, order=0)
        self._lift = None
        if self.has_lift:
            self._lift = z()
            dm.planes_spmv(self.planes["K"], self._mapK, self.g_lift, self._lift)

    def n_load_terms(self):
        return len(self.loads)

    def load_coefficients(self, i):
        return np.array([coef_fn(i) for _, _, coef_fn in self.loads], dtype=np.float64).reshape(-1, 4)

    def step(self, i, coef_dev=None):
        """coef_dev (optional, device [n_terms, 4]): this step's load amplitudes already on the device (end-to-end path)."""
        dm, lib = self.dm, self.lib
        rhs = self._rhs
        rhs.zero_()
        for k, (node, w, coef_fn) in enumerate(self.loads):
            load_axpy(lib, node, w, coef_fn(i) if coef_dev is None else coef_dev[k], rhs)
        if self._lift is not None:
            rhs -= self._lift
        newmark_predict(lib, self._x, self._v, self._a, self.dt, self.gamma, self.beta, self.dir_mask, self._w1, self._w2)
        dm.planes_spmv(self.planes["D"], self._mapD, self._w1, rhs, alpha=-1.0, beta_y=1.0)
        dm.planes_spmv(self.planes["K"], self._mapK, self._w2, rhs, alpha=-1.0, beta_y=1.0)
        dm.block_jacobi_apply(self.dinv, rhs, self._b, self.dir_mask)
        self._guess.guess(self._a, self._a_new)
        info = self.krylov.bicgstab(self.sys, self._b, self._a_new, self.rtol, self.max_iter, self.check_every,
                                    self.floor_factor)
        self._guess.push(self._a)
        # x, v, a are advanced on the halo nodes as well (same arithmetic as on their owner): they need a_new there
        _lib.check(lib.tefem_solver_halo_exchange(self.krylov.handle, _lib.ptr(self._a_new), _lib.stream_ptr()))
        newmark_correct(lib, self._x, self._v, self._a, self._a_new, self.dt, self.gamma, self.beta, self.dir_mask)
        self.solve_info.append(info)
        return info

    def solve_steady(self, idx_t=None):
        """K_ff x_f = F_f - K_fd x_d on the partitioned mesh (reference solvers.py:19-49); call prepare(steady=True) first.
        Returns the solver info; the solution (prescribed values included) is owned_state()[0]."""
        dm, lib = self.dm, self.lib
        rhs = self._rhs
        rhs.zero_()
        for node, w, coef_fn in self.loads:
            load_axpy(lib, node, w, coef_fn(idx_t), rhs)
        if self._lift is not None:
            rhs -= self._lift
        dm.block_jacobi_apply(self.dinv, rhs, self._b, self.dir_mask)
        self._x.zero_()
        info = self.krylov.bicgstab(self.sys, self._b, self._x, self.rtol, self.max_iter, self.check_every, self.floor_factor)
        self._x.masked_fill_(self.dir_mask.bool(), 0.0)
        self.solve_info.append(info)
        return info

    def owned_state(self):
        n = 4 * self.lp.n_own
        return (self._x + self.g_fill)[:n], self._v[:n], self._a[:n]


class _MeshShim:
    """What bench.py needs from a mesh in the distributed case."""

    def __init__(self, dm, n_nodes_global):
        self._dm, self.n_nodes, self.n_dofs = dm, n_nodes_global, 4 * n_nodes_global

    def device_mesh(self):
        return self._dm


def build_distributed_case(cells, rank, world, mat_kwargs, dU, dT, surface_arr, alpha_M, alpha_K, t_end, n_t, rtol,
                           halo="peer", precond_options=None):
    """BASELINE config 4 on `world` GPUs: unit box of cells^3 Kuhn cells, x-slabs of node planes, every rank
    generates only the cells that touch its planes.  Returns (mesh shim, model shim, solver) for bench.py."""
    from .synthetic import kuhn_face_tris, kuhn_node_coords, kuhn_slab_tets
    n = cells
    plane = (n + 1) * (n + 1)
    bounds = slab_ranges(n + 1, plane, world)
    p0, p1 = int(bounds[rank] // plane), int(bounds[rank + 1] // plane)         # owned node planes [p0, p1)
    conn_g, elem_ids = kuhn_slab_tets(n, n, n, p0 - 1, p1)                        # cells touching those planes
    lp = localize(conn_g, elem_ids, bounds, rank)
    coords_of = lambda ids: kuhn_node_coords(ids, n, n, n, 1.0, 1.0, 1.0)
    coords = coords_of(lp.local_nodes)
    tri_g = {4: kuhn_face_tris(n, n, n, 0, False), 5: kuhn_face_tris(n, n, n, 0, True)}
    dirichlet = dirichlet_node_data(tri_g, dU, dT)
    nodes5, w5 = surface_node_weights(coords_of, tri_g[5])
    loads = [(nodes5, w5, lambda i: (surface_arr[0, i], surface_arr[1, i], surface_arr[2, i], 0.0))]
    comm = Comm() if world > 1 else None
    solver = DistributedTransient(lp, coords, np.zeros(lp.conn.shape[0], dtype=np.int32),
                                  [LinearThermoElastic(**mat_kwargs)], dirichlet, loads, alpha_M, alpha_K,
                                  t_end, n_t, rtol=rtol, comm=comm, halo=halo, precond_options=precond_options)
    solver._ss = type("S", (), {})()          # bench.py reads solver._ss.krylov after prepare()
    orig_prepare = solver.prepare

    def prepare():
        orig_prepare()
        solver._ss.krylov = solver.krylov
    solver.prepare = prepare
    model = type("M", (), {})()
    model.dirichlet_dofs_U = dirichlet[0][dirichlet[0] % 4 != 3]
    model.dirichlet_dofs_theta = dirichlet[0][dirichlet[0] % 4 == 3]
    return _MeshShim(solver.dm, (n + 1) ** 3), model, solver
