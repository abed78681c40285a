"""Row-partitioned multi-GPU solves behind the drop-in API: ``LinearTransient(model, ..., comm=...)`` and
``LinearSteadyState(model, comm=...)`` run the reference's algorithm (``solvers.py:19-66, 104-217``) on the GPUs of one node,
one process per GPU (torchrun), for ANY mesh the single-GPU path accepts - several materials, every load kind, non-zero
Dirichlet values.

The reference has no distributed path.  Every rank holds the (host-side) mesh description; the nodes are ordered by their
coordinate along the longest axis of the bounding box (a stable sort: structured boxes whose numbering is already sorted
keep it) and cut into contiguous ranges of equal size; a rank assembles every element that touches an owned node
(``partition.localize``: no communication in assembly, same summation order per row as one GPU).  Dirichlet data and loads
come from the Model's own host-side tables (``Model.create_free_dofs_lists``, ``Model.load_terms``), restricted to the local
nodes.  The solve is ``distributed.DistributedTransient`` (peer-memory halo, in-kernel all-reduces).
"""
import numpy as np
import torch

from .distributed import Comm, DistributedTransient
from .partition import localize, row_ranges


def coordinate_partition(points, world):
    """(order, inv, bounds): new node k is old node order[k] (sorted by the coordinate along the longest axis, stable);
    inv[old] = new; rank r owns the new nodes [bounds[r], bounds[r+1])."""
    pts = np.asarray(points)
    axis = int(np.argmax(pts.max(axis=0) - pts.min(axis=0)))
    order = np.argsort(pts[:, axis], kind="stable")
    inv = np.empty_like(order)
    inv[order] = np.arange(order.shape[0])
    return order, inv, row_ranges(pts.shape[0], world)


class PartitionedModel:
    """The local problem of this rank for a Model, and the maps back to the global (reference) numbering."""

    def __init__(self, model, rank, world):
        mesh = model.mesh
        conn, mat_id, table = mesh.element_tables()
        self.model, self.rank, self.world = model, rank, world
        self.order, self.inv, self.bounds = coordinate_partition(mesh.table_nodes, world)
        self.lp = localize(self.inv[conn], np.arange(conn.shape[0]), self.bounds, rank)
        self.coords_local = mesh.table_nodes[self.order[self.lp.local_nodes]]
        self.mat_id_local = mat_id[self.lp.elem_ids]
        self.mat_table = table
        if model._mask_host is None:
            model.create_free_dofs_lists()
        dofs_old = np.nonzero(model._mask_host)[0]
        dofs_new = 4 * self.inv[dofs_old // 4] + dofs_old % 4
        srt = np.argsort(dofs_new)
        self.dirichlet = (dofs_new[srt], model._g_fill_host[dofs_old][srt], model._g_lift_host[dofs_old][srt])

    def loads(self):
        """[(global nodes in the partition numbering, weights, coef_fn(i))] from Model.load_terms (model.py:120-286)."""
        model = self.model
        out = []
        for k, (node, w, _) in enumerate(model.load_terms(idx_t=None if not self._time_dependent() else 0)):
            nodes_new = self.inv[node.cpu().numpy().astype(np.int64)]
            out.append((nodes_new, w.cpu().numpy(), (lambda i, k=k: model.load_terms(idx_t=i)[k][2])))
        return out

    def _time_dependent(self):
        m = self.model
        for d in (m.dict_nodal_forces, m.dict_surface_forces, m.dict_volume_forces):
            if d is not None and any(np.asarray(v).ndim == 2 for v in d.values()):
                return True
        for d in (m.dict_heat_flux, m.dict_heat_source):
            if d is not None and any(getattr(v, "ndim", 0) == 1 for v in d.values()):
                return True
        return False

    def to_local(self, X_global):
        """[4N] vector in the reference numbering -> [4 n_local] local vector (owned + halo)."""
        nodes_old = self.order[self.lp.local_nodes]
        return np.asarray(X_global).reshape(-1, 4)[nodes_old].reshape(-1)

    def gather(self, owned, comm_device="cuda"):
        """Owned parts [4 n_own] (device tensors) of all ranks -> [4N] numpy vector in the reference numbering, on every rank."""
        import torch.distributed as dist
        sizes = [int(4 * (self.bounds[r + 1] - self.bounds[r])) for r in range(self.world)]
        pad = torch.zeros(max(sizes), dtype=torch.float64, device=owned.device)
        pad[:owned.numel()] = owned
        parts = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(parts, pad)
        new = torch.cat([p[:s] for p, s in zip(parts, sizes)]).cpu().numpy().reshape(-1, 4)
        out = np.empty_like(new)
        out[self.order] = new
        return out.reshape(-1)


def make_comm(comm):
    """comm: a distributed.Comm, or 'auto' / True: one over the default torch.distributed group (must be initialised with
    more than one rank)."""
    if isinstance(comm, Comm):
        return comm
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        raise RuntimeError("comm='auto' needs torch.distributed initialised with more than one rank (torchrun)")
    return Comm()


def distributed_transient(model, x0_full, v0_full, t_end, n_t, gamma, beta, comm, rtol, max_iter, check_every, precond,
                          floor_factor, min_nodes_multilevel, precond_options=None):
    """(PartitionedModel, DistributedTransient) for a Model; x0_full / v0_full: [4N] initial state, reference numbering."""
    pm = PartitionedModel(model, comm.rank, comm.world)
    if precond == "auto":
        precond = "multilevel" if model.mesh.n_nodes >= min_nodes_multilevel else "block_jacobi"
    solver = DistributedTransient(pm.lp, pm.coords_local, pm.mat_id_local, None, pm.dirichlet, pm.loads(), model.alpha_M,
                                  model.alpha_K, t_end, n_t, gamma, beta, rtol=rtol, max_iter=max_iter, check_every=check_every,
                                  comm=comm, precond=precond, floor_factor=floor_factor, mat_table=pm.mat_table,
                                  precond_options=precond_options)
    solver.initial_state = (pm.to_local(x0_full), pm.to_local(v0_full))
    return pm, solver
