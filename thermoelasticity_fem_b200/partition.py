"""Row-block (node-block) partition of the mesh for the multi-GPU path.

The reference has no distributed path.  Ranks own contiguous ranges of node rows (a node's 4 DOFs stay
together: dof = 4*node + c, reference ``elements.py:192-197``).  A rank assembles every element that touches
one of its nodes - interface elements are duplicated on the ranks that share them, so assembly needs NO
communication and every owned row receives all its contributions locally, in the same (global) element
order as a single-GPU assembly.  Columns that reference nodes of other ranks are renumbered into a halo
segment placed after the owned nodes, grouped by owner (ascending global id).

Pure host logic (numpy): exercised on CPU by the ``gloo`` world_size-2 tests.
"""
from dataclasses import dataclass

import numpy as np


def row_ranges(n_nodes, world):
    """Contiguous, as even as possible: rank r owns nodes [bounds[r], bounds[r+1])."""
    return np.array([(n_nodes * r) // world for r in range(world + 1)], dtype=np.int64)


def slab_ranges(n_planes, plane_size, world):
    """Structured boxes: cut between whole x-planes of nodes (node id = plane * plane_size + ...)."""
    cuts = np.array([(n_planes * r) // world for r in range(world + 1)], dtype=np.int64)
    return cuts * plane_size


@dataclass
class LocalProblem:
    rank: int
    world: int
    bounds: np.ndarray          # [world+1] global node ranges
    n_own: int
    n_halo: int
    local_nodes: np.ndarray     # [n_own + n_halo] global node id of every local node (owned first, then halo ascending)
    conn: np.ndarray            # [E_loc, 4] local node numbers, global element order
    elem_ids: np.ndarray        # [E_loc] global element numbers
    nbr: np.ndarray             # neighbour ranks (ascending)
    send_ptr: np.ndarray        # [n_nbr+1]
    send_idx: np.ndarray        # local (owned) node numbers to send, per neighbour, ascending global id
    recv_ptr: np.ndarray        # [n_nbr+1] offsets into the halo segment
    boundary_rows: np.ndarray   # owned rows with at least one halo column
    interior_rows: np.ndarray

    @property
    def n_local(self):
        return self.n_own + self.n_halo

    def to_local(self, global_nodes):
        """Global node ids -> local numbers (-1 when the node is not local)."""
        g = np.asarray(global_nodes, dtype=np.int64)
        lo, hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        out = np.full(g.shape, -1, dtype=np.int64)
        own = (g >= lo) & (g < hi)
        out[own] = g[own] - lo
        halo = self.local_nodes[self.n_own:]
        if halo.size:
            pos = np.searchsorted(halo, g[~own])
            pos_c = np.minimum(pos, halo.size - 1)
            ok = halo[pos_c] == g[~own]
            out[~own] = np.where(ok, self.n_own + pos_c, -1)
        return out


def localize(conn_global, elem_ids, bounds, rank):
    """conn_global: [E', 4] GLOBAL node ids of (at least) every element that touches a node owned by `rank`,
    in ascending global element order; elem_ids their global numbers."""
    world = len(bounds) - 1
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    conn_global = np.asarray(conn_global, dtype=np.int64)
    touches = ((conn_global >= lo) & (conn_global < hi)).any(axis=1)
    conn_g = conn_global[touches]
    elem_ids = np.asarray(elem_ids, dtype=np.int64)[touches]
    nodes = np.unique(conn_g)
    halo = nodes[(nodes < lo) | (nodes >= hi)]
    n_own, n_halo = hi - lo, int(halo.size)
    local_nodes = np.concatenate([np.arange(lo, hi, dtype=np.int64), halo])
    # local numbering of the connectivity
    conn_l = np.where((conn_g >= lo) & (conn_g < hi), conn_g - lo, n_own + np.searchsorted(halo, conn_g))
    # halo grouped by owner (ascending global id => ascending owner)
    owner_of_halo = np.searchsorted(bounds, halo, side="right") - 1
    nbr = np.unique(owner_of_halo)
    recv_ptr = np.concatenate([[0], np.cumsum([int((owner_of_halo == q).sum()) for q in nbr])]).astype(np.int64)
    # send lists: my nodes that share an element with a node owned by q
    owner = np.searchsorted(bounds, conn_g, side="right") - 1            # [E_loc, 4]
    send_lists = []
    nbr_send = []
    for q in range(world):
        if q == rank:
            continue
        has_q = (owner == q).any(axis=1)
        if not has_q.any():
            continue
        mine = conn_g[has_q]
        mine = np.unique(mine[(mine >= lo) & (mine < hi)])
        nbr_send.append(q)
        send_lists.append(mine - lo)
    assert list(nbr_send) == list(nbr), "send / receive neighbour sets must coincide"
    send_ptr = np.concatenate([[0], np.cumsum([len(s) for s in send_lists])]).astype(np.int64)
    send_idx = np.concatenate(send_lists).astype(np.int64) if send_lists else np.zeros(0, dtype=np.int64)
    # owned rows that reference a halo column: the rows of nodes in elements that contain a halo node
    has_halo = (conn_l >= n_own).any(axis=1)
    b = np.unique(conn_l[has_halo])
    boundary_rows = b[b < n_own]
    interior_mask = np.ones(n_own, dtype=bool)
    interior_mask[boundary_rows] = False
    interior_rows = np.nonzero(interior_mask)[0]
    return LocalProblem(rank=rank, world=world, bounds=np.asarray(bounds), n_own=n_own, n_halo=n_halo,
                        local_nodes=local_nodes, conn=conn_l, elem_ids=elem_ids, nbr=nbr.astype(np.int64),
                        send_ptr=send_ptr, send_idx=send_idx, recv_ptr=recv_ptr,
                        boundary_rows=boundary_rows.astype(np.int64), interior_rows=interior_rows.astype(np.int64))
