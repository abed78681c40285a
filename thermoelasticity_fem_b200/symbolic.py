"""Symbolic phase: node-block sparsity pattern, slot map and owner-computes schedule.

Runs once per mesh (not on the repeatable path) with plain torch integer ops on whatever
device the connectivity lives on, so the same code is exercised on CPU tensors by the
``-m "not gpu"`` tests and on the GPU in production.

What it restates from the reference: the set of ``(row, col)`` pairs the assembly loops add
into through ``np.ix_`` (``model.py:91,98,100,102,109,111,114,117``) - a pure function of the
connectivity and of which blocks an operator has - and the element order of those additions
(``mesh.py:51-58``).  The node-pair ("block") pattern is shared by M, D and K; the scalar
touched pattern of each operator follows from the plane maps in ``layout.py``.

Outputs (see include/tefem.h, tefem_assemble):
  rowptr[R+1], colidx[P]     node-block CSR, columns ascending inside a row
  diag_idx[R]                position of block (i, i)
  inc[...]                   incidence entries (tile-local element << 4) | (a << 2) | b of the DIAGONAL and
                             UPPER (row < col) blocks, block after block, ascending element inside a block
  work items                 an upper block (i, j) - its thread also produces the mirrored block (j, i), whose
                             contributions come from the same elements with (a, b) swapped - or a chunk of
                             <= max_chunk incidences of a diagonal block: item_block, item_blockT (mirror block
                             or -1), item_inc (first entry), item_meta (length | (partial slot + 1) << 8)
  split_block, split_meta    diagonal blocks whose list was split: block, first slot | n_chunks << 16
  cta_ptr[n_cta, 12]         per tile: item, staged-element, incidence, split and node ranges
  cta_elems, cta_conn8       elements each tile stages, and their connectivity as 4 tile-local node bytes
  cta_nodes                  nodes whose coordinates each tile stages
"""
from dataclasses import dataclass

import torch

ITEMS_PER_CTA = 256          # = ASM_THREADS of tefem_assemble.cu
ELEM_BUDGET_PER_CTA = 600    # bound on the sum of node degrees per tile (>= staged elements)
MAX_CHUNK = 6                # longest incidence list handled by one thread
MAX_CTA_ELEMS = 4096         # 12 bits of the packed incidence entry
MAX_CTA_NODES = 256          # tile-local node numbers are bytes
CTA_STRIDE = 12


@dataclass
class Schedule:
    n_nodes: int             # number of local nodes (columns)
    n_rows: int              # number of owned node rows
    n_blocks: int
    n_elems: int
    n_items: int
    rowptr: torch.Tensor     # int32 [n_rows+1]
    colidx: torch.Tensor     # int32 [P]
    block_row: torch.Tensor  # int32 [P]
    diag_idx: torch.Tensor   # int32 [n_rows]
    inc_ptr: torch.Tensor    # int32 [n_listed+1]  offsets of the listed (diagonal + upper) blocks into inc
    listed_block: torch.Tensor   # int32 [n_listed] block number of each listed block
    inc: torch.Tensor        # int16 [n_inc + 8] (read as uint16 by the kernels; 8 entries of padding)
    n_cta: int
    max_cta_elems: int
    max_cta_inc: int
    max_cta_slots: int
    max_cta_nodes: int
    cta_ptr: torch.Tensor        # int32 [n_cta, 12]: item0, item1, elem0, elem1, inc0, inc1, split0, split1, node0, node1, 0, 0
    cta_elems: torch.Tensor      # int32 [n_staged]
    cta_conn8: torch.Tensor      # int32 [n_staged] 4 tile-local node numbers, byte a = node a
    cta_nodes: torch.Tensor      # int32 [n_tile_nodes]
    item_block: torch.Tensor     # int32 [n_items]
    item_blockT: torch.Tensor    # int32 [n_items]
    item_inc: torch.Tensor       # int32 [n_items]
    item_meta: torch.Tensor      # int32 [n_items]
    split_block: torch.Tensor    # int32 [n_split]
    split_meta: torch.Tensor     # int32 [n_split]

    def to(self, device):
        kw = {}
        for k, v in self.__dict__.items():
            kw[k] = v.to(device) if isinstance(v, torch.Tensor) else v
        return Schedule(**kw)

    def schedule_bytes(self):
        """Bytes of slot-map / schedule data one assembly launch reads (cta_mat is counted by DeviceMesh)."""
        return sum(t.numel() * t.element_size() for t in
                   (self.inc, self.cta_ptr, self.cta_conn8, self.cta_nodes, self.item_block, self.item_blockT,
                    self.item_inc, self.item_meta, self.split_block, self.split_meta))


def _excl_cumsum(x):
    return torch.cumsum(x, 0) - x


def _ptr_from_counts(counts, n, dev):
    ptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=ptr[1:])
    return ptr


class _TileTooLarge(Exception):
    pass


def build_schedule(conn: torch.Tensor, n_nodes: int, n_rows: int = None,
                   items_per_cta: int = ITEMS_PER_CTA, elem_budget: int = ELEM_BUDGET_PER_CTA,
                   max_chunk: int = MAX_CHUNK) -> Schedule:
    """conn: integer tensor [E, 4] in the reference's element order; node numbers < n_nodes.
    n_rows: only rows of nodes < n_rows are assembled (row-partitioned multi-GPU: owned nodes
    are numbered first); default all.

    A tile must not reference more than 256 nodes / 4096 elements (byte / 12-bit local numbers); meshes whose
    node numbering has little locality (e.g. the Gmsh numbering of data/sandwich.msh) get smaller tiles."""
    for _ in range(12):
        try:
            return _build_schedule(conn, n_nodes, n_rows, items_per_cta, elem_budget, max_chunk)
        except _TileTooLarge:
            items_per_cta = max(16, (3 * items_per_cta) // 4)
            elem_budget = max(32, (3 * elem_budget) // 4)
    raise RuntimeError("could not build a tile schedule within the shared-memory index limits")


def _build_schedule(conn, n_nodes, n_rows, items_per_cta, elem_budget, max_chunk):
    assert conn.dim() == 2 and conn.shape[1] == 4
    assert 1 <= max_chunk <= 255
    dev = conn.device
    E = int(conn.shape[0])
    N = int(n_nodes)
    R = N if n_rows is None else int(n_rows)
    conn = conn.to(torch.int64)
    assert 16 * E < 2 ** 31, "incidence count must fit 31 bits per rank"
    i64 = dict(dtype=torch.int64, device=dev)

    # ---- all (row, col) incidences in (element, a, b) order, keyed by row*N + col
    key = (conn[:, :, None] * N + conn[:, None, :]).reshape(-1)
    if R < N:
        flat = torch.nonzero(conn.repeat_interleave(4, dim=1).reshape(-1) < R).reshape(-1)   # row = conn[e, a]
        key = key[flat]
    else:
        flat = None
    key, order = torch.sort(key, stable=True)        # stable: ascending element inside a block
    src = order if flat is None else flat[order]     # original flat index e*16 + a*4 + b
    del order, flat
    ukey, counts = torch.unique_consecutive(key, return_counts=True)
    del key
    P = int(ukey.numel())
    block_row = torch.div(ukey, N, rounding_mode='floor')
    block_col = ukey - block_row * N
    blocks_per_row = torch.bincount(block_row, minlength=R)
    rowptr = _ptr_from_counts(blocks_per_row, R, dev)
    diag_idx = torch.nonzero(block_row == block_col).reshape(-1)
    assert int(diag_idx.numel()) == R, "every owned node must belong to at least one element"

    # ---- listed blocks: diagonal and upper; the mirror (j, i) of an upper block is produced by the same thread
    listed = block_col >= block_row
    listed_block = torch.nonzero(listed).reshape(-1)
    L = int(listed_block.numel())
    lcounts = counts[listed_block]
    is_upper = (block_col > block_row)[listed_block]
    mirror = torch.full((L,), -1, **i64)
    up = listed_block[is_upper]
    # the mirror block exists in this rank's pattern only when its row (= col of the upper block) is owned
    mkey = block_col[up] * N + block_row[up]
    pos = torch.searchsorted(ukey, mkey).clamp(max=max(P - 1, 0))
    found = ukey[pos] == mkey
    mirror[is_upper] = torch.where(found, pos, torch.full_like(pos, -1))
    assert bool((found | (block_col[up] >= R)).all()), "pattern must be structurally symmetric on the owned rows"
    del ukey, mkey, pos, found, up
    # incidence entries of the listed blocks only
    inc_keep = torch.repeat_interleave(listed, counts)
    src = src[inc_keep]
    del inc_keep
    inc_ptr = _ptr_from_counts(lcounts, L, dev)

    # ---- work items: upper blocks are single items, diagonal blocks are split into balanced chunks
    n_chunks = torch.where(is_upper, torch.ones_like(lcounts),
                           torch.div(lcounts + (max_chunk - 1), max_chunk, rounding_mode='floor'))
    lrow = block_row[listed_block]
    items_per_row = torch.zeros(R, **i64).index_add_(0, lrow, n_chunks)

    # ---- rows -> tiles: cumulative cost, integer arithmetic (work items vs. staged-element budget)
    degree = torch.bincount(conn.reshape(-1), minlength=N)[:R]
    cost = torch.maximum(items_per_row * elem_budget, degree * items_per_cta)
    unit = items_per_cta * elem_budget
    cta_of_row = torch.div(_excl_cumsum(cost), unit, rounding_mode='floor')
    _, cta_of_row = torch.unique_consecutive(cta_of_row, return_inverse=True)    # compact ids
    n_cta = int(cta_of_row[-1].item()) + 1 if R > 0 else 0
    cta_of_listed = cta_of_row[lrow]

    # ---- per-tile staged element lists: unique (tile, element) over the (element, node) incidences
    node = conn.reshape(-1)
    elem_of = torch.arange(E, **i64).repeat_interleave(4)
    owned = node < R
    k2 = torch.unique(cta_of_row[node[owned]] * E + elem_of[owned])             # sorted
    del node, elem_of, owned
    cta_of_k2 = torch.div(k2, E, rounding_mode='floor')
    cta_elems = k2 - cta_of_k2 * E
    cta_elem_ptr = _ptr_from_counts(torch.bincount(cta_of_k2, minlength=n_cta), n_cta, dev)
    max_cta_elems = int((cta_elem_ptr[1:] - cta_elem_ptr[:-1]).max().item()) if n_cta else 0
    if max_cta_elems > MAX_CTA_ELEMS:
        raise _TileTooLarge()

    # ---- per-tile node lists (coordinates staged once per tile) and byte-sized tile-local connectivity
    n_staged = int(cta_elems.numel())
    sconn = conn[cta_elems]                                                      # [n_staged, 4]
    k4 = cta_of_k2[:, None] * N + sconn
    unodes = torch.unique(k4.reshape(-1))                                        # sorted (tile, node)
    cta_of_un = torch.div(unodes, N, rounding_mode='floor')
    cta_nodes = unodes - cta_of_un * N
    cta_node_ptr = _ptr_from_counts(torch.bincount(cta_of_un, minlength=n_cta), n_cta, dev)
    max_cta_nodes = int((cta_node_ptr[1:] - cta_node_ptr[:-1]).max().item()) if n_cta else 0
    if max_cta_nodes > MAX_CTA_NODES:
        raise _TileTooLarge()
    lnode = torch.searchsorted(unodes, k4.reshape(-1)).reshape(n_staged, 4) - cta_node_ptr[cta_of_k2][:, None]
    conn8 = lnode[:, 0] | (lnode[:, 1] << 8) | (lnode[:, 2] << 16) | (lnode[:, 3] << 24)
    conn8 = torch.where(conn8 >= 2 ** 31, conn8 - 2 ** 32, conn8)
    del sconn, k4, unodes, cta_of_un, lnode, cta_of_k2

    # ---- packed incidence entries: (tile-local element << 4) | (a << 2) | b
    e_src = torch.div(src, 16, rounding_mode='floor')
    ab = src - e_src * 16
    cta_of_inc = torch.repeat_interleave(cta_of_listed, lcounts)
    loc = torch.searchsorted(k2, cta_of_inc * E + e_src) - cta_elem_ptr[cta_of_inc]
    del k2, e_src, cta_of_inc, src
    packed = (loc << 4) | ab
    del loc, ab
    packed = torch.where(packed >= 32768, packed - 65536, packed).to(torch.int16)
    # the kernel copies a tile's incidence range with 16-byte loads from the aligned address below it:
    # keep the array readable 8 entries past its end
    packed = torch.cat([packed, torch.zeros(8, dtype=torch.int16, device=dev)])

    # ---- item arrays in (listed block, chunk) order
    n_items = int(n_chunks.sum().item())
    item_l = torch.repeat_interleave(torch.arange(L, **i64), n_chunks)
    chunk = torch.arange(n_items, **i64) - _excl_cumsum(n_chunks)[item_l]
    base_len = torch.div(lcounts, n_chunks, rounding_mode='floor')[item_l]
    rem = (lcounts - torch.div(lcounts, n_chunks, rounding_mode='floor') * n_chunks)[item_l]
    item_len = base_len + (chunk < rem).to(torch.int64)
    item_inc = inc_ptr[item_l] + chunk * base_len + torch.minimum(chunk, rem)
    is_split = n_chunks[item_l] > 1
    cta_of_item = cta_of_listed[item_l]
    cta_item_ptr = _ptr_from_counts(torch.bincount(cta_of_item, minlength=n_cta), n_cta, dev)
    split_run = _excl_cumsum(is_split.to(torch.int64))                           # global rank among split items
    slot = split_run - split_run[cta_item_ptr[:-1].clamp(max=max(n_items - 1, 0))][cta_of_item]
    item_meta = item_len | torch.where(is_split, (slot + 1) << 8, torch.zeros_like(slot))
    item_block = listed_block[item_l]
    item_blockT = mirror[item_l]
    item_upper = is_upper[item_l]
    item_meta = item_meta | (item_upper.to(torch.int64) << 24)
    # split blocks of each tile
    l_split = torch.nonzero(n_chunks > 1).reshape(-1)
    first_item = _excl_cumsum(n_chunks)[l_split]
    split_meta = slot[first_item] | (n_chunks[l_split] << 16)
    split_block = listed_block[l_split]
    cta_split_ptr = _ptr_from_counts(torch.bincount(cta_of_listed[l_split], minlength=n_cta), n_cta, dev)
    slots_per_cta = torch.zeros(n_cta, **i64).index_add_(0, cta_of_item, is_split.to(torch.int64))
    max_cta_slots = int(slots_per_cta.max().item()) if n_cta else 0
    assert max_cta_slots < 65536

    # ---- order the items of each tile: upper blocks first (their threads do more work), then longest first
    # (homogeneous warps); stable inside equal keys
    kind = torch.where(item_upper, torch.zeros_like(item_len), torch.ones_like(item_len))
    perm = torch.sort((cta_of_item * 2 + kind) * (max(max_chunk, int(lcounts.max().item()) if L else 1) + 1)
                      + (max(max_chunk, int(lcounts.max().item()) if L else 1) - item_len), stable=True)[1]
    item_block, item_blockT, item_inc, item_meta = item_block[perm], item_blockT[perm], item_inc[perm], item_meta[perm]
    del perm

    # ---- per-tile ranges
    listed_per_cta = torch.bincount(cta_of_listed, minlength=n_cta)
    cta_l_ptr = _ptr_from_counts(listed_per_cta, n_cta, dev)
    cta_inc_ptr = inc_ptr[cta_l_ptr]
    max_cta_inc = int((cta_inc_ptr[1:] - cta_inc_ptr[:-1]).max().item()) if n_cta else 0
    zero = torch.zeros(n_cta, **i64)
    cta_ptr = torch.stack([cta_item_ptr[:-1], cta_item_ptr[1:], cta_elem_ptr[:-1], cta_elem_ptr[1:],
                           cta_inc_ptr[:-1], cta_inc_ptr[1:], cta_split_ptr[:-1], cta_split_ptr[1:],
                           cta_node_ptr[:-1], cta_node_ptr[1:], zero, zero], dim=1)

    i32 = torch.int32
    return Schedule(n_nodes=N, n_rows=R, n_blocks=P, n_elems=E, n_items=n_items,
                    rowptr=rowptr.to(i32), colidx=block_col.to(i32), block_row=block_row.to(i32),
                    diag_idx=diag_idx.to(i32), inc_ptr=inc_ptr.to(i32), listed_block=listed_block.to(i32), inc=packed,
                    n_cta=n_cta, max_cta_elems=max_cta_elems, max_cta_inc=max_cta_inc, max_cta_slots=max_cta_slots,
                    max_cta_nodes=max_cta_nodes,
                    cta_ptr=cta_ptr.to(i32).contiguous(), cta_elems=cta_elems.to(i32), cta_conn8=conn8.to(i32),
                    cta_nodes=cta_nodes.to(i32),
                    item_block=item_block.to(i32), item_blockT=item_blockT.to(i32), item_inc=item_inc.to(i32),
                    item_meta=item_meta.to(i32), split_block=split_block.to(i32), split_meta=split_meta.to(i32))
