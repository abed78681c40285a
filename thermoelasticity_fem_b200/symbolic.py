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
  rowptr[N+1], colidx[P]            node-block CSR, columns ascending inside a row
  diag_idx[N]                       position of block (i, i)
  inc_ptr[P+1], inc[16E]            per block: contributing (local element, a, b), ascending element
  cta_item_ptr, item_block          blocks owned by each CTA, longest incidence lists first
  cta_elem_ptr, cta_elems           elements each CTA stages in shared memory
"""
from dataclasses import dataclass

import torch

ITEMS_PER_CTA = 256          # = ASM_THREADS of tefem_assemble.cu
ELEM_BUDGET_PER_CTA = 640    # bound on sum of node degrees per CTA (>= staged elements)
MAX_CTA_ELEMS = 4096         # 12 bits of the packed incidence entry


@dataclass
class Schedule:
    n_nodes: int             # number of local nodes (columns)
    n_rows: int              # number of owned node rows
    n_blocks: int
    n_elems: int
    rowptr: torch.Tensor     # int32 [n_rows+1]
    colidx: torch.Tensor     # int32 [P]
    block_row: torch.Tensor  # int32 [P]
    diag_idx: torch.Tensor   # int32 [n_rows]
    inc_ptr: torch.Tensor    # int32 [P+1]  (read as uint32 by the kernels)
    inc: torch.Tensor        # int16 [n_inc + 8] (read as uint16 by the kernels; 8 entries of padding)
    n_cta: int
    max_cta_elems: int
    max_cta_inc: int
    cta_item_ptr: torch.Tensor   # int32 [n_cta+1]
    item_block: torch.Tensor     # int32 [P]
    cta_elem_ptr: torch.Tensor   # int32 [n_cta+1]
    cta_elems: torch.Tensor      # int32 [sum]

    def to(self, device):
        kw = {}
        for k, v in self.__dict__.items():
            kw[k] = v.to(device) if isinstance(v, torch.Tensor) else v
        return Schedule(**kw)

    def schedule_bytes(self):
        """Bytes of slot-map / schedule data one assembly launch reads."""
        return sum(t.numel() * t.element_size() for t in
                   (self.inc_ptr, self.inc, self.cta_item_ptr, self.item_block, self.cta_elem_ptr, self.cta_elems))


def build_schedule(conn: torch.Tensor, n_nodes: int, n_rows: int = None,
                   items_per_cta: int = ITEMS_PER_CTA, elem_budget: int = ELEM_BUDGET_PER_CTA) -> Schedule:
    """conn: integer tensor [E, 4] in the reference's element order; node numbers < n_nodes.
    n_rows: only rows of nodes < n_rows are assembled (row-partitioned multi-GPU: owned nodes
    are numbered first); default all."""
    assert conn.dim() == 2 and conn.shape[1] == 4
    dev = conn.device
    E = int(conn.shape[0])
    N = int(n_nodes)
    R = N if n_rows is None else int(n_rows)
    conn = conn.to(torch.int64)
    assert 16 * E < 2 ** 31, "incidence count must fit 31 bits per rank"

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
    del ukey
    inc_ptr = torch.zeros(P + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=inc_ptr[1:])
    blocks_per_row = torch.bincount(block_row, minlength=R)
    rowptr = torch.zeros(R + 1, dtype=torch.int64, device=dev)
    torch.cumsum(blocks_per_row, 0, out=rowptr[1:])
    is_diag = block_row == block_col
    diag_idx = torch.nonzero(is_diag).reshape(-1)
    assert int(diag_idx.numel()) == R, "every owned node must belong to at least one element"

    # ---- rows -> CTAs: cumulative cost, integer arithmetic (blocks vs. staged-element budget)
    degree = torch.bincount(conn.reshape(-1), minlength=N)[:R]
    cost = torch.maximum(blocks_per_row * elem_budget, degree * items_per_cta)
    unit = items_per_cta * elem_budget
    excl = torch.cumsum(cost, 0) - cost
    cta_of_row = torch.div(excl, unit, rounding_mode='floor')
    # compact CTA ids (a very heavy row can skip an id)
    _, cta_of_row = torch.unique_consecutive(cta_of_row, return_inverse=True)
    n_cta = int(cta_of_row[-1].item()) + 1 if R > 0 else 0

    # ---- per-CTA staged element lists: unique (cta, element) over the (element, node) incidences
    node = conn.reshape(-1)
    elem_of = torch.arange(E, device=dev, dtype=torch.int64).repeat_interleave(4)
    owned = node < R
    k2 = cta_of_row[node[owned]] * E + elem_of[owned]
    del node, elem_of, owned
    k2 = torch.unique(k2)                             # sorted
    cta_of_k2 = torch.div(k2, E, rounding_mode='floor')
    cta_elems = k2 - cta_of_k2 * E
    cta_elem_ptr = torch.zeros(n_cta + 1, dtype=torch.int64, device=dev)
    torch.cumsum(torch.bincount(cta_of_k2, minlength=n_cta), 0, out=cta_elem_ptr[1:])
    max_cta_elems = int((cta_elem_ptr[1:] - cta_elem_ptr[:-1]).max().item()) if n_cta else 0
    assert max_cta_elems <= MAX_CTA_ELEMS, "CTA element list exceeds the 12-bit local index"
    del cta_of_k2

    # ---- packed incidence entries: (local element << 4) | (a << 2) | b
    cta_of_block = cta_of_row[block_row]
    e_src = torch.div(src, 16, rounding_mode='floor')
    ab = src - e_src * 16
    cta_of_inc = torch.repeat_interleave(cta_of_block, counts)
    loc = torch.searchsorted(k2, cta_of_inc * E + e_src) - cta_elem_ptr[cta_of_inc]
    del k2, e_src, cta_of_inc, src
    packed = (loc << 4) | ab
    del loc, ab
    packed = torch.where(packed >= 32768, packed - 65536, packed).to(torch.int16)
    # the kernel copies a CTA's incidence range with 16-byte loads from the aligned address below it:
    # keep the array readable 8 entries past its end
    packed = torch.cat([packed, torch.zeros(8, dtype=torch.int16, device=dev)])

    # ---- items of each CTA, longest lists first (homogeneous warps), stable inside equal lengths
    maxlen = int(counts.max().item()) if P else 0
    k3 = cta_of_block * (maxlen + 1) + (maxlen - counts)
    item_block = torch.sort(k3, stable=True)[1]
    cta_item_ptr = torch.zeros(n_cta + 1, dtype=torch.int64, device=dev)
    torch.cumsum(torch.bincount(cta_of_block, minlength=n_cta), 0, out=cta_item_ptr[1:])
    # incidence range of each CTA (its blocks are the contiguous range cta_item_ptr[c] .. cta_item_ptr[c+1])
    cta_inc = inc_ptr[cta_item_ptr[1:]] - inc_ptr[cta_item_ptr[:-1]]
    max_cta_inc = int(cta_inc.max().item()) if n_cta else 0

    i32 = torch.int32
    return Schedule(n_nodes=N, n_rows=R, n_blocks=P, n_elems=E,
                    rowptr=rowptr.to(i32), colidx=block_col.to(i32), block_row=block_row.to(i32),
                    diag_idx=diag_idx.to(i32), inc_ptr=inc_ptr.to(i32), inc=packed,
                    n_cta=n_cta, max_cta_elems=max_cta_elems, max_cta_inc=max_cta_inc,
                    cta_item_ptr=cta_item_ptr.to(i32), item_block=item_block.to(i32),
                    cta_elem_ptr=cta_elem_ptr.to(i32), cta_elems=cta_elems.to(i32))
