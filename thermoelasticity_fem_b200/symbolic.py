"""Symbolic phase: node-block sparsity pattern, slot map and owner-computes tile schedule.

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
  tile-packed schedule       a tile owns a contiguous range of node rows; everything it needs is stored in
                             per-tile segments that start on 16-byte boundaries (each segment is fetched with
                             one bulk copy into shared memory):
    cta_hdr[n_cta, 12]         n_items, n_elems, n_inc, n_split, n_nodes, the five segment offsets, first block,
                               n_pieces (pieces-first layout, see build_schedule; 0 = inline layout)
    item_block/inc/meta        work items: a block, or a piece of a block's incidence list - lists are cut where
                               the material of the elements changes (every item is material-uniform, the kernel
                               applies the material constants once per item) and long uniform runs (the diagonal
                               blocks) into chunks; items are stored in block order, so consecutive threads
                               write consecutive addresses of every plane (coalesced stores);
                               item_inc = first incidence entry (tile-relative) | tile-local block << 16,
                               item_meta = length | (partial slot + 1) << 8 | material << 24
    inc                        incidence entries (tile-local element << 4) | (a << 2) | b, block after block,
                               ascending element inside a block (the reference's summation order)
    split_block/meta           blocks with more than one item: tile-local block, first slot | n_items << 16
    cta_elems, cta_conn8       elements the tile stages, and their connectivity as 4 tile-local node bytes
    cta_nodes                  nodes whose coordinates the tile stages
"""
import os
from dataclasses import dataclass

import torch

ITEMS_PER_CTA = 512          # blocks per tile = two passes of the 256 threads of tefem_assemble.cu (B200 sweep, profiles/README.md)
ELEM_BUDGET_PER_CTA = 620    # bound on the sum of node degrees per tile (>= staged elements; ~480 on Kuhn meshes)
MAX_CHUNK = 8                # lists longer than 2 * MAX_CHUNK are split into chunks of <= MAX_CHUNK
ITEM_SORT_WINDOW = 1         # > 1: items ordered longest-first inside windows of this many consecutive items (measured:
                             # helps the K-only launch by <= 12 %, scatters the 29 plane stores of M + D + K: off)
MAX_SMEM_BYTES = 220 * 1024  # dynamic shared memory limit of assemble_kernel (csrc/tefem_assemble.cu)
MAX_CTA_ELEMS = 4096         # 12 bits of the packed incidence entry
MAX_CTA_NODES = 256          # tile-local node numbers are bytes
# True: the pieces of split blocks come first in a tile's item list and every block has one item in block order behind
# them (see build_schedule); False: pieces inline + split lists (build_schedule(pieces_first=...) for A/B measurements).
PIECES_FIRST = True
# Bank-aware tile-local numbering of the staged elements (optimize_banks).  On by default: it cannot change a value (the
# kernel and the lists are untouched, CPU emulation: bitwise) and the CPU model that reproduces ncu's wavefront counts predicts
# -22 % / -15 % gather wavefronts; measured under ncu in round 2: -16 % / -15 % of all shared-memory wavefronts
# (profiles/README.md).  build_schedule(bank_opt=False) keeps the staged order.
BANK_OPT = True
# Fixed-pitch placement of the incidence lists (pitch_lists): values untouched, kernel untouched, modelled 1.7 -> 0.7
# wavefronts per element for the 2-byte entry loads - but ~20 % more entries per tile, and the default tiles leave only
# 7.8 KB of shared memory before the fused M + D + K launch drops from two CTAs per SM to one (2 x (111.9 + 1) KB of 228 KB):
# with the lists pitched the margin shrinks to < 1 KB.  OFF by default (build_schedule(pitch=True) enables; measured in
# round 2 with 448-block tiles: slower, profiles/README.md); when enabled it is skipped for schedules where it would cost
# a CTA per SM.
PITCH_LISTS = False
SM_SMEM_BYTES = 228 * 1024     # shared memory of a B200 SM; every resident CTA reserves 1 KB of it
HDR_STRIDE = 12
H_NITEMS, H_NELEMS, H_NINC, H_NSPLIT, H_NNODES, H_OITEMS, H_OELEMS, H_OINC, H_OSPLIT, H_ONODES, H_BLOCK0, H_NPIECE = range(12)


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
    block_inc: torch.Tensor  # int32 [P] absolute position of each block's incidence list in `inc`
    block_len: torch.Tensor  # int32 [P] its length
    inc: torch.Tensor        # int16 [padded] (read as uint16 by the kernels)
    n_cta: int
    max_cta_items: int
    max_cta_elems: int
    max_cta_inc: int
    max_cta_slots: int
    max_cta_split: int
    max_cta_nodes: int
    cta_hdr: torch.Tensor        # int32 [n_cta, 12]
    cta_elems: torch.Tensor      # int32 [padded] (-1 in the padding)
    cta_conn8: torch.Tensor      # int32 [padded] 4 tile-local node numbers, byte a = node a
    cta_nodes: torch.Tensor      # int32 [padded] (padding repeats node 0 of the tile)
    item_block: torch.Tensor     # int32 [padded]
    item_inc: torch.Tensor       # int32 [padded] tile-relative
    item_meta: torch.Tensor      # int32 [padded]
    split_block: torch.Tensor    # int32 [padded] tile-local block
    split_meta: torch.Tensor     # int32 [padded]
    pieces_first: bool = False   # item layout of the tiles (build_schedule)
    bank_optimized: bool = False # staged elements renumbered per tile by optimize_banks
    lists_pitched: bool = False  # incidence lists re-placed at a fixed pitch by pitch_lists (block_inc is then stale)

    def to(self, device):
        kw = {}
        for k, v in self.__dict__.items():
            kw[k] = v.to(device) if isinstance(v, torch.Tensor) else v
        return Schedule(**kw)

    def h_max(self):
        """The h_max[7] argument of tefem_assemble (shared-memory sizes of the largest tile, item layout)."""
        import numpy as np
        return np.array([self.max_cta_items, self.max_cta_elems, self.max_cta_inc, self.max_cta_slots, self.max_cta_split,
                         self.max_cta_nodes, int(self.pieces_first)], dtype=np.int32)

    def smem_bytes(self, with_D=True, with_K=True):
        """Dynamic shared memory one CTA of assemble_kernel needs for the largest tile (mirrors launch_assemble_pf /
        stage_layout of csrc/tefem_assemble.cu): element records + parked pieces + two prefetch stages + barriers + tables.
        The fused M + D + K launch (default arguments) is the largest instantiation."""
        def up16(b):
            return (b + 15) & ~15
        rec = 17 if with_D else 13
        live = (13 if (with_K or with_D) else 0) + 1 + (4 if with_D else 0)
        slot = live | 1
        rec_doubles = (self.max_cta_elems * rec + 15) & ~15
        part_doubles = (self.max_cta_slots * slot + 15) & ~15
        stage = (64 + up16(24 * ((self.max_cta_nodes + 1) & ~1)) + up16(4 * (self.max_cta_elems + 3)) + 2 * up16(4 * (self.max_cta_items + 3))
                 + 2 * up16(4 * (self.max_cta_split + 3)) + up16(2 * (self.max_cta_inc + 7)))
        return 8 * (rec_doubles + part_doubles) + 2 * stage + 16 + 20 * 8

    def schedule_bytes(self):
        """Bytes of slot-map / schedule data one assembly launch reads (cta_xyz is counted by DeviceMesh)."""
        return sum(t.numel() * t.element_size() for t in
                   (self.inc, self.cta_hdr, self.cta_conn8, self.item_inc, self.item_meta, self.split_block, self.split_meta))


def pitch_lists(S: "Schedule", pitch_blocks: int = 6, pitch_pieces: int = 10, keep_occupancy: bool = True) -> "Schedule":
    """Re-places the incidence lists of every tile at a fixed pitch: item k of a pass starts at entry pitch * (slots before it),
    a list longer than the pitch takes several slots.  The kernel reads its entries as 2-byte shared-memory loads,
    lane l at inc[start_l + step]: with the lists packed back to back the 32 start words of a warp fall into ~12 of the 32
    four-byte banks (ncu: 1.8 wavefronts per element for these loads, 0.8 conflict-free); with a pitch of 3 (blocks: 6
    entries) or 5 (pieces: 10 entries) words - both odd, hence invertible modulo 32 - the start words of 32 consecutive items
    hit 32 different banks at every step.  Costs ~20 % more incidence entries per tile (2 bytes each); items, lists, their
    order and the values are unchanged - only item_inc's start field, the incidence array and its segment sizes change.
    keep_occupancy: return S unchanged when the larger segments would reduce the CTAs per SM of the fused launch."""
    dev = S.cta_hdr.device
    n_cta = S.n_cta
    if n_cta == 0 or S.lists_pitched:
        return S
    i64 = dict(dtype=torch.int64, device=dev)
    H = S.cta_hdr.to(torch.int64)
    n_it, o_it, n_inc, o_inc, n_piece = (H[:, k].contiguous() for k in (H_NITEMS, H_OITEMS, H_NINC, H_OINC, H_NPIECE))
    tot_items = int(n_it.sum().item())
    tile_of_item = torch.repeat_interleave(torch.arange(n_cta, **i64), n_it)
    rank = torch.arange(tot_items, **i64) - (torch.cumsum(n_it, 0) - n_it)[tile_of_item]      # item index inside its tile
    pos = o_it[tile_of_item] + rank                                                          # position in the padded item arrays
    inc_word = S.item_inc[pos].to(torch.int64) & 0xFFFFFFFF
    meta = S.item_meta[pos].to(torch.int64) & 0xFFFFFFFF
    start = inc_word & 0xFFFF
    length = meta & 0xFF
    is_piece = rank < n_piece[tile_of_item]
    if S.pieces_first:                                   # combine items (block pass, slot field set) carry a piece count, no list
        length = torch.where(~is_piece & (((meta >> 8) & 0xFFFF) != 0), torch.zeros_like(length), length)
    pitch = torch.where(is_piece, torch.full_like(length, pitch_pieces), torch.full_like(length, pitch_blocks))
    take = torch.div(length + pitch - 1, pitch, rounding_mode='floor') * pitch
    csum = torch.cumsum(take, 0) - take
    tile_first = (torch.cumsum(n_it, 0) - n_it).clamp(max=max(tot_items - 1, 0))
    new_start = csum - csum[tile_first][tile_of_item]
    new_n_inc = torch.zeros(n_cta, **i64).index_add_(0, tile_of_item, take)
    if int(new_n_inc.max().item()) >= 65536:
        return S                                         # start offsets are 16-bit: keep the packed lists
    padded = torch.div(new_n_inc + 7, 8, rounding_mode='floor') * 8
    new_o_inc = torch.cumsum(padded, 0) - padded
    # ---- move the entries: entry k of item i goes from o_inc + start + k to new_o_inc + new_start + k
    tot_ent = int(length.sum().item())
    item_of_ent = torch.repeat_interleave(torch.arange(tot_items, **i64), length)
    k = torch.arange(tot_ent, **i64) - (torch.cumsum(length, 0) - length)[item_of_ent]
    t_ent = tile_of_item[item_of_ent]
    src = o_inc[t_ent] + start[item_of_ent] + k
    dst = new_o_inc[t_ent] + new_start[item_of_ent] + k
    new_inc = torch.zeros(int(padded.sum().item()) + 8, dtype=S.inc.dtype, device=dev)
    new_inc[dst] = S.inc[src]
    new_item_inc = S.item_inc.clone()
    w = (inc_word & ~0xFFFF) | new_start
    new_item_inc[pos] = torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(S.item_inc.dtype)
    new_hdr = S.cta_hdr.clone()
    new_hdr[:, H_NINC] = new_n_inc.to(new_hdr.dtype)
    new_hdr[:, H_OINC] = new_o_inc.to(new_hdr.dtype)
    assert int(new_o_inc[-1].item()) + int(padded[-1].item()) < 2 ** 31
    kw = dict(S.__dict__)
    kw.update(cta_hdr=new_hdr, item_inc=new_item_inc, inc=new_inc, max_cta_inc=int(new_n_inc.max().item()), lists_pitched=True)
    out = Schedule(**kw)

    def ctas_per_sm(sched, margin):
        return (SM_SMEM_BYTES - margin) // (sched.smem_bytes() + 1024)
    if keep_occupancy and ctas_per_sm(out, 2048) < ctas_per_sm(S, 0):
        return S                                         # the longer incidence segments would cost a resident CTA per SM
    return out


def optimize_banks(S: "Schedule", strides: int = 3, n_sweeps: int = 3, n_threads: int = 0, return_cost: bool = False,
                   chunk: int = 1 << 26):
    """Bank-aware tile-local numbering of the staged elements (native host routine tefem_schedule_bank_numbering,
    csrc/tefem_schedule.cu): returns a Schedule whose cta_elems / cta_conn8 are permuted inside blocks of 16 staged elements
    per tile and whose incidence entries carry the new local numbers.  Items, lists, summation order and therefore the
    assembled values are unchanged (bitwise); only the shared-memory banks the kernel's gathers hit change.
    strides: bit 0 = count the collisions of the kernels without D (record stride 13), bit 1 = with D (stride 17).
    The search runs on the host (threads over tiles) on copies of the schedule arrays; the permutation is applied on the
    device the schedule lives on, in chunks of `chunk` entries."""
    import ctypes

    from . import _lib
    lib = _lib.load()
    dev = S.cta_hdr.device
    n_cta = S.n_cta
    if n_cta == 0:
        return (S, (0, 0)) if return_cost else S
    hdr = S.cta_hdr.cpu().contiguous()
    item_inc, item_meta, inc_h = S.item_inc.cpu().contiguous(), S.item_meta.cpu().contiguous(), S.inc.cpu().contiguous()
    new_local = torch.zeros(S.cta_elems.numel(), dtype=torch.int32)
    n_slots = torch.zeros(n_cta, dtype=torch.int32)
    cost = (ctypes.c_int64 * 2)()
    p = _lib.ptr
    _lib.check(lib.tefem_schedule_bank_numbering(n_cta, p(hdr), p(item_inc), p(item_meta), p(inc_h), int(strides), int(n_sweeps),
                                                 int(n_threads), p(new_local), p(n_slots), cost))
    del item_inc, item_meta, inc_h
    assert torch.equal(n_slots, hdr[:, H_NELEMS])                    # a bijection: the segment sizes do not change
    H = S.cta_hdr.to(torch.int64)
    n_el, o_el, n_inc, o_inc = (H[:, k].contiguous() for k in (H_NELEMS, H_OELEMS, H_NINC, H_OINC))
    new_local = new_local.to(dev)
    # ---- staged element arrays: entry q of tile t moves to o_el[t] + new_local[q]
    new_elems, new_conn8 = S.cta_elems.clone(), S.cta_conn8.clone()
    n_staged = int(S.cta_elems.numel())
    for s0 in range(0, n_staged, chunk):
        q = torch.arange(s0, min(s0 + chunk, n_staged), dtype=torch.int64, device=dev)
        t = torch.searchsorted(o_el, q, right=True) - 1
        valid = (q - o_el[t]) < n_el[t]
        q, t = q[valid], t[valid]
        dest = o_el[t] + new_local[q].to(torch.int64)
        new_elems[dest] = S.cta_elems[q]
        new_conn8[dest] = S.cta_conn8[q]
    # ---- incidence entries: element field -> new local number
    new_inc = S.inc.clone()
    n_ent = int(S.inc.numel())
    for s0 in range(0, n_ent, chunk):
        q = torch.arange(s0, min(s0 + chunk, n_ent), dtype=torch.int64, device=dev)
        t = torch.searchsorted(o_inc, q, right=True) - 1
        valid = (q - o_inc[t]) < n_inc[t]
        q, t = q[valid], t[valid]
        w = S.inc[q].to(torch.int64) & 0xFFFF
        w = (new_local[o_el[t] + (w >> 4)].to(torch.int64) << 4) | (w & 15)
        new_inc[q] = torch.where(w >= 32768, w - 65536, w).to(torch.int16)
    kw = dict(S.__dict__)
    kw.update(cta_elems=new_elems, cta_conn8=new_conn8, inc=new_inc, bank_optimized=True)
    out = Schedule(**kw)
    return (out, (int(cost[0]), int(cost[1]))) if return_cost else out


def _excl_cumsum(x):
    return torch.cumsum(x, 0) - x


def _ptr_from_counts(counts, n, dev):
    ptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=ptr[1:])
    return ptr


def _pad_segments(values, seg_of, seg_ptr, n_seg, align, fill):
    """values are grouped by segment (seg_of ascending, seg_ptr their unpadded offsets); returns the array with
    every segment start aligned to `align` entries, the padded offsets [n_seg + 1] and the padded positions."""
    counts = seg_ptr[1:] - seg_ptr[:-1]
    padded = torch.div(counts + (align - 1), align, rounding_mode='floor') * align
    pptr = _ptr_from_counts(padded, n_seg, values.device)
    pos = pptr[seg_of] + (torch.arange(values.numel(), device=values.device, dtype=torch.int64) - seg_ptr[seg_of])
    out = torch.full((int(pptr[-1].item()),), fill, dtype=values.dtype, device=values.device)
    out[pos] = values
    return out, pptr, pos


class _TileTooLarge(Exception):
    pass


def build_schedule(conn: torch.Tensor, n_nodes: int, n_rows: int = None,
                   items_per_cta: int = ITEMS_PER_CTA, elem_budget: int = ELEM_BUDGET_PER_CTA,
                   max_chunk: int = MAX_CHUNK, mat_id: torch.Tensor = None,
                   item_sort_window: int = None, pieces_first: bool = None, tile_cost: str = None,
                   bank_opt: bool = None, pitch: bool = None) -> Schedule:
    """conn: integer tensor [E, 4] in the reference's element order; node numbers < n_nodes.
    n_rows: only rows of nodes < n_rows are assembled (row-partitioned multi-GPU: owned nodes
    are numbered first); default all.
    mat_id: material row of every element (< 256); default: one material (row 0).
    pieces_first: item layout of a tile.  False: items in (block, chunk) order, the pieces of split blocks park
    their values in slots and the split lists name the blocks to combine afterwards.  True: the tile's item list
    starts with ALL pieces of its split blocks (cta_hdr[11] = their number; piece k parks in slot k), followed by
    exactly one item per block in block order - a plain block (its whole list), or a "combine" item
    (item_meta = n_pieces | (first slot + 1) << 8) that adds the parked pieces in list order.  The kernel runs the
    pieces, synchronises, then the block items: warps hold lists of similar length (the diagonal chunks no longer
    sit between the short off-diagonal lists) and EVERY block is stored by the thread at its block-order position
    (no scattered stores of the combined blocks).  Same chunks, same summation order: bitwise the same values.
    pitch: re-place the incidence lists at a fixed pitch (pitch_lists; default PITCH_LISTS).
    bank_opt: renumber the staged elements of every tile so that the kernel's shared-memory gathers hit fewer bank conflicts
    (optimize_banks; default BANK_OPT).
    tile_cost: what `items_per_cta` bounds - 'items' (all work items of the tile's rows; default of the inline layout)
    or 'blocks' (its blocks = the threads of the block-order pass; default of the pieces-first layout).

    A tile must not reference more than 256 nodes / 4096 elements (byte / 12-bit local numbers) and must fit the
    kernel's shared memory (Schedule.smem_bytes); meshes whose node numbering has little locality (e.g. the Gmsh
    numbering of data/sandwich.msh), very high valences or oversized requests get smaller tiles."""
    pieces_first = PIECES_FIRST if pieces_first is None else bool(pieces_first)
    if tile_cost is None:
        tile_cost = 'blocks' if pieces_first else 'items'
    assert tile_cost in ('items', 'blocks')
    for _ in range(12):
        try:
            S = _build_schedule(conn, n_nodes, n_rows, items_per_cta, elem_budget, max_chunk, mat_id,
                                ITEM_SORT_WINDOW if item_sort_window is None else item_sort_window,
                                pieces_first, tile_cost)
            if S.smem_bytes() <= MAX_SMEM_BYTES:            # the kernel would refuse the launch: smaller tiles
                if (PITCH_LISTS if pitch is None else pitch):
                    S = pitch_lists(S)
                if not (BANK_OPT if bank_opt is None else bank_opt):
                    return S
                try:
                    return optimize_banks(S)
                except Exception as exc:                    # an optimisation, not a requirement: keep the staged order
                    import warnings
                    warnings.warn(f"bank-aware numbering skipped: {exc}")
                    return S
        except _TileTooLarge:
            pass
        items_per_cta = max(16, (3 * items_per_cta) // 4)
        elem_budget = max(32, (3 * elem_budget) // 4)
    raise RuntimeError("could not build a tile schedule within the shared-memory index limits")


def _build_schedule(conn, n_nodes, n_rows, items_per_cta, elem_budget, max_chunk, mat_id=None, item_sort_window=1,
                    pieces_first=False, tile_cost='items'):
    assert conn.dim() == 2 and conn.shape[1] == 4
    assert 1 <= max_chunk <= 127
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
    del ukey
    inc_ptr = _ptr_from_counts(counts, P, dev)
    n_inc = int(src.numel())
    # ---- material-uniform runs of every block's incidence list ("sub-blocks", in list order)
    blk_of_inc = torch.repeat_interleave(torch.arange(P, **i64), counts)
    new_run = torch.zeros(n_inc, dtype=torch.bool, device=dev)
    new_run[inc_ptr[:-1]] = True
    if mat_id is not None:
        mat_id = mat_id.to(device=dev, dtype=torch.int64)
        assert int(mat_id.numel()) == E and (E == 0 or (int(mat_id.min()) >= 0 and int(mat_id.max()) < 256))
        mat_of_inc = mat_id[torch.div(src, 16, rounding_mode='floor')]
        new_run[1:] |= mat_of_inc[1:] != mat_of_inc[:-1]
    run_start = torch.nonzero(new_run).reshape(-1)                               # unpadded absolute position
    n_runs = int(run_start.numel())
    run_block = blk_of_inc[run_start]
    run_len = torch.diff(run_start, append=torch.tensor([n_inc], **i64))
    run_mat = mat_of_inc[run_start] if mat_id is not None else torch.zeros(n_runs, **i64)
    del blk_of_inc, new_run
    blocks_per_row = torch.bincount(block_row, minlength=R)
    rowptr = _ptr_from_counts(blocks_per_row, R, dev)
    diag_idx = torch.nonzero(block_row == block_col).reshape(-1)
    assert int(diag_idx.numel()) == R, "every owned node must belong to at least one element"

    # ---- work items: lists longer than 2 * max_chunk (the diagonal blocks) are split into balanced chunks
    n_chunks = torch.where(run_len > 2 * max_chunk, torch.div(run_len + (max_chunk - 1), max_chunk, rounding_mode='floor'),
                           torch.ones_like(run_len))                             # per run
    items_per_block = torch.zeros(P, **i64).index_add_(0, run_block, n_chunks)
    items_per_row = torch.zeros(R, **i64).index_add_(0, block_row, items_per_block)

    # ---- rows -> tiles: cumulative cost, integer arithmetic (work items vs. staged-element budget)
    degree = torch.bincount(conn.reshape(-1), minlength=N)[:R]
    cost = torch.maximum((items_per_row if tile_cost == 'items' else blocks_per_row) * elem_budget, degree * items_per_cta)
    unit = items_per_cta * elem_budget
    cta_of_row = torch.div(_excl_cumsum(cost), unit, rounding_mode='floor')
    _, cta_of_row = torch.unique_consecutive(cta_of_row, return_inverse=True)    # compact ids
    n_cta = int(cta_of_row[-1].item()) + 1 if R > 0 else 0
    cta_of_block = cta_of_row[block_row]

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
    del sconn, k4, unodes, lnode

    # ---- packed incidence entries: (tile-local element << 4) | (a << 2) | b
    e_src = torch.div(src, 16, rounding_mode='floor')
    ab = src - e_src * 16
    cta_of_inc = torch.repeat_interleave(cta_of_block, counts)
    loc = torch.searchsorted(k2, cta_of_inc * E + e_src) - cta_elem_ptr[cta_of_inc]
    del k2, e_src, src
    packed = (loc << 4) | ab
    del loc, ab
    packed = torch.where(packed >= 32768, packed - 65536, packed).to(torch.int16)

    # ---- tile-packed segments (16-byte aligned starts)
    blocks_per_cta = torch.bincount(cta_of_block, minlength=n_cta)
    cta_blk_ptr = _ptr_from_counts(blocks_per_cta, n_cta, dev)
    cta_inc_ptr = inc_ptr[cta_blk_ptr]                                           # unpadded incidence range of each tile
    inc_p, inc_pptr, inc_pos = _pad_segments(packed, cta_of_inc, cta_inc_ptr, n_cta, 8, 0)
    del packed, cta_of_inc
    block_inc_abs = inc_pos[inc_ptr[:-1]]                                        # every block has >= 1 incidence
    run_inc_abs = inc_pos[run_start]
    del inc_pos
    max_cta_inc = int((cta_inc_ptr[1:] - cta_inc_ptr[:-1]).max().item()) if n_cta else 0

    # ---- item arrays in (block, chunk) order
    n_items = int(n_chunks.sum().item())
    item_run = torch.repeat_interleave(torch.arange(n_runs, **i64), n_chunks)
    item_blk = run_block[item_run]                                               # ascending: runs are in block order
    chunk = torch.arange(n_items, **i64) - _excl_cumsum(n_chunks)[item_run]
    base_len = torch.div(run_len, n_chunks, rounding_mode='floor')[item_run]
    rem = (run_len - torch.div(run_len, n_chunks, rounding_mode='floor') * n_chunks)[item_run]
    item_len = base_len + (chunk < rem).to(torch.int64)
    if n_items and int(item_len.max().item()) > 255:
        raise _TileTooLarge()
    cta_of_item = cta_of_block[item_blk]
    # first entry relative to the tile's (padded) incidence segment
    item_inc = run_inc_abs[item_run] - inc_pptr[cta_of_item] + chunk * base_len + torch.minimum(chunk, rem)
    cta_item_ptr = _ptr_from_counts(torch.bincount(cta_of_item, minlength=n_cta), n_cta, dev)
    max_cta_items = int((cta_item_ptr[1:] - cta_item_ptr[:-1]).max().item()) if n_cta else 0
    blocks_per_cta = torch.bincount(cta_of_block, minlength=n_cta)
    if max_cta_inc >= 65536 or (n_cta and int(blocks_per_cta.max().item()) >= 65536):
        raise _TileTooLarge()
    # blocks with several items park partial sums in shared-memory slots (rank among the tile's split items)
    is_split = items_per_block[item_blk] > 1
    split_run = _excl_cumsum(is_split.to(torch.int64))                           # global rank among split items
    slot = split_run - split_run[cta_item_ptr[:-1].clamp(max=max(n_items - 1, 0))][cta_of_item]
    item_meta = item_len | torch.where(is_split, (slot + 1) << 8, torch.zeros_like(slot)) | (run_mat[item_run] << 24)
    item_meta = torch.where(item_meta >= 2 ** 31, item_meta - 2 ** 32, item_meta)            # int32 bit pattern
    local_blk = item_blk - cta_blk_ptr[cta_of_item]                              # tiles own contiguous block ranges
    item_inc = item_inc | (local_blk << 16)
    blk_split = torch.nonzero(items_per_block > 1).reshape(-1)
    first_item = _excl_cumsum(items_per_block)[blk_split]
    split_meta = slot[first_item] | (items_per_block[blk_split] << 16)
    cta_of_split = cta_of_block[blk_split]
    split_local = blk_split - cta_blk_ptr[cta_of_split]
    cta_split_ptr = _ptr_from_counts(torch.bincount(cta_of_split, minlength=n_cta), n_cta, dev)
    max_cta_split = int((cta_split_ptr[1:] - cta_split_ptr[:-1]).max().item()) if n_cta else 0
    slots_per_cta = torch.zeros(n_cta, **i64).index_add_(0, cta_of_item, is_split.to(torch.int64))
    max_cta_slots = int(slots_per_cta.max().item()) if n_cta else 0
    assert max_cta_slots < 65536
    # inside windows of ITEM_SORT_WINDOW consecutive items of a tile: longest lists first (fuller warps in the
    # incidence loop) - the blocks a warp writes stay within one window, so its plane stores remain a few lines
    if item_sort_window > 1 and n_items:
        rank_in_cta = torch.arange(n_items, **i64) - cta_item_ptr[cta_of_item]
        win = torch.div(rank_in_cta, item_sort_window, rounding_mode='floor')
        maxlen = int(item_len.max().item())
        nwin = int(win.max().item()) + 1
        perm = torch.sort((cta_of_item * nwin + win) * (maxlen + 1) + (maxlen - item_len), stable=True)[1]
        item_blk, item_inc, item_meta = item_blk[perm], item_inc[perm], item_meta[perm]
        del perm, rank_in_cta, win                                               # cta_of_item is unchanged by the permutation

    n_piece_cta = torch.zeros(n_cta, **i64)
    if pieces_first and n_items:
        assert item_sort_window <= 1, "the windowed item order applies to the inline layout only"
        # pieces of split blocks (already in (tile, block, chunk) order; piece k of a tile parks in slot k) ...
        pc = torch.nonzero(is_split).reshape(-1)
        n_piece_cta = slots_per_cta
        # ... then one item per block, in block order
        first_item_all = _excl_cumsum(items_per_block)
        blk_split_mask = items_per_block > 1
        if int(items_per_block.max().item()) > 255:
            raise RuntimeError("a block has more than 255 pieces (material changes x chunks)")
        b_inc = torch.where(blk_split_mask, torch.zeros(P, **i64), item_inc[first_item_all] & 0xffff)
        b_inc = b_inc | ((torch.arange(P, **i64) - cta_blk_ptr[cta_of_block]) << 16)
        b_meta = torch.where(blk_split_mask, items_per_block | ((slot[first_item_all] + 1) << 8) | (run_mat[item_run[first_item_all]] << 24),
                             item_meta[first_item_all])
        b_meta = torch.where(b_meta >= 2 ** 31, b_meta - 2 ** 32, b_meta)
        cat_cta = torch.cat([cta_of_item[pc], cta_of_block])
        perm = torch.sort(cat_cta * 2 + torch.cat([torch.zeros_like(pc), torch.ones(P, **i64)]), stable=True)[1]
        item_blk = torch.cat([item_blk[pc], torch.arange(P, **i64)])[perm]
        item_inc = torch.cat([item_inc[pc], b_inc])[perm]
        item_meta = torch.cat([item_meta[pc], b_meta])[perm]
        cta_of_item = cat_cta[perm]
        n_items = int(item_blk.numel())
        cta_item_ptr = _ptr_from_counts(torch.bincount(cta_of_item, minlength=n_cta), n_cta, dev)
        max_cta_items = int((cta_item_ptr[1:] - cta_item_ptr[:-1]).max().item())
        # no split lists in this layout
        split_local = split_local[:0]
        split_meta = split_meta[:0]
        cta_of_split = cta_of_split[:0]
        cta_split_ptr = torch.zeros(n_cta + 1, **i64)
        max_cta_split = 0
        del pc, perm, cat_cta, b_inc, b_meta
    item_block_p, item_pptr, _ = _pad_segments(item_blk, cta_of_item, cta_item_ptr, n_cta, 4, 0)
    item_inc_p, _, _ = _pad_segments(item_inc, cta_of_item, cta_item_ptr, n_cta, 4, 0)
    item_meta_p, _, _ = _pad_segments(item_meta, cta_of_item, cta_item_ptr, n_cta, 4, 0)      # padding: length 0
    split_block_p, split_pptr, _ = _pad_segments(split_local, cta_of_split, cta_split_ptr, n_cta, 4, 0)
    split_meta_p, _, _ = _pad_segments(split_meta, cta_of_split, cta_split_ptr, n_cta, 4, 0)
    if split_block_p.numel() == 0:                                               # keep the pointers valid (16-byte aligned)
        split_block_p = torch.zeros(4, **i64)
        split_meta_p = torch.zeros(4, **i64)
    elems_p, elem_pptr, _ = _pad_segments(cta_elems, cta_of_k2, cta_elem_ptr, n_cta, 4, -1)
    conn8_p, _, _ = _pad_segments(conn8, cta_of_k2, cta_elem_ptr, n_cta, 4, 0)
    nodes_p, node_pptr, _ = _pad_segments(cta_nodes, cta_of_un, cta_node_ptr, n_cta, 2, -1)
    # padding entries of the node list repeat the tile's first node (a valid address for the coordinate copy)
    if nodes_p.numel():
        first_node = cta_nodes[cta_node_ptr[:-1]]
        tile_of_slot = torch.searchsorted(node_pptr[1:], torch.arange(nodes_p.numel(), **i64), right=True)
        nodes_p = torch.where(nodes_p < 0, first_node[tile_of_slot.clamp(max=max(n_cta - 1, 0))], nodes_p)

    def cnt(ptr):
        return ptr[1:] - ptr[:-1]
    cta_hdr = torch.stack([cnt(cta_item_ptr), cnt(cta_elem_ptr), cnt(cta_inc_ptr), cnt(cta_split_ptr), cnt(cta_node_ptr),
                           item_pptr[:-1], elem_pptr[:-1], inc_pptr[:-1], split_pptr[:-1], node_pptr[:-1], cta_blk_ptr[:-1],
                           n_piece_cta], dim=1)
    assert int(inc_pptr[-1].item()) < 2 ** 31

    i32 = torch.int32
    return Schedule(n_nodes=N, n_rows=R, n_blocks=P, n_elems=E, n_items=n_items,
                    rowptr=rowptr.to(i32), colidx=block_col.to(i32), block_row=block_row.to(i32),
                    diag_idx=diag_idx.to(i32), block_inc=block_inc_abs.to(i32), block_len=counts.to(i32), inc=inc_p,
                    n_cta=n_cta, max_cta_items=max_cta_items, max_cta_elems=max_cta_elems, max_cta_inc=max_cta_inc,
                    max_cta_slots=max_cta_slots, max_cta_split=max_cta_split, max_cta_nodes=max_cta_nodes,
                    cta_hdr=cta_hdr.to(i32).contiguous(), cta_elems=elems_p.to(i32), cta_conn8=conn8_p.to(i32),
                    cta_nodes=nodes_p.to(i32), item_block=item_block_p.to(i32), item_inc=item_inc_p.to(i32),
                    item_meta=item_meta_p.to(i32), split_block=split_block_p.to(i32), split_meta=split_meta_p.to(i32),
                    pieces_first=bool(pieces_first))
