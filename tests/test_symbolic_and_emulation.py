"""Symbolic phase (pattern, slot map, CTA schedule) and the K1 per-thread routines, on the CPU.

The per-thread routines of the CUDA kernel are compiled for the host by tests/host_emulation/build_emu.py
(test infrastructure, never loaded by the package) and executed sequentially; the result must match the oracle."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

from oracle import tefem_oracle as orc
from thermoelasticity_fem_b200 import layout
from thermoelasticity_fem_b200.symbolic import build_schedule
from thermoelasticity_fem_b200.synthetic import kuhn_box
import parity

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emulation"))

MAT2 = (2300, 64e9, 0.1, 1., 1., 4e-6, 20.)


@pytest.fixture(scope="module")
def emu():
    import build_emu
    return ctypes.CDLL(build_emu.build())


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def planes_to_csc(vals, pmap, S, n):
    import scipy.sparse as sp
    br = S.block_row.numpy().astype(np.int64)
    bc = S.colidx.numpy().astype(np.int64)
    r_, c_, v_ = [], [], []
    for r in range(4):
        for c in range(4):
            pl = pmap[r, c]
            if pl >= 0:
                r_.append(4 * br + r)
                c_.append(4 * bc + c)
                v_.append(vals[pl])
    A = sp.csr_array((np.concatenate(v_), (np.concatenate(r_), np.concatenate(c_))), shape=(4 * S.n_rows, n)).tocsc()
    A.sort_indices()
    return A


def emu_assemble(emu, S, pts, conn, mat_id, rows, ops, aM, aK):
    P = S.n_blocks
    nD = layout.n_planes(layout.plane_map_D(aM, aK))
    vM, vK, vD = np.full((3, P), np.nan), np.full((13, P), np.nan), np.full((nD, P), np.nan)
    bad = np.full(2, 2 ** 31 - 1, dtype=np.int32)
    a = {k: getattr(S, k).numpy() for k in ("cta_ptr", "cta_elems", "cta_conn8", "cta_nodes", "item_block", "item_blockT",
                                            "item_inc", "item_meta", "split_block", "split_meta", "inc")}
    cta_mat = np.ascontiguousarray(mat_id.astype(np.int32)[a["cta_elems"]])
    mask = sum({"M": 1, "K": 2, "D": 4}[c] for c in ops)
    rc = emu.tefem_emu_assemble(ptr(pts), ptr(rows), ctypes.c_int32(S.n_cta), ctypes.c_int32(S.max_cta_elems),
                                ctypes.c_int32(S.max_cta_inc), ctypes.c_int32(S.max_cta_slots), ctypes.c_int32(S.max_cta_nodes),
                                ptr(a["cta_ptr"]), ptr(a["cta_elems"]), ptr(a["cta_conn8"]), ptr(cta_mat), ptr(a["cta_nodes"]),
                                ptr(a["item_block"]), ptr(a["item_blockT"]), ptr(a["item_inc"]),
                                ptr(a["item_meta"]), ptr(a["split_block"]), ptr(a["split_meta"]),
                                ptr(a["inc"]), ctypes.c_int64(P), mask, ctypes.c_double(aM), ctypes.c_double(aK),
                                ptr(vM), ptr(vK), ptr(vD), ptr(bad))
    assert rc == 0
    for nm, v in (("M", vM), ("K", vK), ("D", vD)):
        if nm in ops and bad.min() == 2 ** 31 - 1:                               # (a bad Jacobian yields inf / nan values)
            assert not np.isnan(v).any(), f"{nm}: some block was never written"     # every value written (exactly once)
    return dict(M=vM, K=vK, D=vD), bad


def check_schedule_invariants(S, conn, n_nodes, max_chunk=6):
    E = conn.shape[0]
    rowptr, colidx, br = S.rowptr.numpy(), S.colidx.numpy(), S.block_row.numpy()
    assert rowptr[0] == 0 and rowptr[-1] == S.n_blocks
    for i in range(0, S.n_rows, max(1, S.n_rows // 50)):
        cols = colidx[rowptr[i]:rowptr[i + 1]]
        assert np.all(np.diff(cols) > 0)
        assert colidx[S.diag_idx.numpy()[i]] == i
    incp = S.inc_ptr.numpy().astype(np.int64)
    lb = S.listed_block.numpy()
    assert np.array_equal(lb, np.nonzero(colidx >= br)[0])               # diagonal + upper blocks carry the lists
    cp = S.cta_ptr.numpy()
    ib, ibT, ii, im = S.item_block.numpy(), S.item_blockT.numpy(), S.item_inc.numpy().astype(np.int64), S.item_meta.numpy()
    sb, sm = S.split_block.numpy(), S.split_meta.numpy()
    assert cp[0, 0] == 0 and cp[-1, 1] == S.n_items and np.array_equal(cp[1:, 0], cp[:-1, 1])
    assert np.array_equal(cp[1:, 4], cp[:-1, 5]) and cp[-1, 5] == incp[-1]
    ilen = im & 0xff
    upper = (im >> 24) & 1
    assert ilen.min() >= 1 and ilen[upper == 0].max() <= max_chunk
    assert np.array_equal(upper == 1, colidx[ib] > br[ib])
    # mirrors: block (j, i) of an upper block (i, j), or -1 when row j is not owned; every lower block is produced once
    has = ibT >= 0
    assert np.all(upper[has] == 1) and np.array_equal(br[ibT[has]], colidx[ib[has]]) and np.array_equal(colidx[ibT[has]], br[ib[has]])
    assert np.all(colidx[ib[(upper == 1) & ~has]] >= S.n_rows)
    assert np.array_equal(np.sort(ibT[has]), np.nonzero(colidx < br)[0])
    # the chunks of every listed block tile its incidence list exactly, in order
    pos_of_block = np.full(S.n_blocks, -1)
    pos_of_block[lb] = np.arange(lb.size)
    order = np.lexsort((ii, ib))
    b_sorted, s_sorted, l_sorted = ib[order], ii[order], ilen[order]
    first = np.r_[True, b_sorted[1:] != b_sorted[:-1]]
    assert np.array_equal(s_sorted[first], incp[pos_of_block[b_sorted[first]]])
    assert np.array_equal((s_sorted + l_sorted)[~np.r_[first[1:], True]], s_sorted[~first])
    last = np.r_[first[1:], True]
    assert np.array_equal((s_sorted + l_sorted)[last], incp[pos_of_block[b_sorted[last]] + 1])
    assert np.array_equal(np.unique(ib), lb)
    cn8 = S.cta_conn8.numpy().astype(np.int64) & 0xFFFFFFFF
    ce, cnodes = S.cta_elems.numpy(), S.cta_nodes.numpy()
    for c in range(S.n_cta):
        i0, i1, e0, e1, n0, n1, s0, s1, k0, k1 = cp[c, :10]
        assert e1 - e0 <= S.max_cta_elems and n1 - n0 <= S.max_cta_inc and k1 - k0 <= S.max_cta_nodes <= 256
        assert np.all(ii[i0:i1] >= n0) and np.all(ii[i0:i1] + ilen[i0:i1] <= n1)
        slots = ((im[i0:i1] >> 8) & 0xffff) - 1
        used = np.sort(slots[slots >= 0])
        assert np.array_equal(used, np.arange(used.size)) and used.size <= S.max_cta_slots
        tot = 0
        for k in range(s0, s1):
            slot0, n = sm[k] & 0xffff, sm[k] >> 16
            sel = (ib[i0:i1] == sb[k])
            assert n >= 2 and sel.sum() == n
            ss = slots[sel][np.argsort(ii[i0:i1][sel])]
            assert np.array_equal(ss, np.arange(slot0, slot0 + n))      # chunk order = slot order
            tot += n
        assert tot == used.size
        # byte connectivity decodes to the global connectivity through the tile's node list
        loc = np.stack([(cn8[e0:e1] >> (8 * a)) & 0xff for a in range(4)], axis=1)
        assert loc.max() < k1 - k0 and np.array_equal(cnodes[k0:k1][loc], conn[ce[e0:e1]])
    assert S.inc.numel() == incp[-1] + 8
    # every incidence entry decodes to an element of the tile's list that really contains (row, col)
    inc = S.inc.numpy().astype(np.int64) & 0xFFFF
    l_ptr = np.searchsorted(incp, cp[:, 4])                            # first listed block of each tile
    cta_of_listed = np.searchsorted(l_ptr, np.arange(lb.size), side="right") - 1
    rng = np.random.default_rng(0)
    for q in rng.integers(0, lb.size, size=400):
        p, c = lb[q], cta_of_listed[q]
        last_e = -1
        for k in range(incp[q], incp[q + 1]):
            w = inc[k]
            e = ce[cp[c, 2] + (w >> 4)]
            a, b = (w >> 2) & 3, w & 3
            assert conn[e, a] == br[p] and conn[e, b] == colidx[p]
            assert e > last_e                            # ascending element order = the reference's summation order
            last_e = e
    if S.n_rows == n_nodes:
        assert incp[-1] == 10 * E                        # 4 diagonal + 6 upper incidences per element


@pytest.mark.parametrize("aM,aK", [(0.2, 0.2), (0.3, 0.0), (0.0, 0.0)])
def test_sandwich_schedule_and_emulated_assembly(sandwich, emu, aM, aK):
    pts, blocks, groups = sandwich
    m = orc.Material(*MAT2)
    conn, mat_id, rows = orc.element_table(groups["tet"], {1: m, 2: m, 3: m})
    S = build_schedule(torch.from_numpy(conn), pts.shape[0])
    assert S.n_blocks == 39985
    if aM == 0.2:
        check_schedule_invariants(S, conn, pts.shape[0])
    vals, bad = emu_assemble(emu, S, pts, conn, mat_id, rows, "MKD", aM, aK)
    assert bad[0] == 2 ** 31 - 1 and bad[1] == 2 ** 31 - 1
    ops = orc.assemble(pts, conn, rows, mat_id, aM, aK)
    ops["M"] = orc.drop_component_zeros_M(ops["M"])
    if aK == 0. and aM != 0.:
        ops["D"] = orc.drop_cross_component_uu(ops["D"])
    n = 4 * pts.shape[0]
    for nm, pm in (("M", layout.plane_map_M()), ("K", layout.plane_map_K()), ("D", layout.plane_map_D(aM, aK))):
        A = planes_to_csc(vals[nm], pm, S, n)
        parity.check_pattern(A, None, ops[nm])
        assert parity.scaled_error(A, ops[nm]) < parity.VALUE_TOL, nm


def test_single_operator_launches_match_fused(sandwich, emu):
    pts, blocks, groups = sandwich
    m = orc.Material(*MAT2)
    conn, mat_id, rows = orc.element_table(groups["tet"], {1: m, 2: m, 3: m})
    S = build_schedule(torch.from_numpy(conn), pts.shape[0])
    fused, _ = emu_assemble(emu, S, pts, conn, mat_id, rows, "MKD", 0.2, 0.2)
    for nm in "MKD":
        single, _ = emu_assemble(emu, S, pts, conn, mat_id, rows, nm, 0.2, 0.2)
        assert np.array_equal(single[nm], fused[nm]), nm


def test_kuhn_multimaterial_and_bad_jacobian(emu):
    pts, tets, tris, nodes = kuhn_box(5, 3, 4, jitter=0.2, seed=1)
    mats = {1: orc.Material(2300, 64e9, 0.1, 1., 750., 4e-6, 20.), 2: orc.Material(7800, 210e9, 0.3, 45., 460., 1.2e-5, 25.),
            3: orc.Material(2700, 70e9, 0.33, 200., 900., 2.3e-5, 30.)}
    conn, mat_id, rows = orc.element_table(tets, mats)
    for items in (256, 64):                       # small tiles: many CTAs, exercises the staging offsets
        chunk = 6 if items == 256 else 3
        S = build_schedule(torch.from_numpy(conn), pts.shape[0], items_per_cta=items, elem_budget=max(160, items),
                           max_chunk=chunk)
        check_schedule_invariants(S, conn, pts.shape[0], chunk)
        vals, bad = emu_assemble(emu, S, pts, conn, mat_id, rows, "MKD", 0.1, 0.4)
        ops = orc.assemble(pts, conn, rows, mat_id, 0.1, 0.4)
        A = planes_to_csc(vals["K"], layout.plane_map_K(), S, 4 * pts.shape[0])
        parity.check_pattern(A, None, ops["K"])
        assert parity.scaled_error(A, ops["K"]) < parity.VALUE_TOL
        D = planes_to_csc(vals["D"], layout.plane_map_D(0.1, 0.4), S, 4 * pts.shape[0])
        assert parity.scaled_error(D, ops["D"]) < parity.VALUE_TOL
    bad_conn = conn.copy()
    bad_conn[[17, 90]] = bad_conn[[17, 90]][:, [0, 2, 1, 3]]        # two inverted elements: the first one is reported
    bad_conn[40, 3] = bad_conn[40, 2]                               # and a degenerate one in between
    S = build_schedule(torch.from_numpy(bad_conn), pts.shape[0])
    _, bad = emu_assemble(emu, S, pts, bad_conn, mat_id, rows, "K", 0., 0.)
    assert bad[0] == 17 and bad[1] == 40


def test_row_restricted_schedule_matches_global_rows():
    """Row-partitioned assembly (multi-GPU): a schedule restricted to the first R rows has exactly the global
    pattern's first R rows and the same incidence lists."""
    pts, tets, tris, nodes = kuhn_box(6, 3, 3)
    conn = np.concatenate(list(tets.values()))
    N = pts.shape[0]
    full = build_schedule(torch.from_numpy(conn), N)
    R = N // 2
    part = build_schedule(torch.from_numpy(conn), N, n_rows=R)
    nb = int(full.rowptr[R])
    assert part.n_blocks == nb and part.n_rows == R
    assert torch.equal(part.rowptr, full.rowptr[:R + 1]) and torch.equal(part.colidx, full.colidx[:nb])
    nl = int((full.listed_block < nb).sum())
    assert torch.equal(part.listed_block, full.listed_block[:nl]) and torch.equal(part.inc_ptr, full.inc_ptr[:nl + 1])
    # mirrors whose row is not owned are dropped, the others are identical
    assert torch.equal(part.rowptr, full.rowptr[:R + 1])


def test_emulated_element_matrices_match_oracle(emu):
    pts, tets, tris, nodes = kuhn_box(3, 2, 2, jitter=0.25, seed=3)
    mats = {1: orc.Material(2300, 64e9, 0.1, 1., 750., 4e-6, 20.), 2: orc.Material(7800, 210e9, 0.3, 45., 460., 1.2e-5, 25.),
            3: orc.Material(2700, 70e9, 0.33, 200., 900., 2.3e-5, 30.)}
    conn, mat_id, rows = orc.element_table(tets, mats)
    E = conn.shape[0]
    out = {"Muu": np.zeros((E, 12, 12)), "Dtu": np.zeros((E, 4, 12)), "Dtt": np.zeros((E, 4, 4)),
           "Kuu": np.zeros((E, 12, 12)), "Kut": np.zeros((E, 12, 4)), "Ktt": np.zeros((E, 4, 4))}
    bad = np.full(2, 2 ** 31 - 1, dtype=np.int32)
    conn32, mid32 = conn.astype(np.int32), mat_id.astype(np.int32)
    rc = emu.tefem_emu_element_matrices(ptr(pts), ptr(conn32), ptr(mid32), ptr(rows), ctypes.c_int64(E), 63,
                                        ptr(out["Muu"]), ptr(out["Dtu"]), ptr(out["Dtt"]), ptr(out["Kuu"]),
                                        ptr(out["Kut"]), ptr(out["Ktt"]), ptr(bad))
    assert rc == 0
    em = orc.element_matrices_batch(pts, conn, rows, mat_id)
    for nm in out:
        assert np.max(np.abs(out[nm] - em[nm])) <= 1e-14 * np.abs(em[nm]).max(), nm
    # and against the per-call restatement, which is bit-identical to the reference
    for e in range(0, E, 5):
        pc = orc.element_matrices_percall(pts[conn[e]], mats[int(list(tets)[mat_id[e]])])
        for nm in out:
            assert np.max(np.abs(out[nm][e] - pc[nm])) <= 1e-14 * np.abs(pc[nm]).max(), nm
