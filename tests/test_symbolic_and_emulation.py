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
    a = {k: getattr(S, k).numpy() for k in ("cta_hdr", "cta_elems", "cta_conn8", "cta_nodes", "item_inc", "item_meta",
                                            "split_block", "split_meta", "inc")}
    cta_xyz = np.ascontiguousarray(pts[a["cta_nodes"]])
    h_max = S.h_max()
    mask = sum({"M": 1, "K": 2, "D": 4}[c] for c in ops)
    rc = emu.tefem_emu_assemble(ptr(rows), ctypes.c_int(rows.shape[0]), ctypes.c_int32(S.n_cta), ptr(h_max), ptr(a["cta_hdr"]),
                                ptr(a["cta_elems"]), ptr(a["cta_conn8"]), ptr(cta_xyz), ptr(a["item_inc"]),
                                ptr(a["item_meta"]), ptr(a["split_block"]), ptr(a["split_meta"]), ptr(a["inc"]),
                                ctypes.c_int64(P), mask, ctypes.c_double(aM), ctypes.c_double(aK),
                                ptr(vM), ptr(vK), ptr(vD), ptr(bad))
    assert rc == 0
    for nm, v in (("M", vM), ("K", vK), ("D", vD)):
        if nm in ops and bad.min() == 2 ** 31 - 1:                               # (a bad Jacobian yields inf / nan values)
            assert not np.isnan(v).any(), f"{nm}: some block was never written"     # every value written (exactly once)
    return dict(M=vM, K=vK, D=vD), bad


def check_schedule_invariants(S, conn, n_nodes, max_chunk=8, mat_id=None):
    E = conn.shape[0]
    rowptr, colidx, br = S.rowptr.numpy(), S.colidx.numpy(), S.block_row.numpy()
    assert rowptr[0] == 0 and rowptr[-1] == S.n_blocks
    for i in range(0, S.n_rows, max(1, S.n_rows // 50)):
        cols = colidx[rowptr[i]:rowptr[i + 1]]
        assert np.all(np.diff(cols) > 0)
        assert colidx[S.diag_idx.numpy()[i]] == i
    H = S.cta_hdr.numpy().astype(np.int64)
    binc, blen = S.block_inc.numpy().astype(np.int64), S.block_len.numpy().astype(np.int64)
    ib, ii, im = S.item_block.numpy(), S.item_inc.numpy().astype(np.int64), S.item_meta.numpy()
    sb, sm = S.split_block.numpy(), S.split_meta.numpy().astype(np.int64) & 0xFFFFFFFF
    inc = S.inc.numpy().astype(np.int64) & 0xFFFF
    cn8 = S.cta_conn8.numpy().astype(np.int64) & 0xFFFFFFFF
    ce, cnodes = S.cta_elems.numpy(), S.cta_nodes.numpy()
    assert H[:, 0].sum() == S.n_items and blen.sum() == (16 * E if S.n_rows == n_nodes else blen.sum())
    seen_blocks = []
    for c in range(S.n_cta):
        n_it, n_el, n_inc, n_sp, n_nd, o_it, o_el, o_inc, o_sp, o_nd, p0 = H[c, :11]
        # 16-byte aligned segment starts, sizes within the declared maxima
        assert o_it % 4 == 0 and o_el % 4 == 0 and o_inc % 8 == 0 and o_sp % 4 == 0 and o_nd % 2 == 0
        assert n_it <= S.max_cta_items and n_el <= S.max_cta_elems and n_inc <= S.max_cta_inc
        assert n_sp <= S.max_cta_split and n_nd <= S.max_cta_nodes <= 256
        blk, start, meta = ib[o_it:o_it + n_it], ii[o_it:o_it + n_it] & 0xffff, im[o_it:o_it + n_it]
        ilen_all = meta & 0xff
        assert np.array_equal(ii[o_it:o_it + n_it] >> 16, blk - p0)
        # block order up to the longest-first ordering inside windows of consecutive items
        from thermoelasticity_fem_b200.symbolic import ITEM_SORT_WINDOW as W
        for w0 in range(0, n_it, W):
            assert np.all(np.diff(ilen_all[w0:w0 + W]) <= 0)
            if w0 + W < n_it:
                assert blk[w0:w0 + W].max() <= blk[w0 + W:].min()
        ilen = meta & 0xff
        slots = ((meta >> 8) & 0xffff) - 1
        imat = (meta.astype(np.int64) >> 24) & 0xff
        assert ilen.min() >= 1
        assert np.all(start >= 0) and np.all(start + ilen <= n_inc)
        assert np.all(ilen <= 2 * max_chunk)
        # every item is material-uniform and carries that material
        for k in range(0, n_it, max(1, n_it // 40)):
            els = ce[o_el + (inc[o_inc + start[k]:o_inc + start[k] + ilen[k]] >> 4)]
            want = 0 if mat_id is None else mat_id[els]
            assert np.all(want == imat[k])
        used = np.sort(slots[slots >= 0])
        assert np.array_equal(used, np.arange(used.size)) and used.size <= S.max_cta_slots
        # the chunks of every block tile its incidence list exactly, in order
        order = np.lexsort((start, blk))
        b_s, s_s, l_s = blk[order], start[order], ilen[order]
        first = np.r_[True, b_s[1:] != b_s[:-1]]
        last = np.r_[first[1:], True]
        assert np.array_equal(s_s[first] + o_inc, binc[b_s[first]])
        assert np.array_equal((s_s + l_s)[~last], s_s[~first])
        assert np.array_equal((s_s + l_s)[last] + o_inc, binc[b_s[last]] + blen[b_s[last]])
        seen_blocks.append(np.unique(blk))
        tot = 0
        for k in range(o_sp, o_sp + n_sp):
            slot0, n = sm[k] & 0xffff, sm[k] >> 16
            sel = blk == p0 + sb[k]
            assert n >= 2 and sel.sum() == n
            assert np.array_equal(slots[sel][np.argsort(start[sel])], np.arange(slot0, slot0 + n))   # list order = slot order
            tot += n
        assert tot == used.size
        # byte connectivity decodes to the global connectivity through the tile's node list
        loc = np.stack([(cn8[o_el:o_el + n_el] >> (8 * a)) & 0xff for a in range(4)], axis=1)
        assert loc.max() < n_nd and np.array_equal(cnodes[o_nd:o_nd + n_nd][loc], conn[ce[o_el:o_el + n_el]])
        # incidence entries decode to elements of the tile that really contain (row, col), ascending element
        for p in np.unique(blk)[:: max(1, len(np.unique(blk)) // 6)]:
            last_e = -1
            for k in range(binc[p], binc[p] + blen[p]):
                w = inc[k]
                e = ce[o_el + (w >> 4)]
                assert conn[e, (w >> 2) & 3] == br[p] and conn[e, w & 3] == colidx[p]
                assert e > last_e                        # = the reference's summation order
                last_e = e
    assert np.array_equal(np.concatenate(seen_blocks), np.arange(S.n_blocks))       # tiles own contiguous block ranges


@pytest.mark.parametrize("aM,aK", [(0.2, 0.2), (0.3, 0.0), (0.0, 0.0)])
def test_sandwich_schedule_and_emulated_assembly(sandwich, emu, aM, aK):
    pts, blocks, groups = sandwich
    m = orc.Material(*MAT2)
    conn, mat_id, rows = orc.element_table(groups["tet"], {1: m, 2: m, 3: m})
    S = build_schedule(torch.from_numpy(conn), pts.shape[0], mat_id=torch.from_numpy(mat_id))
    assert S.n_blocks == 39985
    if aM == 0.2:
        S_inline = build_schedule(torch.from_numpy(conn), pts.shape[0], mat_id=torch.from_numpy(mat_id), pieces_first=False,
                                  pitch=False)       # the invariants below describe the packed lists (block_inc)
        check_schedule_invariants(S_inline, conn, pts.shape[0], mat_id=mat_id)
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
    S = build_schedule(torch.from_numpy(conn), pts.shape[0], mat_id=torch.from_numpy(mat_id))
    fused, _ = emu_assemble(emu, S, pts, conn, mat_id, rows, "MKD", 0.2, 0.2)
    for nm in "MKD":
        single, _ = emu_assemble(emu, S, pts, conn, mat_id, rows, nm, 0.2, 0.2)
        assert np.array_equal(single[nm], fused[nm]), nm


def test_kuhn_multimaterial_and_bad_jacobian(emu):
    pts, tets, tris, nodes = kuhn_box(5, 3, 4, jitter=0.2, seed=1)
    mats = {1: orc.Material(2300, 64e9, 0.1, 1., 750., 4e-6, 20.), 2: orc.Material(7800, 210e9, 0.3, 45., 460., 1.2e-5, 25.),
            3: orc.Material(2700, 70e9, 0.33, 200., 900., 2.3e-5, 30.)}
    conn, mat_id, rows = orc.element_table(tets, mats)
    for items, pf in ((256, False), (256, True), (64, False), (64, True)):     # small tiles: many CTAs, exercises the offsets
        chunk = 6 if items == 256 else 3
        S = build_schedule(torch.from_numpy(conn), pts.shape[0], items_per_cta=items, elem_budget=max(160, items),
                           max_chunk=chunk, mat_id=torch.from_numpy(mat_id), pieces_first=pf)
        if not pf:
            check_schedule_invariants(build_schedule(torch.from_numpy(conn), pts.shape[0], items_per_cta=items,
                                                     elem_budget=max(160, items), max_chunk=chunk, mat_id=torch.from_numpy(mat_id),
                                                     pieces_first=False, pitch=False), conn, pts.shape[0], chunk, mat_id=mat_id)
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
    S = build_schedule(torch.from_numpy(bad_conn), pts.shape[0], mat_id=torch.from_numpy(mat_id))
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
    assert torch.equal(part.block_len, full.block_len[:nb])


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


def check_pieces_first_invariants(S, S0):
    """Pieces-first layout against the inline layout S0 of the same mesh / parameters: same tiles, same pieces (in the
    same order, piece k -> slot k), exactly one item per block in block order, combine items cover their pieces."""
    assert S.pieces_first and not S0.pieces_first
    assert S.n_cta == S0.n_cta and S.max_cta_split == 0 and S.max_cta_slots == S0.max_cta_slots
    H, H0 = S.cta_hdr.numpy().astype(np.int64), S0.cta_hdr.numpy().astype(np.int64)
    assert np.array_equal(H[:, [1, 2, 4, 6, 7, 9, 10]], H0[:, [1, 2, 4, 6, 7, 9, 10]]) and np.all(H[:, 3] == 0)
    assert torch.equal(S.inc, S0.inc) and torch.equal(S.cta_conn8, S0.cta_conn8)
    ii, im, ib = S.item_inc.numpy().astype(np.int64), S.item_meta.numpy().astype(np.int64), S.item_block.numpy()
    ii0, im0, ib0 = S0.item_inc.numpy().astype(np.int64), S0.item_meta.numpy().astype(np.int64), S0.item_block.numpy()
    for c in range(S.n_cta):
        n_it, o_it, p0, n_pc = H[c, 0], H[c, 5], H[c, 10], H[c, 11]
        n0, o0 = H0[c, 0], H0[c, 5]
        slot0 = ((im0[o0:o0 + n0] >> 8) & 0xffff) - 1
        pieces0 = np.nonzero(slot0 >= 0)[0]
        assert n_pc == pieces0.size
        # the pieces are the inline layout's split items, unchanged and in the same order; piece k parks in slot k
        assert np.array_equal(ii[o_it:o_it + n_pc], ii0[o0 + pieces0]) and np.array_equal(im[o_it:o_it + n_pc], im0[o0 + pieces0])
        assert np.array_equal(((im[o_it:o_it + n_pc] >> 8) & 0xffff) - 1, np.arange(n_pc))
        # one item per block, in block order
        nb = n_it - n_pc
        nxt = H[c + 1, 10] if c + 1 < S.n_cta else S.n_blocks
        assert nb == nxt - p0
        bi, bm = ii[o_it + n_pc:o_it + n_it], im[o_it + n_pc:o_it + n_it]
        assert np.array_equal(ib[o_it + n_pc:o_it + n_it], p0 + np.arange(nb)) and np.array_equal(bi >> 16, np.arange(nb))
        bslot = ((bm >> 8) & 0xffff) - 1
        plain = bslot < 0
        # plain blocks: the inline layout's single item of that block
        single0 = np.nonzero(slot0 < 0)[0]
        assert np.array_equal(ib0[o0 + single0], p0 + np.nonzero(plain)[0])
        assert np.array_equal(bi[plain], ii0[o0 + single0]) and np.array_equal(bm[plain], im0[o0 + single0])
        # combine items: consecutive, disjoint slot ranges that tile [0, n_pieces) and belong to that block
        cs, cn = bslot[~plain], bm[~plain] & 0xff
        assert np.array_equal(cs, np.cumsum(cn) - cn) and cn.sum() == n_pc and (cn.size == 0 or cn.min() >= 2)
        assert np.array_equal(np.repeat(p0 + np.nonzero(~plain)[0], cn), ib[o_it:o_it + n_pc])


@pytest.mark.parametrize("mesh_kind", ["sandwich", "kuhn3"])
def test_pieces_first_layout_is_bitwise_the_inline_layout(sandwich, emu, mesh_kind):
    if mesh_kind == "sandwich":
        pts, blocks, groups = sandwich
        m = orc.Material(*MAT2)
        conn, mat_id, rows = orc.element_table(groups["tet"], {1: m, 2: m, 3: m})
        variants = [dict()]
    else:
        pts, tets, tris, nodes = kuhn_box(5, 4, 4, jitter=0.2, seed=2)
        mats = {1: orc.Material(2300, 64e9, 0.1, 1., 750., 4e-6, 20.), 2: orc.Material(7800, 210e9, 0.3, 45., 460., 1.2e-5, 25.),
                3: orc.Material(2700, 70e9, 0.33, 200., 900., 2.3e-5, 30.)}
        conn, mat_id, rows = orc.element_table(tets, mats)
        variants = [dict(), dict(items_per_cta=64, elem_budget=160, max_chunk=3), dict(items_per_cta=700, elem_budget=800, max_chunk=6)]
    for kw in variants:
        S0 = build_schedule(torch.from_numpy(conn), pts.shape[0], mat_id=torch.from_numpy(mat_id), pieces_first=False,
                            bank_opt=False, pitch=False, **kw)
        S1 = build_schedule(torch.from_numpy(conn), pts.shape[0], mat_id=torch.from_numpy(mat_id), pieces_first=True,
                            tile_cost='items', bank_opt=False, pitch=False, **kw)
        check_pieces_first_invariants(S1, S0)
        S2 = build_schedule(torch.from_numpy(conn), pts.shape[0], mat_id=torch.from_numpy(mat_id), pieces_first=True, **kw)
        assert S2.n_cta <= S1.n_cta            # default tile cost of this layout: blocks per tile (the block-order pass)
        for ops, aM, aK in (("MKD", 0.2, 0.3), ("K", 0., 0.), ("MD", 0.3, 0.)):
            v0, bad0 = emu_assemble(emu, S0, pts, conn, mat_id, rows, ops, aM, aK)
            for S in (S1, S2):                 # the values depend on the chunks only, not on the tiles
                v1, bad1 = emu_assemble(emu, S, pts, conn, mat_id, rows, ops, aM, aK)
                for nm in ops:
                    assert np.array_equal(v0[nm], v1[nm]), (kw, ops, nm)
                assert np.array_equal(bad0, bad1)


def test_bank_conflict_model_of_the_incidence_loop():
    """tools/k1_model.py (the CPU model behind the K1 analysis in profiles/README.md): every incidence of the schedule is
    replayed exactly once, the conflict-free bound is loads / 16 and the modelled wavefronts lie between that bound and one
    wavefront per active lane."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import k1_model
    pts, tets, tris, nodes = kuhn_box(6, 5, 4)
    conn = np.concatenate(list(tets.values()))
    for pf in (False, True):
        S = build_schedule(torch.from_numpy(conn), pts.shape[0], pieces_first=pf)
        for with_d, loads in ((False, 7), (True, 10)):
            r = k1_model.model(S, 17 if with_d else 13, with_d)
            assert r["lane_steps"] == 16 * conn.shape[0]                    # one lane-step per (element, a, b)
            assert abs(r["ideal"] - loads * r["lane_steps"] / 16.0) < 1e-9
            assert r["ideal"] <= r["wavefronts"] <= loads * r["lane_steps"]
            skip = k1_model.model(S, 17 if with_d else 13, with_d, skip_equal_ab=True)
            assert abs(skip["ideal"] - (loads * r["lane_steps"] - 3 * 4 * conn.shape[0]) / 16.0) < 1e-9    # a == b: 4 per element
        p1 = k1_model.model_phase1(S)
        assert p1["staged"] == int(S.cta_hdr[:, 1].sum()) and p1["wavefronts"] >= 25 * p1["staged"] / 16.0


@pytest.mark.parametrize("pf", [False, True])
def test_edge_meshes_single_tet_and_fan(emu, pf):
    """A mesh of one tetrahedron (no split block at all) and a fan of 40 tets around one edge (two nodes of valence 40:
    long lists cut into many pieces, also with a tiny chunk length) in both item layouts."""
    m = orc.Material(2300, 64e9, 0.1, 1., 750., 4e-6, 20.)
    single = (np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.]]), {1: np.array([[0, 1, 2, 3]])})
    n = 40
    ang = np.linspace(0, 2 * np.pi, n, endpoint=False)
    ring = np.stack([np.cos(ang), np.sin(ang), np.zeros(n)], 1)
    fan = (np.vstack([[0, 0, -1.], [0, 0, 1.], ring]), {1: np.array([[0, 2 + k, 2 + (k + 1) % n, 1] for k in range(n)])})
    for pts, tg in (single, fan):
        conn, mat_id, rows = orc.element_table(tg, {1: m})
        ops = orc.assemble(pts, conn, rows, mat_id, 0.1, 0.2)
        for chunk in (8, 2):
            S = build_schedule(torch.from_numpy(conn), pts.shape[0], mat_id=torch.from_numpy(mat_id), pieces_first=pf, max_chunk=chunk)
            vals, bad = emu_assemble(emu, S, pts, conn, mat_id, rows, "MKD", 0.1, 0.2)
            assert bad.min() == 2 ** 31 - 1
            for nm, pm in (("K", layout.plane_map_K()), ("D", layout.plane_map_D(0.1, 0.2))):
                A = planes_to_csc(vals[nm], pm, S, 4 * pts.shape[0])
                parity.check_pattern(A, None, ops[nm])
                assert parity.scaled_error(A, ops[nm]) < parity.VALUE_TOL


@pytest.mark.parametrize("seed", [0, 3])
def test_random_delaunay_meshes_both_layouts(emu, seed):
    """Unstructured meshes with irregular valences (up to ~60 elements per node: many pieces per diagonal block) and a
    material interface, random tile parameters: the two item layouts give bitwise the same planes and match the oracle."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    pts = rng.random((int(rng.integers(150, 450)), 3)) * np.array([2.0, 0.5, 0.3])
    tets = Delaunay(pts).simplices.astype(np.int64)
    X = pts[tets]
    det = np.linalg.det(np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0], X[:, 3] - X[:, 0]], 1))
    tets[det < 0] = tets[det < 0][:, [0, 2, 1, 3]]                        # the reference's orientation convention
    tets = tets[np.abs(det) > 1e-6 * np.abs(det).max()]                   # drop slivers
    used = np.unique(tets)
    remap = -np.ones(pts.shape[0], dtype=np.int64)
    remap[used] = np.arange(used.size)
    pts, tets = pts[used], remap[tets]
    left = pts[tets].mean(axis=1)[:, 0] < 1.0
    mats = {1: orc.Material(2300, 64e9, 0.1, 1., 750., 4e-6, 20.), 2: orc.Material(7800, 210e9, 0.3, 45., 460., 1.2e-5, 25.)}
    conn, mat_id, rows = orc.element_table({1: tets[left], 2: tets[~left]}, mats)
    kw = dict(items_per_cta=int(rng.choice([64, 256, 512])), elem_budget=int(rng.choice([160, 310, 620])), max_chunk=int(rng.choice([3, 6, 8])))
    vals = {}
    for pf in (False, True):
        S = build_schedule(torch.from_numpy(conn), pts.shape[0], mat_id=torch.from_numpy(mat_id), pieces_first=pf, **kw)
        vals[pf], bad = emu_assemble(emu, S, pts, conn, mat_id, rows, "MKD", 0.2, 0.1)
        assert bad.min() == 2 ** 31 - 1
    for nm in "MKD":
        assert np.array_equal(vals[False][nm], vals[True][nm]), nm
    ops = orc.assemble(pts, conn, rows, mat_id, 0.2, 0.1)
    for nm, pm in (("K", layout.plane_map_K()), ("D", layout.plane_map_D(0.2, 0.1)), ("M", layout.plane_map_M())):
        A = planes_to_csc(vals[True][nm], pm, S, 4 * pts.shape[0])
        ref = orc.drop_component_zeros_M(ops[nm]) if nm == "M" else ops[nm]
        parity.check_pattern(A, None, ref)
        assert parity.scaled_error(A, ref) < parity.VALUE_TOL, nm


def test_schedule_fits_the_kernels_shared_memory():
    """build_schedule shrinks tiles until the largest one fits the dynamic shared memory of assemble_kernel (the launch would
    be refused otherwise); the default tiles measured on the B200 (ncu: 111.9 KB for the fused launch) are reproduced."""
    from thermoelasticity_fem_b200.symbolic import MAX_SMEM_BYTES
    pts, tets, tris, nodes = kuhn_box(14, 14, 14, 1., 1., 1., n_layers=1)
    conn = torch.from_numpy(tets[1])
    S = build_schedule(conn, pts.shape[0])
    assert 90_000 < S.smem_bytes() < 120_000 and S.smem_bytes(with_D=False) < S.smem_bytes()
    big = build_schedule(conn, pts.shape[0], items_per_cta=4000, elem_budget=5000)
    assert big.smem_bytes() <= MAX_SMEM_BYTES and big.max_cta_nodes <= 256 and big.n_cta < S.n_cta


@pytest.mark.parametrize("pf", [True, False])
def test_bank_aware_numbering_keeps_the_values_bitwise(emu, pf):
    """symbolic.optimize_banks (native host routine tefem_schedule_bank_numbering): the staged elements of every tile are
    renumbered inside blocks of 16, the incidence entries follow; the schedule still decodes to the mesh, the emulated
    assembly is BITWISE unchanged, and the modelled shared-memory wavefronts of the incidence loop drop."""
    from thermoelasticity_fem_b200.symbolic import optimize_banks
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import k1_model
    pts, tets, tris, nodes = kuhn_box(9, 8, 7, jitter=0.15, seed=5)
    mats = {1: orc.Material(2300, 64e9, 0.1, 1., 750., 4e-6, 20.), 2: orc.Material(7800, 210e9, 0.3, 45., 460., 1.2e-5, 25.),
            3: orc.Material(2700, 70e9, 0.33, 200., 900., 2.3e-5, 30.)}
    conn, mat_id, rows = orc.element_table(tets, mats)
    S = build_schedule(torch.from_numpy(conn), pts.shape[0], mat_id=torch.from_numpy(mat_id), pieces_first=pf, bank_opt=False)
    S2, (before, after) = optimize_banks(S, return_cost=True, n_threads=2)
    assert S2.bank_optimized and not S.bank_optimized and after < 0.8 * before
    H, H2 = S.cta_hdr.numpy().astype(np.int64), S2.cta_hdr.numpy().astype(np.int64)
    assert np.array_equal(H, H2) and S2.max_cta_elems == S.max_cta_elems            # a bijection: no holes, same segment sizes
    assert torch.equal(S.item_inc, S2.item_inc) and torch.equal(S.item_meta, S2.item_meta)
    ce, ce2 = S.cta_elems.numpy(), S2.cta_elems.numpy()
    cn8 = S2.cta_conn8.numpy().astype(np.int64) & 0xFFFFFFFF
    inc, inc2 = S.inc.numpy().astype(np.int64) & 0xFFFF, S2.inc.numpy().astype(np.int64) & 0xFFFF
    cnodes = S2.cta_nodes.numpy()
    moved = 0
    for c in range(S.n_cta):
        n_el, n_inc, n_nd, o_el, o_inc, o_nd = H[c, 1], H[c, 2], H[c, 4], H[c, 6], H[c, 7], H[c, 9]
        a, b = ce[o_el:o_el + n_el], ce2[o_el:o_el + n_el]
        for b0 in range(0, n_el, 16):                                               # a permutation inside every block of 16
            assert np.array_equal(np.sort(a[b0:b0 + 16]), np.sort(b[b0:b0 + 16]))
        moved += int((a != b).sum())
        loc = np.stack([(cn8[o_el:o_el + n_el] >> (8 * k)) & 0xff for k in range(4)], axis=1)
        assert np.array_equal(cnodes[o_nd:o_nd + n_nd][loc], conn[b])               # byte connectivity follows the elements
        w, w2 = inc[o_inc:o_inc + n_inc], inc2[o_inc:o_inc + n_inc]
        assert np.array_equal(w & 15, w2 & 15) and np.array_equal(a[w >> 4], b[w2 >> 4])   # same (element, a, b), new local number
    assert moved > 0
    v1, bad1 = emu_assemble(emu, S, pts, conn, mat_id, rows, "MKD", 0.2, 0.1)
    v2, bad2 = emu_assemble(emu, S2, pts, conn, mat_id, rows, "MKD", 0.2, 0.1)
    for nm in "MKD":
        assert np.array_equal(v1[nm], v2[nm]), nm
    for with_d in (False, True):
        r1, r2 = k1_model.model(S, 17 if with_d else 13, with_d, True), k1_model.model(S2, 17 if with_d else 13, with_d, True)
        assert r2["wavefronts"] < 0.93 * r1["wavefronts"] and r2["ideal"] == r1["ideal"]
    assert k1_model.model_phase1(S2)["wavefronts"] == k1_model.model_phase1(S)["wavefronts"]   # phase 1 is untouched



@pytest.mark.parametrize("pf", [True, False])
def test_fixed_pitch_lists_keep_the_values_bitwise(emu, pf):
    """symbolic.pitch_lists: every list is moved to a multiple of the pitch (6 entries for blocks, 10 for pieces), entry by
    entry unchanged; the emulated assembly is bitwise the same and the start words of 32 consecutive items hit 32 different
    four-byte banks."""
    pts, tets, tris, nodes = kuhn_box(8, 7, 7, jitter=0.1, seed=6)
    mats = {1: orc.Material(2300, 64e9, 0.1, 1., 750., 4e-6, 20.), 2: orc.Material(7800, 210e9, 0.3, 45., 460., 1.2e-5, 25.),
            3: orc.Material(2700, 70e9, 0.33, 200., 900., 2.3e-5, 30.)}
    conn, mat_id, rows = orc.element_table(tets, mats)
    common = dict(mat_id=torch.from_numpy(mat_id), pieces_first=pf, bank_opt=False)
    S0 = build_schedule(torch.from_numpy(conn), pts.shape[0], pitch=False, **common)
    from thermoelasticity_fem_b200.symbolic import pitch_lists
    S1 = pitch_lists(S0, keep_occupancy=False)
    assert S1.lists_pitched and not S0.lists_pitched and S1.max_cta_inc > S0.max_cta_inc and S1.smem_bytes() <= 220 * 1024
    H0, H1 = S0.cta_hdr.numpy().astype(np.int64), S1.cta_hdr.numpy().astype(np.int64)
    keep = [c for c in range(12) if c not in (2, 7)]                    # everything but n_inc / o_inc
    assert np.array_equal(H0[:, keep], H1[:, keep]) and torch.equal(S0.item_meta, S1.item_meta)
    i0, i1 = S0.item_inc.numpy().astype(np.int64) & 0xFFFFFFFF, S1.item_inc.numpy().astype(np.int64) & 0xFFFFFFFF
    assert np.array_equal(i0 >> 16, i1 >> 16)
    inc0, inc1 = S0.inc.numpy(), S1.inc.numpy()
    meta = S0.item_meta.numpy().astype(np.int64) & 0xFFFFFFFF
    for c in range(S0.n_cta):
        n_it, o_it, n_piece = H0[c, 0], H0[c, 5], H0[c, 11]
        assert H1[c, 7] % 8 == 0 and H1[c, 2] <= S1.max_cta_inc
        for k in range(n_it):
            ln = meta[o_it + k] & 0xff
            if pf and k >= n_piece and ((meta[o_it + k] >> 8) & 0xffff) != 0:
                continue                                                 # combine item: no list
            a, b = H0[c, 7] + (i0[o_it + k] & 0xffff), H1[c, 7] + (i1[o_it + k] & 0xffff)
            assert np.array_equal(inc0[a:a + ln], inc1[b:b + ln])
            base = 0 if k < n_piece else (i1[o_it + n_piece] & 0xffff)       # the block pass starts behind the pieces' slots
            assert ((i1[o_it + k] & 0xffff) - base) % (10 if k < n_piece else 6) == 0 and (i1[o_it + k] & 0xffff) + ln <= H1[c, 2]
        lo = n_piece if pf else 0
        sl = slice(o_it + lo, o_it + min(n_it, lo + 32))                   # the first warp of the block pass
        starts, lens = i1[sl] & 0xffff, meta[sl] & 0xff
        gathers = ~(((meta[sl] >> 8) & 0xffff) != 0) if pf else np.ones(lens.size, dtype=bool)     # combine items load nothing
        if np.all(lens[gathers] <= 6):                                     # one slot each: consecutive multiples of 3 words
            assert np.unique((starts[gathers] >> 1) % 32).size == int(gathers.sum())
    v0, _ = emu_assemble(emu, S0, pts, conn, mat_id, rows, "MKD", 0.2, 0.1)
    v1, _ = emu_assemble(emu, S1, pts, conn, mat_id, rows, "MKD", 0.2, 0.1)
    for nm in "MKD":
        assert np.array_equal(v0[nm], v1[nm]), nm
