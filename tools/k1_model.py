"""CPU model of the shared-memory traffic of K1's incidence loop (no GPU needed).

For a schedule built by symbolic.build_schedule it replays the thread mapping of assemble_kernel (item k of a pass ->
thread k % 256, warps of 32 consecutive threads, every warp iterates to the longest list it holds) and counts, per
half-warp and per 8-byte load, the shared-memory wavefronts: the number of DISTINCT addresses that fall into the most
loaded of the 16 eight-byte banks (equal addresses are a broadcast).  Output: wavefronts per element of the gathers,
lane fill, and the conflict-free bound - the quantities profiles/README.md compares with ncu's
l1tex__data_pipe_lsu_wavefronts_mem_shared.

usage: python tools/k1_model.py [cells] [variant ...]   variant = items,budget,chunk[,pieces_first[,bank_opt[,pitch]]]
(bank_opt: 0 = staged order, 1 = symbolic.optimize_banks for both record strides, 2 = stride 17 only, 3 = stride 13 only;
pitch: 1 = symbolic.pitch_lists, 0 = lists packed back to back)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from thermoelasticity_fem_b200.symbolic import build_schedule  # noqa: E402
from thermoelasticity_fem_b200.synthetic import kuhn_box  # noqa: E402

THREADS = 256


def _max_bank_load(addr, active):
    """addr [n_halfwarps, 16] (double index), active same shape (bool) -> wavefronts per half-warp."""
    n = addr.shape[0]
    a = np.where(active, addr, -1 - np.arange(16)[None, :])          # inactive lanes: unique negatives (ignored below)
    a_sorted = np.sort(a, axis=1)
    first = np.ones_like(a_sorted, dtype=bool)
    first[:, 1:] = a_sorted[:, 1:] != a_sorted[:, :-1]
    valid = first & (a_sorted >= 0)
    bank = np.where(valid, a_sorted % 16, 16)
    load = np.zeros((n, 17), dtype=np.int64)
    np.add.at(load, (np.repeat(np.arange(n), 16), bank.reshape(-1)), 1)
    return load[:, :16].max(axis=1)


def model(S, rec_stride=13, with_d=False, skip_equal_ab=False, tiles=None):
    """Returns dict(wavefronts, ideal, lane_steps, warp_steps, blocks) summed over the tiles (default: all) of the
    schedule (incidence loop only)."""
    H = S.cta_hdr.numpy().astype(np.int64)
    ii = S.item_inc.numpy().astype(np.int64) & 0xFFFFFFFF
    im = S.item_meta.numpy().astype(np.int64) & 0xFFFFFFFF
    inc = S.inc.numpy().astype(np.int64) & 0xFFFF
    tot_w = tot_ideal = lane_steps = warp_steps = n_blocks = 0
    for c in (range(S.n_cta) if tiles is None else tiles):
        n_it, o_it, o_inc = H[c, 0], H[c, 5], H[c, 7]
        n_blocks += (H[c + 1, 10] if c + 1 < S.n_cta else S.n_blocks) - H[c, 10]
        n_piece = H[c, 11]
        start = ii[o_it:o_it + n_it] & 0xffff
        ln = im[o_it:o_it + n_it] & 0xff
        combine = np.zeros(n_it, dtype=bool)
        if n_piece > 0:                         # pieces-first layout: block items that only add parked pieces gather nothing
            slot = (im[o_it:o_it + n_it] >> 8) & 0xffff
            combine[n_piece:] = slot[n_piece:] != 0
            ln = np.where(combine, 0, ln)
        # thread mapping: two item ranges (pieces, then block items) each mapped from thread 0
        ranges = [(0, n_piece), (n_piece, n_it)] if n_piece > 0 else [(0, n_it)]
        for lo, hi in ranges:
            n = hi - lo
            if n <= 0:
                continue
            npad = -(-n // 32) * 32
            st = np.zeros(npad, dtype=np.int64)
            le = np.zeros(npad, dtype=np.int64)
            st[:n], le[:n] = start[lo:hi], ln[lo:hi]
            st, le = st.reshape(-1, 32), le.reshape(-1, 32)
            wmax = le.max(axis=1)
            for k in range(int(wmax.max()) if wmax.size else 0):
                act_w = wmax > k                                   # warps still iterating
                if not act_w.any():
                    break
                act = (le > k) & act_w[:, None]
                w = np.where(act, inc[np.minimum(o_inc + st + k, inc.size - 1)], 0)
                e, a, b = w >> 4, (w >> 2) & 3, w & 3
                base = e * rec_stride
                loads = [base + 3 * a + i for i in range(3)] + [base + 12]
                masks = [act] * 4
                nb_act = act & (a != b) if skip_equal_ab else act
                loads += [base + 3 * b + i for i in range(3)]
                masks += [nb_act] * 3
                if with_d:                                         # GT[c][b]: b == 0 -> 13 + c, else 3 (c + 1) + b - 1
                    for cc in range(3):
                        loads.append(np.where(b == 0, base + 13 + cc, base + 3 * (cc + 1) + b - 1))
                        masks.append(act)
                sel = act_w
                for ad, mk in zip(loads, masks):
                    hw_a = ad[sel].reshape(-1, 16)
                    hw_m = mk[sel].reshape(-1, 16)
                    tot_w += int(_max_bank_load(hw_a, hw_m).sum())
                    tot_ideal += int(hw_m.sum())                   # 1/16 wavefront per active lane
                lane_steps += int(act.sum())
                warp_steps += int(act_w.sum())
    return dict(wavefronts=tot_w, ideal=tot_ideal / 16.0, lane_steps=lane_steps, warp_steps=warp_steps, blocks=n_blocks)


def model_entry_loads(S, tiles=None):
    """Shared-memory wavefronts of the 2-byte incidence-entry loads (lane l reads inc[start_l + step]; 32 lanes, 32
    four-byte banks, equal words are a broadcast)."""
    H = S.cta_hdr.numpy().astype(np.int64)
    ii = S.item_inc.numpy().astype(np.int64) & 0xFFFFFFFF
    im = S.item_meta.numpy().astype(np.int64) & 0xFFFFFFFF
    tot = 0
    for c in (range(S.n_cta) if tiles is None else tiles):
        n_it, o_it, n_piece = H[c, 0], H[c, 5], H[c, 11]
        start = ii[o_it:o_it + n_it] & 0xffff
        ln = im[o_it:o_it + n_it] & 0xff
        if n_piece > 0:
            slot = (im[o_it:o_it + n_it] >> 8) & 0xffff
            combine = np.zeros(n_it, dtype=bool)
            combine[n_piece:] = slot[n_piece:] != 0
            ln = np.where(combine, 0, ln)
        for lo, hi in ([(0, n_piece), (n_piece, n_it)] if n_piece > 0 else [(0, n_it)]):
            n = hi - lo
            if n <= 0:
                continue
            npad = -(-n // 32) * 32
            st = np.zeros(npad, dtype=np.int64)
            le = np.zeros(npad, dtype=np.int64)
            st[:n], le[:n] = start[lo:hi], ln[lo:hi]
            st, le = st.reshape(-1, 32), le.reshape(-1, 32)
            wmax = le.max(axis=1)
            for k in range(int(wmax.max())):
                aw = wmax > k
                act = (le > k) & aw[:, None]
                word = np.where(act, (st + k) >> 1, -1 - np.arange(32)[None, :])[aw]
                srt = np.sort(word, axis=1)
                first = np.ones_like(srt, dtype=bool)
                first[:, 1:] = srt[:, 1:] != srt[:, :-1]
                bank = np.where(first & (srt >= 0), srt % 32, 32)
                load = np.zeros((srt.shape[0], 33), dtype=np.int64)
                np.add.at(load, (np.repeat(np.arange(srt.shape[0]), 32), bank.reshape(-1)), 1)
                tot += int(load[:, :32].max(axis=1).sum())
    return tot


def model_phase1(S, rec_stride=13, tiles=None):
    """Shared-memory wavefronts of phase 1 (thread le -> staged element le): 12 coordinate loads gathered by tile-local
    node (xyz[3 * node + c]) and rec_stride record stores at le * rec_stride + k."""
    H = S.cta_hdr.numpy().astype(np.int64)
    cn8 = S.cta_conn8.numpy().astype(np.int64) & 0xFFFFFFFF
    tot = staged = 0
    for c in (range(S.n_cta) if tiles is None else tiles):
        n_el, o_el = H[c, 1], H[c, 6]
        npad = -(-n_el // 16) * 16
        act = np.arange(npad) < n_el
        w = np.zeros(npad, dtype=np.int64)
        w[:n_el] = cn8[o_el:o_el + n_el]
        le = np.arange(npad)
        for a in range(4):
            node = (w >> (8 * a)) & 0xff
            for cc in range(3):
                tot += int(_max_bank_load((3 * node + cc).reshape(-1, 16), act.reshape(-1, 16)).sum())
        for k in range(rec_stride):
            tot += int(_max_bank_load((le * rec_stride + k).reshape(-1, 16), act.reshape(-1, 16)).sum())
        staged += n_el
    return dict(wavefronts=tot, staged=staged)


if __name__ == "__main__":
    cells = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    variants = sys.argv[2:] or ["256,310,8"]
    pts, tets, tris, nodes = kuhn_box(cells, cells, cells, 1.0, 1.0, 1.0, n_layers=1, index_dtype=np.int32)
    conn = torch.from_numpy(tets[1].astype(np.int64))
    E, N = conn.shape[0], pts.shape[0]
    for v in variants:
        f = [int(x) for x in v.split(",")]
        kw = dict(items_per_cta=f[0], elem_budget=f[1], max_chunk=f[2])
        if len(f) > 3:
            kw["pieces_first"] = bool(f[3])
        S = build_schedule(conn, N, bank_opt=False, pitch=bool(f[5]) if len(f) > 5 else False, **kw)
        if len(f) > 4 and f[4]:
            from thermoelasticity_fem_b200.symbolic import optimize_banks
            S = optimize_banks(S, strides={1: 3, 2: 2, 3: 1}[f[4]])
        staged = int(S.cta_hdr[:, 1].sum())
        tiles = range(S.n_cta // 2, min(S.n_cta, S.n_cta // 2 + 40))
        nb_t = sum(int((S.cta_hdr[c + 1, 10] if c + 1 < S.n_cta else S.n_blocks) - S.cta_hdr[c, 10]) for c in tiles)
        print(f"{v:14s} incidence-entry loads: {model_entry_loads(S, tiles) / (nb_t / 2.5):.2f} wavefronts/elem "
              f"(max entries per tile {S.max_cta_inc})", flush=True)
        for rs in (13, 17):
            r1 = model_phase1(S, rs, tiles)
            nb = sum(int((S.cta_hdr[c + 1, 10] if c + 1 < S.n_cta else S.n_blocks) - S.cta_hdr[c, 10]) for c in tiles)
            print(f"{v:14s} phase 1 (record stride {rs}): {r1['wavefronts'] / (nb / 2.5):.2f} wavefronts/elem, "
                  f"{r1['wavefronts'] / r1['staged']:.2f} per staged element, staged/elem={r1['staged'] / (nb / 2.5):.2f}", flush=True)
        for with_d in (False, True):
            for skip in (False, True):
                # a sample of tiles from the middle of the mesh (interior rows: 15 blocks per row = 2.5 per element)
                tiles = range(S.n_cta // 2, min(S.n_cta, S.n_cta // 2 + 40))
                r = model(S, 17 if with_d else 13, with_d, skip, tiles)
                el = r['blocks'] / 2.5
                print(f"{v:14s} {'MDK' if with_d else 'K  '} skip_ab={int(skip)} n_cta={S.n_cta} staged/E={staged / E:.2f} "
                      f"gather wavefronts/elem={r['wavefronts'] / el:.2f} (conflict-free {r['ideal'] / el:.2f}) "
                      f"fill={r['lane_steps'] / (32 * r['warp_steps']):.3f} lane-steps/elem={r['lane_steps'] / el:.2f}", flush=True)
