"""TEST INFRASTRUCTURE — numpy restatement of the reference's tiled-CSR data contract
and of the intermediate arrays of its three SpGEMM steps.

Every function cites the reference site it restates (paths under /root/reference).
Vectorised numpy only; sizes are meant for test inputs (up to a few million nonzeros).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

TS = 16  # tile size, spgemm.cu:727


@dataclass
class Tiled:
    """Arrays of section 2.2 for one matrix."""
    rows: int
    cols: int
    tile_rows: int          # spgemm.cu:840-843
    tile_cols: int
    cnt: int                # number of non-empty tiles
    tile_row: np.ndarray    # int32[cnt]   high half of the sorted key, spgemm.cu:131-134
    tile_col: np.ndarray    # int32[cnt]   low half  (== *_tileColIdx, spgemm.cu:1001-1006)
    tile_nnz_ptr: np.ndarray  # int64[cnt+1] == *_perTileNnz (exclusive scan), spgemm.cu:866-878
    vals: np.ndarray        # float64[nnz] tile-major, row-major inside a tile, spgemm.cu:211-222
    row_col_idx: np.ndarray  # uint8[nnz] (r<<4)|c, spgemm.cu:195,221
    masks: np.ndarray       # uint16[cnt,16] bit c of masks[t,r]  <=> (r,c) present, spgemm.cu:196-200
    row_ptr: np.ndarray     # uint8[cnt,16] exclusive scan of row popcounts, spgemm.cu:205-209
    masks_t: np.ndarray     # uint16[cnt,16] bit r of masks_t[t,c] <=> (r,c) present, spgemm.cu:228-258
    tile_row_ptr: np.ndarray  # int32[tile_rows+1] CSR of the tile-structure matrix, spgemm.cu:985-999


def _popcount16(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint32)
    x = x - ((x >> 1) & 0x5555)
    x = (x & 0x3333) + ((x >> 2) & 0x3333)
    x = (x + (x >> 4)) & 0x0F0F
    return ((x + (x >> 8)) & 0x1F).astype(np.int32)


def tile_format(rows: int, cols: int, I, J, V, transpose: bool = False) -> Tiled:
    """COO -> tiled CSR, restating spgemm.cu:112-258 and :840-1031.

    ``transpose`` swaps the roles of I/J and rows/cols first, as the CLI does for B in
    the A*A^T mode (spgemm.cu:788-792)."""
    I = np.asarray(I, np.int64); J = np.asarray(J, np.int64); V = np.asarray(V, np.float64)
    if transpose:
        I, J, rows, cols = J, I, cols, rows
    tile_rows = (rows - 1 + TS) // TS
    tile_cols = (cols - 1 + TS) // TS
    # decide_which_tile (spgemm.cu:112-135): key = (I>>4)<<32 | (J>>4); then COO is sorted by
    # (I,J) (spgemm.cu:894-897) and each tile picks its entries row by row (spgemm.cu:180-222):
    # overall order = (tile key, r, c).
    full = ((I >> 4) << 36) | ((J >> 4) << 8) | ((I & 15) << 4) | (J & 15)
    order = np.argsort(full, kind="stable")
    full = full[order]
    vals = V[order]
    tkey = full >> 8
    rc = (full & 0xFF).astype(np.uint8)
    if full.size:
        head = np.concatenate([[True], tkey[1:] != tkey[:-1]])
    else:
        head = np.zeros(0, bool)
    starts = np.flatnonzero(head)
    cnt = int(starts.size)
    tile_nnz_ptr = np.concatenate([starts, [full.size]]).astype(np.int64)
    tile_row = (tkey[starts] >> 28).astype(np.int32)
    tile_col = (tkey[starts] & ((1 << 28) - 1)).astype(np.int32)
    tid = np.cumsum(head) - 1 if full.size else np.zeros(0, np.int64)
    r = (rc >> 4).astype(np.int64); c = (rc & 15).astype(np.int64)
    masks = np.zeros((cnt, TS), np.uint16)
    masks_t = np.zeros((cnt, TS), np.uint16)
    np.bitwise_or.at(masks, (tid, r), (1 << c).astype(np.uint16))      # ballot, spgemm.cu:196
    np.bitwise_or.at(masks_t, (tid, c), (1 << r).astype(np.uint16))    # spgemm.cu:243-253
    pc = _popcount16(masks)
    row_ptr = (np.cumsum(pc, axis=1) - pc).astype(np.uint8)            # spgemm.cu:205-209
    tile_row_ptr = np.zeros(tile_rows + 1, np.int64)
    np.add.at(tile_row_ptr, tile_row.astype(np.int64) + 1, 1)          # spgemm.cu:990-999
    tile_row_ptr = np.cumsum(tile_row_ptr).astype(np.int32)
    return Tiled(rows, cols, tile_rows, tile_cols, cnt, tile_row, tile_col, tile_nnz_ptr, vals, rc,
                 masks, row_ptr, masks_t, tile_row_ptr)


@dataclass
class TiledProduct:
    """Intermediate arrays of steps 1-2 for C' = A'*B' (tile level) and C's masks."""
    c_row_ptr: np.ndarray     # int64[tile_rows_A+1]  _C_rowPtr, spgemm.cu:1168
    c_tile_row: np.ndarray    # int32[n_ctiles]       _C_tileRowIdx, spgemm.cu:378
    c_tile_col: np.ndarray    # int32[n_ctiles]       _C_tileColIdx, ascending in a tile row, spgemm.cu:379
    pair_ptr: np.ndarray      # int64[n_ctiles+1]     pairs_insertion_offset, spgemm.cu:1242
    pairs_a: np.ndarray       # int32[n_pairs]        A tile ids, ascending k inside a C' tile, spgemm.cu:430
    pairs_b: np.ndarray       # int32[n_pairs]        B tile ids (CSR order), spgemm.cu:428-431
    c_masks: np.ndarray       # uint16[n_ctiles,16]   row masks of C tiles (Ctiles_mask, spgemm.cu:533-543)
    c_tile_nnz_ptr: np.ndarray  # int64[n_ctiles+1]   _C_perTileNnz after the scan, spgemm.cu:1288


def tiled_product(A: Tiled, B: Tiled, keep_empty: bool = True) -> TiledProduct:
    """Steps 1 and 2 at tile level.

    ``keep_empty=True`` restates the reference exactly: step 1 (spgemm.cu:271-384, or the
    NSPARSE hash path) works on the tile STRUCTURE only, so C' contains tiles whose mask
    product is empty and pairs that contribute nothing.  ``keep_empty=False`` drops pairs
    whose 16x16 boolean product is empty (and then tiles without pairs): exactly the
    non-empty tiles of C.
    """
    assert A.tile_cols == B.tile_rows
    # expansion: A tile t=(i,k) meets every B tile in B' row k
    blen = np.diff(B.tile_row_ptr.astype(np.int64))
    k = A.tile_col.astype(np.int64)
    rep = blen[k]
    ta = np.repeat(np.arange(A.cnt, dtype=np.int64), rep)
    start = np.repeat(B.tile_row_ptr[k].astype(np.int64), rep)
    within = np.arange(ta.size, dtype=np.int64) - np.repeat(np.cumsum(rep) - rep, rep)
    tb = start + within
    if not keep_empty:
        a_col_occ = np.bitwise_or.reduce(A.masks, axis=1)           # columns k present in the A tile
        b_row_occ = np.bitwise_or.reduce(B.masks_t, axis=1)         # rows k present in the B tile
        ok = (a_col_occ[ta] & b_row_occ[tb]) != 0
        ta, tb = ta[ok], tb[ok]
    ci = A.tile_row[ta].astype(np.int64)
    cj = B.tile_col[tb].astype(np.int64)
    order = np.lexsort((ta, cj, ci))      # (i, j) then ascending A tile id == ascending k
    ta, tb, ci, cj = ta[order], tb[order], ci[order], cj[order]
    key = ci * np.int64(max(B.tile_cols, 1)) + cj
    head = np.concatenate([[True], key[1:] != key[:-1]]) if key.size else np.zeros(0, bool)
    starts = np.flatnonzero(head)
    n_ct = int(starts.size)
    pair_ptr = np.concatenate([starts, [key.size]]).astype(np.int64)
    c_tile_row = ci[starts].astype(np.int32)
    c_tile_col = cj[starts].astype(np.int32)
    c_row_ptr = np.zeros(A.tile_rows + 1, np.int64)
    np.add.at(c_row_ptr, c_tile_row.astype(np.int64) + 1, 1)
    c_row_ptr = np.cumsum(c_row_ptr)
    # masks: Cmask[r] |= OR_{k in Amask[r]} Bmask[k]  (equivalent to the AND test against the
    # transposed B mask in spgemm.cu:533-540)
    pm = np.zeros((ta.size, TS), np.uint16)
    am = A.masks[ta]; bm = B.masks[tb]
    for kk in range(TS):
        sel = ((am >> kk) & 1).astype(bool)            # [pairs,16]  A has (r,kk)
        pm |= np.where(sel, bm[:, kk][:, None], 0).astype(np.uint16)
    c_masks = np.zeros((n_ct, TS), np.uint16)
    if ta.size:
        c_masks = np.bitwise_or.reduceat(pm, starts, axis=0)
    nnz_t = _popcount16(c_masks).sum(axis=1).astype(np.int64)
    c_tile_nnz_ptr = np.concatenate([[0], np.cumsum(nnz_t)]).astype(np.int64)
    return TiledProduct(c_row_ptr, c_tile_row, c_tile_col, pair_ptr, ta.astype(np.int32),
                        tb.astype(np.int32), c_masks, c_tile_nnz_ptr)


def product_to_coo(P: TiledProduct):
    """Coordinates of C in tile order (tile, r, c) from the C masks — what sanitize_C
    (spgemm.cu:663-695) emits before the final sort."""
    t, r = np.nonzero(P.c_masks)
    rows, cols = [], []
    for c in range(TS):
        sel = ((P.c_masks[t, r] >> c) & 1).astype(bool)
        rows.append((P.c_tile_row[t[sel]].astype(np.int64) << 4) + r[sel])
        cols.append((P.c_tile_col[t[sel]].astype(np.int64) << 4) + c)
    rows = np.concatenate(rows) if rows else np.zeros(0, np.int64)
    cols = np.concatenate(cols) if cols else np.zeros(0, np.int64)
    o = np.lexsort((cols, rows))
    return rows[o].astype(np.int32), cols[o].astype(np.int32)
