"""Host-side model of warp trip counts for the dense-tile kernels (numpy, no GPU): what a lane mapping or loop
shape would cost on a cage15-shaped stencil product BEFORE it is written.  The figures quoted in
profiles/r02_step3_windows.md come from `python tools/lane_sim.py 40` (grid 40 x 41 x 41; the per-warp ratios
match ncu's source counters of the full-size product: 9.65 product trips at 12.8 lanes simulated, 10.2 at 12 measured).
    python tools/lane_sim.py [grid]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pem_spgemm_b200 import synth
from oracle import tiles

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows, cols, I, J, V = synth.cage_like(n, n + 1, n + 1)
A = tiles.tile_format(rows, cols, I, J, V)
P = tiles.tiled_product(A, A, keep_empty=False)
npairs = np.diff(P.pair_ptr)
# nonzeros of C in tile order, row-major inside a tile (the window kernel's thread order)
t_idx, r_idx = np.nonzero(P.c_masks)
et, er, ec = [], [], []
for c in range(16):
    sel = ((P.c_masks[t_idx, r_idx] >> c) & 1).astype(bool)
    et.append(t_idx[sel]); er.append(r_idx[sel]); ec.append(np.full(int(sel.sum()), c))
et, er, ec = np.concatenate(et), np.concatenate(er), np.concatenate(ec)
o = np.lexsort((ec, er, et)); et, er, ec = et[o], er[o], ec[o]
E, maxnp = et.size, int(npairs.max())
pc = np.zeros((E, maxnp), np.int32)            # products of nonzero e with the j-th pair of its tile
for j in range(maxnp):
    has = npairs[et] > j
    pi = P.pair_ptr[et[has]] + j
    pc[has, j] = tiles._popcount16(A.masks[P.pairs_a[pi], er[has]] & A.masks_t[P.pairs_b[pi], ec[has]])
tot = pc.sum(1)
print(f"C tiles {npairs.size}, pairs {npairs.sum()}, nonzeros {E}, products {tot.sum()} ({tot.sum() / E:.2f} per nonzero, "
      f"{(pc > 0).sum() / npairs[et].sum():.2f} of the (nonzero, pair) combinations hit)")
print("products per nonzero, histogram:", np.bincount(tot)[:20])
win = (P.pair_ptr[:-1] // 128)[et]             # 128-pair windows


def warps(order):
    w = win[order]
    first = np.concatenate([[True], w[1:] != w[:-1]])
    ws = np.flatnonzero(first)
    pos = np.arange(E) - np.repeat(ws, np.diff(np.concatenate([ws, [E]])))
    return np.cumsum(np.concatenate([[0], ((pos[1:] % 32 == 0) | first[1:]).astype(int)]))


def report(order, label):
    wid = warps(order); nw = wid[-1] + 1
    smax = lambda x: np.maximum.reduceat(x, np.flatnonzero(np.concatenate([[True], wid[1:] != wid[:-1]])))
    p = pc[order]
    pair_trips = smax(npairs[et][order]).sum() / nw
    nested = sum(smax(p[:, j]).sum() for j in range(maxnp)) / nw     # pair loop outside, product loop inside (shipped)
    flat = smax(tot[order]).sum() / nw                                 # one flattened product loop per lane
    print(f"{label:34s} pair trips {pair_trips:5.2f} | product trips nested {nested:5.2f} (lanes {tot.sum() / nested / nw:4.1f}) "
          f"flat {flat:5.2f} (lanes {tot.sum() / flat / nw:4.1f})")


print("-- step 3, per warp of 32 nonzeros")
report(np.arange(E), "natural order")
report(np.lexsort((tot, win)), "sorted by products inside a window")
report(np.lexsort((npairs[et], win)), "sorted by pair count inside a window")
# step 2, row-mask form: lane = pair, sixteen row loops; a warp pays the largest row population per row
pa = P.pairs_a[: P.pairs_a.size // 128 * 128]
rp = tiles._popcount16(A.masks[pa]).reshape(-1, 128, 16)
cost = lambda b: b.reshape(b.shape[0], 4, 32, 16).max(axis=2).sum(axis=2).mean()
words = rp[:, :, 0::2] + rp[:, :, 1::2]
print("-- step 2 (row-mask form), loop trips per warp of 32 pairs")
print(f"sixteen row loops {cost(rp):.1f} | eight two-row loops {words.reshape(words.shape[0], 4, 32, 8).max(axis=2).sum(axis=2).mean():.1f} | "
      f"largest A tile of the warp {rp.sum(2).reshape(-1, 4, 32).max(axis=2).mean():.1f} | mean A tile {rp.sum(2).mean():.1f}")
srt = np.take_along_axis(rp, np.argsort(-rp.sum(2), axis=1, kind="stable")[:, :, None], axis=1)
print(f"pairs of a 128-pair block sorted by A-tile population: {cost(srt):.1f}")
