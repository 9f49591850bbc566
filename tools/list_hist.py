"""Developer probe: histogram of pair-list lengths and per-row stats after step 1."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pem_spgemm_b200 as pem
from pem_spgemm_b200 import synth
ctx = pem.Context(0)
for k in [int(x) for x in sys.argv[1:]]:
    name, tb, (rows, cols, I, J, V) = synth.config(k)
    A = ctx.convert_coo(rows, cols, I, J, V)
    B = ctx.convert_coo(rows, cols, I, J, V, transpose=True) if tb else A
    C = ctx.step1(A, B)
    pp = C.array("pair_ptr"); ln = np.diff(pp)
    rp = C.array("row_ptr"); D = np.diff(rp)
    La = np.diff(A.array("tile_row_ptr"))
    print(name, "tiles", ln.size, "pairs", int(pp[-1]))
    for lo, hi in [(1, 1), (2, 2), (3, 8), (9, 24), (25, 64), (65, 256), (257, 1024), (1025, 10**9)]:
        m = (ln >= lo) & (ln <= hi)
        print(f"   len {lo}-{hi}: lists {int(m.sum())} pairs {int(ln[m].sum())}")
    long_per_row = np.add.reduceat((ln > 24).astype(np.int64), rp[:-1][D > 0]) if ln.size else []
    print("   rows with long lists", int((np.asarray(long_per_row) > 0).sum()), "max long lists in a row", int(np.max(long_per_row)) if len(long_per_row) else 0)
    print("   D: max", int(D.max()), "mean", float(D.mean()), " La: max", int(La.max()))
    C.free()
    if B is not A: B.free()
    A.free()
