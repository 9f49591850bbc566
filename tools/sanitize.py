"""Small end-to-end products for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python tools/sanitize.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pem_spgemm_b200 as pem
from pem_spgemm_b200 import synth
ctx = pem.Context(0)
for k in (1, 2, 3, 4):
    name, tb, (rows, cols, I, J, V) = synth.config(k, small=True)
    if I.size > 40000:
        keep = np.arange(I.size) % 4 == 0
        I, J, V = I[keep], J[keep], V[keep]
    A = ctx.convert_coo(rows, cols, I, J, V)
    B = ctx.convert_coo(rows, cols, I, J, V, transpose=True) if tb else A
    for path in (1, 3, 4, 5):
        for owner, step2 in ((2, 1), (3, 2), (1, 0), (4, 3), (0, 0)):
            ctx.set_option(pem.OPT_STEP1_PATH, path); ctx.set_option(pem.OPT_OWNER, owner); ctx.set_option(pem.OPT_STEP2_KERNEL, step2)
            C = ctx.spgemm(A, B)
            s = C.checksum(); r, c, v = C.to_coo(); rp, cc, vv = C.to_csr(); C.free()
    print(name, "ok", r.size, flush=True)
ctx.close()
