import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pem_spgemm_b200 as pem
from pem_spgemm_b200 import synth
from oracle import host
owner = int(sys.argv[1]); k = int(sys.argv[2]); small = len(sys.argv) > 3
ctx = pem.Context(0)
name, tb, (rows, cols, I, J, V) = synth.config(k, small=small)
A = ctx.convert_coo(rows, cols, I, J, V)
B = ctx.convert_coo(rows, cols, I, J, V, transpose=True) if tb else A
ctx.set_option(pem.OPT_OWNER, owner)
print("start", name, owner, flush=True)
t = pem.Times()
C = ctx.spgemm(A, B, times=t)
print("done", C.info.nnz, C.info.tiles, t.step1_ms, t.step2_ms, t.step3_ms, flush=True)
if small:
    _, _, oC = host.spgemm_from_coo(rows, cols, I, J, V, tb)
    r, c, v = C.to_coo(); ro, co, vo = oC.to_coo()
    print("match", np.array_equal(r, ro), np.array_equal(c, co), np.array_equal(v, vo), flush=True)
