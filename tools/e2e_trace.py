"""Phase timing of the end-to-end loop of bench.py (host COO -> convert -> spgemm -> checksum) with the
allocator's miss counter: python tools/e2e_trace.py 4"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pem_spgemm_b200 as pem
from pem_spgemm_b200 import synth
k = int(sys.argv[1])
name, tb, (rows, cols, I, J, V) = synth.config(k)
tI, tJ, tV = (torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in (I, J, V))
ctx = pem.Context(0)
A = ctx.convert_coo(rows, cols, tI.numpy(), tJ.numpy(), tV.numpy()); B = ctx.transpose(A) if tb else A
for _ in range(4):
    C = ctx.spgemm(A, B); C.free()
for r in range(5):
    m0 = ctx.pool_mallocs; t0 = time.perf_counter()
    A2 = ctx.convert_coo(rows, cols, tI.numpy(), tJ.numpy(), tV.numpy()); ctx.sync(); t1 = time.perf_counter()
    B2 = ctx.transpose(A2) if tb else A2; ctx.sync(); t2 = time.perf_counter()
    tm = pem.Times(); C2 = ctx.spgemm(A2, B2, times=tm); t3 = time.perf_counter()
    chk = C2.checksum(); ctx.sync(); t4 = time.perf_counter()
    C2.free()
    if B2 is not A2: B2.free()
    A2.free()
    print(f"{name} e2e rep {r}: convert {1e3*(t1-t0):.2f} transpose {1e3*(t2-t1):.2f} spgemm {1e3*(t3-t2):.2f} "
          f"(steps {tm.step1_ms:.2f}/{tm.step2_ms:.2f}/{tm.step3_ms:.2f}) checksum {1e3*(t4-t3):.2f} total {1e3*(t4-t0):.2f} ms | "
          f"pool mallocs +{ctx.pool_mallocs - m0} pool {ctx.pool_bytes / 2**30:.2f} GiB", flush=True)
ctx.close()
