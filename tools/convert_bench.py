"""Conversion timing loop (host COO in pinned memory -> tiled CSR): python tools/convert_bench.py 4 [--reps 4]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pem_spgemm_b200 as pem
from pem_spgemm_b200 import synth
ap = argparse.ArgumentParser(); ap.add_argument("config", type=int); ap.add_argument("--reps", type=int, default=4)
a = ap.parse_args()
name, tb, (rows, cols, I, J, V) = synth.config(a.config)
tI = torch.from_numpy(np.ascontiguousarray(I)).pin_memory(); tJ = torch.from_numpy(np.ascontiguousarray(J)).pin_memory()
tV = torch.from_numpy(np.ascontiguousarray(V)).pin_memory()
ctx = pem.Context(0)
for r in range(a.reps):
    t = pem.Times(); t0 = time.perf_counter()
    A = ctx.convert_coo(rows, cols, tI.numpy(), tJ.numpy(), tV.numpy(), times=t)
    ctx.sync(); w = (time.perf_counter() - t0) * 1e3
    print(f"{name} rep {r}: convert wall {w:.2f} ms (library total {t.convert_total_ms:.2f}, tile kernel {t.convert_kernel_ms:.3f}) "
          f"= {(I.nbytes + J.nbytes + V.nbytes) / w / 1e6:.1f} GB/s of COO", flush=True)
    A.free()
dI, dJ, dV = (x.cuda() for x in (tI, tJ, tV)); torch.cuda.synchronize()
for r in range(2):
    t0 = time.perf_counter()
    A = ctx.convert_coo(rows, cols, dI.data_ptr(), dJ.data_ptr(), dV.data_ptr(), nnz=I.size)
    ctx.sync(); w = (time.perf_counter() - t0) * 1e3
    print(f"{name} device input rep {r}: convert wall {w:.2f} ms", flush=True)
    A.free()
ctx.close()
