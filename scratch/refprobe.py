import os, subprocess, sys, time
sys.path.insert(0, os.getcwd())
import pem_spgemm_b200 as pem
from pem_spgemm_b200 import synth
ref = os.path.join(os.getcwd(), "oracle/_ref/pemspgemm_ref")
os.makedirs("/tmp/rp", exist_ok=True)
for name, gen, tb in [("lap600", lambda: synth.laplacian2d(600), False), ("webbase", lambda: synth.config(2)[2], False)]:
    rows, cols, I, J, V = gen()
    mtx = f"/tmp/rp/{name}.mtx"
    pem.mtx_write(mtx, rows, cols, I, J, V)
    t0 = time.time()
    p = subprocess.run([ref, mtx, "0"] + (["1"] if tb else []), cwd="/tmp/rp", capture_output=True, text=True, timeout=900)
    print(name, "rc", p.returncode, "wall %.1f" % (time.time() - t0))
    print(p.stdout[-1800:]); print("STDERR:", p.stderr[-1500:], flush=True)
print(open("/tmp/rp/pemspgemm_benchmark_result.csv").read() if os.path.exists("/tmp/rp/pemspgemm_benchmark_result.csv") else "no csv")
