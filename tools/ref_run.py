"""Run the reference's sm_100 rebuild (oracle/_ref/pemspgemm_ref) on the named configs whose step 1
takes its SPA path (B tile columns <= 16384); NSPARSE-path inputs hang on sm_100.
    python tools/ref_run.py 1 3      -> prints the reference's CSV row per config"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pem_spgemm_b200 as pem
from pem_spgemm_b200 import synth
ref = os.path.join(ROOT, "oracle", "_ref", "pemspgemm_ref")
work = "/tmp/pem_refrun"; os.makedirs(work, exist_ok=True)
for k in [int(x) for x in sys.argv[1:]]:
    name, tb, (rows, cols, I, J, V) = synth.config(k)
    tile_cols = ((rows if tb else cols) + 15) // 16
    if tile_cols > 16384:
        print(f"config {k} {name}: skipped, B has {tile_cols} tile columns -> NSPARSE path (hangs on sm_100)"); continue
    mtx = os.path.join(work, name + ".mtx")
    pem.mtx_write(mtx, rows, cols, I, J, V)
    csv = os.path.join(work, "pemspgemm_benchmark_result.csv")
    if os.path.exists(csv): os.remove(csv)
    t0 = time.time()
    try:
        p = subprocess.run([ref, mtx, "0"] + (["1"] if tb else []), cwd=work, capture_output=True, text=True, timeout=420)
        print(f"config {k} {name}: rc {p.returncode} wall {time.time()-t0:.1f}s")
        print("  ", "\n   ".join(p.stdout.strip().splitlines()[-22:]))
        if os.path.exists(csv): print("   CSV:", open(csv).read().strip())
        if p.returncode != 0: print("   STDERR:", p.stderr[-500:])
    except subprocess.TimeoutExpired:
        print(f"config {k} {name}: TIMEOUT after {time.time()-t0:.0f}s")
