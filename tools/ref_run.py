"""Run a rebuild of the reference (oracle/_ref/<bin>) through its own CLI on the named inputs and print
its report tail and CSV row.
    python tools/ref_run.py [--bin pemspgemm_ref61] [--timeout 300] lap600 1 2 3 4
Inputs: a BASELINE config number, or lap<g> for a g x g 2-D Laplacian (lap600 is the smallest shape
class that takes the reference's NSPARSE step 1, spgemm.cu:1142).  A run that exceeds the time-out is
reported as such (the compute_100 rebuild does not leave "step1 using NSPARSE" on sm_100)."""
import argparse, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pem_spgemm_b200 as pem
from pem_spgemm_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--bin", default="pemspgemm_ref")
ap.add_argument("--timeout", type=int, default=300)
ap.add_argument("inputs", nargs="+")
a = ap.parse_args()
ref = os.path.join(ROOT, "oracle", "_ref", a.bin)
work = f"/tmp/pem_refrun_{a.bin}"; os.makedirs(work, exist_ok=True)
for k in a.inputs:
    if k.startswith("lap"):
        name, tb, (rows, cols, I, J, V) = k, False, synth.laplacian2d(int(k[3:]))
    else:
        name, tb, (rows, cols, I, J, V) = synth.config(int(k))
    tile_cols = ((rows if tb else cols) + 15) // 16
    path = "NSPARSE" if tile_cols > 16384 else "SPA"
    mtx = os.path.join(work, name + ".mtx")
    pem.mtx_write(mtx, rows, cols, I, J, V)
    csv = os.path.join(work, "pemspgemm_benchmark_result.csv")
    if os.path.exists(csv): os.remove(csv)
    t0 = time.time()
    try:
        p = subprocess.run([ref, mtx, "0"] + (["1"] if tb else []), cwd=work, capture_output=True, text=True, timeout=a.timeout)
        print(f"[{a.bin}] {k} {name} ({path} step 1, {tile_cols} B tile columns): rc {p.returncode} wall {time.time()-t0:.1f}s")
        print("  ", "\n   ".join(p.stdout.strip().splitlines()[-22:]))
        if os.path.exists(csv): print("   CSV:", open(csv).read().strip())
        if p.returncode != 0: print("   STDERR:", p.stderr[-500:])
    except subprocess.TimeoutExpired as e:
        tail = (e.stdout or b"")[-300:]
        print(f"[{a.bin}] {k} {name} ({path} step 1): TIMEOUT after {time.time()-t0:.0f}s; stdout tail: {tail!r}")
    sys.stdout.flush()
