"""Developer timing loop: per-step times of the engine on the named configs (not the bench contract).
    python tools/quick_bench.py 2 3 4 [--check] [--keep-empty] [--reps N]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pem_spgemm_b200 as pem
from pem_spgemm_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("configs", nargs="+", type=int)
ap.add_argument("--check", action="store_true")
ap.add_argument("--keep-empty", type=int, default=0)
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--small", action="store_true")
ap.add_argument("--panels", type=int, default=1, help="multiply in this many sequential tile-row panels (results freed panel by panel)")
ap.add_argument("--owner", type=int, default=0)
ap.add_argument("--step1", type=int, default=0)
ap.add_argument("--plans", type=int, default=1)
ap.add_argument("--step2", type=int, default=0)
ap.add_argument("--graphs", type=int, default=1)
ap.add_argument("--sweep", default="", help="comma list of owner:small_nnz:small_pairs variants timed on the same operands, e.g. 0:8:64,0:4:64,2:8:64")
a = ap.parse_args()
ctx = pem.Context(0)
ctx.set_option(pem.OPT_KEEP_EMPTY_TILES, a.keep_empty)
ctx.set_option(pem.OPT_OWNER, a.owner)
ctx.set_option(pem.OPT_STEP1_PATH, a.step1)
ctx.set_option(pem.OPT_STEP2_KERNEL, a.step2)
ctx.set_option(pem.OPT_SIZE_PLANS, a.plans)
ctx.set_option(pem.OPT_GRAPHS, a.graphs)
for k in a.configs:
    t0 = time.time()
    name, tb, (rows, cols, I, J, V) = synth.config(k, small=a.small)
    tg = time.time() - t0
    tc = pem.Times()
    A = ctx.convert_coo(rows, cols, I, J, V, times=tc)
    B = ctx.convert_coo(rows, cols, I, J, V, transpose=True) if tb else A
    flop = ctx.count_flop(A, B)
    best = None
    bounds = ctx.partition_panels(A, B, a.panels)
    for var in [v for v in a.sweep.split(",") if v]:
        ow, se, sp = (int(x) for x in var.split(":"))
        ctx.set_option(pem.OPT_OWNER, ow); ctx.set_option(pem.OPT_S3_SMALL_NNZ, se); ctx.set_option(pem.OPT_S3_SMALL_PAIRS, sp)
        bt = None
        for r in range(a.reps):
            t = pem.Times()
            C = ctx.spgemm(A, B, times=t)
            cs = C.checksum()
            C.free()
            if bt is None or t.step3_ms < bt.step3_ms:
                bt = t; bk = ctx.kernel_ms()
        print(f"   sweep owner {ow} small_nnz {se} small_pairs {sp}: step1 {bt.step1_ms:.3f} step2 {bt.step2_ms:.3f} step3 {bt.step3_ms:.3f} total {bt.total_ms:.3f} | numeric kernel {bk['step3_numeric']:.3f} | checksum {cs[0].hex()} {cs[1].hex()}", flush=True)
    ctx.set_option(pem.OPT_OWNER, a.owner); ctx.set_option(pem.OPT_S3_SMALL_NNZ, 8); ctx.set_option(pem.OPT_S3_SMALL_PAIRS, 64)
    for r in range(a.reps):
        t = pem.Times()
        tot = dict(tiles=0, pairs=0, nnz=0, tile_products=0)
        for pn in range(a.panels):
            tp = pem.Times()
            C = ctx.spgemm(A, B, times=tp, panel=(int(bounds[pn]), int(bounds[pn + 1])) if a.panels > 1 else None)
            for f in ("step1_ms", "step2_ms", "step3_ms", "total_ms"):
                setattr(t, f, getattr(t, f) + getattr(tp, f))
            i = C.info
            for f in tot:
                tot[f] += getattr(i, f)
            if pn + 1 < a.panels:
                C.free()
        class info: pass
        for f in tot:
            setattr(info, f, tot[f])
        if best is None or t.total_ms < best.total_ms:
            best = t
        if os.environ.get('PEM_QB_VERBOSE'):
            print(f'   rep {r}: step1 {t.step1_ms:.3f} step2 {t.step2_ms:.3f} step3 {t.step3_ms:.3f} total {t.total_ms:.3f} pool {ctx.pool_bytes/2**30:.2f} GiB'
                  f' | last panel kernels {" ".join(f"{k}={v:.3f}" for k, v in ctx.kernel_ms().items())}', flush=True)
        if r + 1 < a.reps:
            C.free()
    print(f"{name}: gen {tg:.1f}s nnzA {A.info.nnz} tilesA {A.info.tiles} flop {flop} | convert {tc.convert_total_ms:.2f} ms "
          f"(tile kernel {tc.convert_kernel_ms:.3f}) | C tiles {info.tiles} pairs {info.pairs} nnz {info.nnz} tile_products {info.tile_products} | "
          f"step1 {best.step1_ms:.3f} step2 {best.step2_ms:.3f} step3 {best.step3_ms:.3f} total {best.total_ms:.3f} ms "
          f"=> {2*flop/best.total_ms/1e6:.1f} GFLOP/s | pool {ctx.pool_bytes/2**30:.2f} GiB | graph launches {ctx.graph_replays}", flush=True)
    if a.check:
        from oracle import host
        oA, oB, oC = host.spgemm_from_coo(rows, cols, I, J, V, tb)
        s, ab = C.checksum()
        ok = (oC.nnz == info.nnz) and host.flop(oA, oB) == flop
        r, c, v = C.to_coo()
        ro, co, vo = oC.to_coo()
        ok = ok and np.array_equal(r, ro) and np.array_equal(c, co)
        exact = ok and np.array_equal(v, vo)
        print(f"   check vs oracle: structure {'OK' if ok else 'MISMATCH'}, values {'bit-exact' if exact else 'DIFFER'}", flush=True)
    C.free()
    if B is not A:
        B.free()
    A.free()
ctx.close()
