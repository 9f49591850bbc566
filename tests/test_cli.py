"""The drop-in process surface: `pemspgemm <mtx> [0/1] [1]` (reference: spgemm.cu:720-1568,
README.md:39-53).  Argument handling is checked on CPU (it happens before any CUDA call); the
report, CSV row and COO dump are checked on the GPU against the host oracle."""
import os
import subprocess

import numpy as np
import pytest

import pem_spgemm_b200 as pem
from pem_spgemm_b200 import synth


def _run(args, cwd, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([pem.CLI_PATH] + args, cwd=cwd, env=e, capture_output=True, text=True, timeout=300)


def test_cli_usage_and_rectangular_messages(tmp_path):
    assert os.path.exists(pem.CLI_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    out = _run([], tmp_path)                                    # spgemm.cu:722-725
    assert out.returncode == 1 and "Provide a matrix market file path. Exiting." in out.stdout
    out = _run(["a", "0", "1", "extra"], tmp_path)
    assert out.returncode == 1 and "Provide a matrix market file path. Exiting." in out.stdout
    rows, cols, I, J, V = synth.random_sparse(5, 9, 12, seed=1)
    mtx = str(tmp_path / "rect.mtx")
    pem.mtx_write(mtx, rows, cols, I, J, V)
    out = _run([mtx, "0"], tmp_path)                            # spgemm.cu:782-786
    assert out.returncode == 1 and "input is rectangular. Only AAt is possible. Exiting." in out.stdout
    out = _run([str(tmp_path / "missing.mtx"), "0"], tmp_path)  # the reference would crash; we report
    assert out.returncode == 2 and "cannot read" in out.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("aat", [False, True])
def test_cli_report_csv_and_dump(tmp_path, aat):
    from oracle import host
    if aat:
        rows, cols, I, J, V = synth.random_sparse(70, 300, 1500, seed=5)
    else:
        rows, cols, I, J, V = synth.laplacian2d(24)
    os.makedirs(tmp_path / "in")
    mtx = str(tmp_path / "in" / "case_a.mtx")
    pem.mtx_write(mtx, rows, cols, I, J, V)
    env = {"PEM_DUMP_DIR": str(tmp_path), "PEM_REPEAT": "2", "PEM_WARMUP": "1"}
    out = _run([mtx, "1"] + (["1"] if aat else []), tmp_path, env)
    assert out.returncode == 0, out.stderr
    oA, oB, oC = host.spgemm_from_coo(rows, cols, I, J, V, aat)
    ro, co, vo = oC.to_coo()
    flop = host.flop(oA, oB)
    # report lines (spgemm.cu:1406-1422)
    for needle in ("<---Program done--->", f"Flop count: {flop}", f"C nnz: {ro.size}", "pemSpGEMM took",
                   "Saving results to", "CLEANING UP RESOURCES"):
        assert needle in out.stdout, needle
    # COO dump (spgemm.cu:1527-1560): NNZ without newline, 0-based rows/cols, values with 17 digits
    assert open(tmp_path / "SPGEMM_RESULT_NNZ.txt").read() == str(ro.size)
    r = np.loadtxt(tmp_path / "SPGEMM_RESULT_ROWS.txt", dtype=np.int64, ndmin=1)
    c = np.loadtxt(tmp_path / "SPGEMM_RESULT_COLS.txt", dtype=np.int64, ndmin=1)
    v = np.loadtxt(tmp_path / "SPGEMM_RESULT_VALS.txt", dtype=np.float64, ndmin=1)
    assert np.array_equal(r, ro) and np.array_equal(c, co)
    np.testing.assert_allclose(v, vo, rtol=1e-12, atol=1e-16)   # std::fixed, 17 decimals
    # the text itself is what `std::fixed << setprecision(max_digits10)` prints (spgemm.cu:1555-1558)
    lines = open(tmp_path / "SPGEMM_RESULT_VALS.txt").read().split("\n")
    assert lines[-1] == "" and lines[:-1] == ["%.17f" % x for x in vo]
    # "C tiles" is the reference's count: every structurally reachable tile, empty ones included (spgemm.cu:1420)
    from oracle import tiles as otiles
    P = otiles.tiled_product(otiles.tile_format(rows, cols, I, J, V), otiles.tile_format(rows, cols, I, J, V, transpose=aat),
                             keep_empty=True)
    assert f"C tiles: {P.c_tile_col.size} " in out.stdout
    # CSV (spgemm.cu:1424-1450, README.md:51-53): one row, preceded by a newline, 14 fields, no header
    raw = open(tmp_path / "pemspgemm_benchmark_result.csv").read()
    assert raw.startswith("\n") and raw.count("\n") == 1
    f = raw.strip().split(",")
    assert len(f) == 14 and f[0] == "case_a" and int(f[1]) == flop and int(f[2]) == ro.size
    assert f[3] == f"{flop / ro.size:.2f}"
    assert all(len(x.split(".")[1]) == 2 for x in f[3:])
    # the same product in 3 sequential tile-row panels: identical dump
    env3 = dict(env, PEM_PANELS="3", PEM_DUMP_DIR=str(tmp_path / "p3"), PEM_CSV=str(tmp_path / "p3.csv"))
    os.makedirs(tmp_path / "p3")
    out = _run([mtx, "1"] + (["1"] if aat else []), tmp_path, env3)
    assert out.returncode == 0, out.stderr
    assert f"C nnz: {ro.size}" in out.stdout
    for name in ("NNZ", "ROWS", "COLS", "VALS"):
        assert open(tmp_path / "p3" / f"SPGEMM_RESULT_{name}.txt").read() == open(tmp_path / f"SPGEMM_RESULT_{name}.txt").read()
    out = _run([mtx, "0"] + (["1"] if aat else []), tmp_path, env)      # second run appends
    assert out.returncode == 0 and "Not saving results. Exiting." in out.stdout
    assert open(tmp_path / "pemspgemm_benchmark_result.csv").read().count("\n") == 2


@pytest.mark.gpu
def test_cli_symmetric_file_is_expanded_with_the_diagonal_once(tmp_path):
    """A `symmetric` Matrix Market file holds the lower triangle; the reader expands it to general form and
    emits each diagonal entry ONCE.  (fast_matrix_market v1.7.6, which the reference links, would by default add
    an extra zero-valued (i,i) duplicate per diagonal entry - duplicates are undefined behaviour in the
    reference, spgemm.cu:186-191 - so this engine's printed Nnz differs from the reference's for such files;
    DESIGN.md section 8 and INTEGRATION.md list the deviation.)"""
    from oracle import host
    n = 40
    rng = np.random.default_rng(4)
    I = rng.integers(0, n, 300); J = rng.integers(0, n, 300)
    lo = np.unique(np.stack([np.maximum(I, J), np.minimum(I, J)], 1), axis=0)        # lower triangle incl. diagonal
    V = rng.uniform(-1, 1, len(lo))
    mtx = str(tmp_path / "sym.mtx")
    with open(mtx, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real symmetric\n")
        f.write(f"{n} {n} {len(lo)}\n")
        for (i, j), v in zip(lo, V):
            f.write(f"{i + 1} {j + 1} {float(v)!r}\n")
    off = lo[:, 0] != lo[:, 1]
    gI = np.concatenate([lo[:, 0], lo[off, 1]]).astype(np.int32); gJ = np.concatenate([lo[:, 1], lo[off, 0]]).astype(np.int32)
    gV = np.concatenate([V, V[off]])
    rows, cols, rI, rJ, rV, sym = pem.mtx_read(mtx)
    assert sym and rI.size == gI.size
    out = _run([mtx, "1"], tmp_path, {"PEM_DUMP_DIR": str(tmp_path), "PEM_REPEAT": "1", "PEM_WARMUP": "0"})
    assert out.returncode == 0, out.stderr
    assert f"Nnz: {gI.size}" in out.stdout
    _, _, oC = host.spgemm_from_coo(n, n, gI, gJ, gV, False)
    ro, co, vo = oC.to_coo()
    r = np.loadtxt(tmp_path / "SPGEMM_RESULT_ROWS.txt", dtype=np.int64, ndmin=1)
    c = np.loadtxt(tmp_path / "SPGEMM_RESULT_COLS.txt", dtype=np.int64, ndmin=1)
    v = np.loadtxt(tmp_path / "SPGEMM_RESULT_VALS.txt", dtype=np.float64, ndmin=1)
    assert np.array_equal(r, ro) and np.array_equal(c, co)
    np.testing.assert_allclose(v, vo, rtol=1e-12, atol=1e-16)


@pytest.mark.gpu
def test_cpp_abi_client(tmp_path):
    """examples/abi_demo.cpp: a C++ program that uses only include/pemspgemm.h (the Laplacian is
    symmetric, so A*A^T through pem_tiled_transpose must equal A^2; figures for the 256 grid are the
    pinned config-1 answers of SURVEY.md section 8c)."""
    demo = os.path.join(os.path.dirname(pem.CLI_PATH), "abi_demo")
    out = subprocess.run([demo, "256"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "nnzA=326656 flop=1629192 C_tiles=43364 C_nnz=846852 sum=1032.0 abs_sum=4178952.0 C[0,0]=18.0" in out.stdout
    assert "A*A^T: C_nnz=846852 sum=1032.0 abs_sum=4178952.0" in out.stdout
