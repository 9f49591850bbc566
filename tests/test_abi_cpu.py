"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/pemspgemm.h declares, and its host-only entry points (Matrix Market I/O) work.
No compute call is made here (no GPU in this container)."""
import ctypes
import os
import re

import numpy as np
import pytest

import pem_spgemm_b200 as pem
from pem_spgemm_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pemspgemm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pem_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(pem.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in pemspgemm.h but not exported"
    assert sorted(pem.EXPORTS) == names


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pem.PemError) as e:
        pem.Context(0)
    assert e.value.code == -6


def test_mtx_roundtrip(tmp_path):
    rows, cols, I, J, V = synth.random_sparse(123, 77, 900, seed=5)
    p = str(tmp_path / "m.mtx")
    pem.mtx_write(p, rows, cols, I, J, V)
    r2, c2, I2, J2, V2, sym = pem.mtx_read(p)
    assert (r2, c2, sym) == (rows, cols, False)
    assert np.array_equal(I, I2) and np.array_equal(J, J2) and np.array_equal(V, V2)  # bit-exact values


def test_mtx_symmetric_pattern_complex(tmp_path):
    p = tmp_path / "s.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real symmetric\n% comment\n3 3 3\n1 1 2.5\n2 1 -1\n3 2 4e0\n")
    r, c, I, J, V, sym = pem.mtx_read(str(p))
    assert sym and r == 3 and I.size == 5          # diagonal once, off-diagonals mirrored
    got = {(int(a), int(b)): float(v) for a, b, v in zip(I, J, V)}
    assert got == {(0, 0): 2.5, (1, 0): -1.0, (0, 1): -1.0, (2, 1): 4.0, (1, 2): 4.0}
    q = tmp_path / "p.mtx"
    q.write_text("%%MatrixMarket matrix coordinate pattern general\n2 3 2\n1 3\n2 2\n")
    r, c, I, J, V, sym = pem.mtx_read(str(q))
    assert (r, c) == (2, 3) and list(V) == [1.0, 1.0] and list(J) == [2, 1]
    z = tmp_path / "z.mtx"
    z.write_text("%%MatrixMarket matrix coordinate complex general\n2 2 1\n2 1 3.5 -9\n")
    r, c, I, J, V, sym = pem.mtx_read(str(z))
    assert list(V) == [3.5]                         # real part, spgemm.cu:99-107


def test_mtx_errors(tmp_path):
    with pytest.raises(pem.PemError):
        pem.mtx_read(str(tmp_path / "missing.mtx"))
    b = tmp_path / "b.mtx"
    b.write_text("%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1 1.0\n")
    with pytest.raises(pem.PemError):
        pem.mtx_read(str(b))
    o = tmp_path / "o.mtx"
    o.write_text("%%MatrixMarket matrix coordinate real general\n2 2 1\n3 1 1.0\n")
    with pytest.raises(pem.PemError):
        pem.mtx_read(str(o))


def test_mtx_large_parallel_parse(tmp_path):
    rows, cols, I, J, V = synth.random_sparse(50_000, 50_000, 300_000, seed=9)
    p = str(tmp_path / "big.mtx")
    pem.mtx_write(p, rows, cols, I, J, V)
    assert os.path.getsize(p) > (1 << 20)           # takes the multi-threaded path
    r2, c2, I2, J2, V2, _ = pem.mtx_read(p)
    assert np.array_equal(I, I2) and np.array_equal(J, J2) and np.array_equal(V, V2)
