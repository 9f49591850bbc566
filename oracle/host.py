"""TEST INFRASTRUCTURE — ctypes front end of ``oracle/gustavson.cpp`` (OpenMP Gustavson).

See ``oracle/__init__.py`` for who may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc + OpenMP)."""
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "gustavson.cpp"))):
        subprocess.check_call(["make", "-s", "-C", _HERE, "_build/liboracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_max_threads.restype = C.c_int
        L.oracle_set_threads.restype = None
        L.oracle_set_threads.argtypes = [C.c_int]
        L.oracle_coo_to_csr.restype = C.c_int
        L.oracle_coo_to_csr.argtypes = [C.c_int, C.c_int, C.c_int64, _i32p, _i32p, _f64p, _i64p, _i32p, _f64p]
        L.oracle_csr_transpose.restype = None
        L.oracle_csr_transpose.argtypes = [C.c_int, C.c_int, _i64p, _i32p, _f64p, _i64p, _i32p, _f64p]
        L.oracle_flop.restype = C.c_uint64
        L.oracle_flop.argtypes = [C.c_int, _i64p, _i32p, _i64p]
        L.oracle_spgemm_symbolic.restype = C.c_int64
        L.oracle_spgemm_symbolic.argtypes = [C.c_int, C.c_int, _i64p, _i32p, _i64p, _i32p, _i64p]
        L.oracle_spgemm_numeric.restype = None
        L.oracle_spgemm_numeric.argtypes = [C.c_int, C.c_int, _i64p, _i32p, _f64p, _i64p, _i32p, _f64p,
                                            _i64p, _i32p, _f64p]
        _f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
        L.oracle_spgemm_numeric_f32.restype = None
        L.oracle_spgemm_numeric_f32.argtypes = [C.c_int, C.c_int, _i64p, _i32p, _f32p, _i64p, _i32p, _f32p,
                                                _i64p, _i32p, _f32p]
        _lib = L
    return _lib


def set_threads(n: int) -> int:
    """Use n OpenMP threads from now on (torchrun exports OMP_NUM_THREADS=1); returns the count in effect."""
    lib().oracle_set_threads(int(n))
    return int(lib().oracle_max_threads())


@dataclass
class CSR:
    rows: int
    cols: int
    ptr: np.ndarray  # int64[rows+1]
    idx: np.ndarray  # int32[nnz], ascending inside a row
    val: np.ndarray  # float64[nnz]

    @property
    def nnz(self) -> int:
        return int(self.ptr[-1])

    def to_coo(self):
        """(rows, cols, vals) sorted by (row, col) — the order of the reference's dump
        (/root/reference/spgemm.cu:1515-1518)."""
        r = np.repeat(np.arange(self.rows, dtype=np.int32), np.diff(self.ptr))
        return r, self.idx, self.val


def coo_to_csr(rows, cols, I, J, V) -> CSR:
    I = np.ascontiguousarray(I, np.int32); J = np.ascontiguousarray(J, np.int32)
    V = np.ascontiguousarray(V, np.float64)
    n = I.size
    ptr = np.zeros(rows + 1, np.int64); idx = np.empty(n, np.int32); val = np.empty(n, np.float64)
    rc = lib().oracle_coo_to_csr(rows, cols, n, I, J, V, ptr, idx, val)
    if rc == -1:
        raise ValueError("coordinate out of range")
    if rc == -2:
        raise ValueError("duplicate coordinate")
    return CSR(rows, cols, ptr, idx, val)


def transpose(A: CSR) -> CSR:
    ptr = np.zeros(A.cols + 1, np.int64); idx = np.empty(A.nnz, np.int32); val = np.empty(A.nnz, np.float64)
    lib().oracle_csr_transpose(A.rows, A.cols, A.ptr, A.idx, A.val, ptr, idx, val)
    return CSR(A.cols, A.rows, ptr, idx, val)


def row_slice(A: CSR, r0: int, r1: int) -> CSR:
    """Rows [r0, r1) of A as a CSR of its own (views of idx/val): the oracle side of a tile-row panel."""
    r0 = max(0, min(r0, A.rows)); r1 = max(r0, min(r1, A.rows))
    b, e = int(A.ptr[r0]), int(A.ptr[r1])
    return CSR(r1 - r0, A.cols, np.ascontiguousarray(A.ptr[r0:r1 + 1] - b), A.idx[b:e], A.val[b:e])


def flop(A: CSR, B: CSR) -> int:
    return int(lib().oracle_flop(A.rows, A.ptr, A.idx, B.ptr))


def spgemm(A: CSR, B: CSR, timing: dict | None = None) -> CSR:
    """C = A*B, structural zeros kept, fma accumulation in ascending k."""
    assert A.cols == B.rows
    t0 = time.perf_counter()
    cp = np.zeros(A.rows + 1, np.int64)
    nnz = lib().oracle_spgemm_symbolic(A.rows, B.cols, A.ptr, A.idx, B.ptr, B.idx, cp)
    cj = np.empty(nnz, np.int32); cx = np.empty(nnz, np.float64)
    lib().oracle_spgemm_numeric(A.rows, B.cols, A.ptr, A.idx, A.val, B.ptr, B.idx, B.val, cp, cj, cx)
    if timing is not None:
        timing["seconds"] = time.perf_counter() - t0
        timing["threads"] = int(lib().oracle_max_threads())
    return CSR(A.rows, B.cols, cp, cj, cx)


def spgemm_f32(A: CSR, B: CSR) -> CSR:
    """The fp32 instantiation: values rounded to float32 first, single-precision fma in ascending k."""
    assert A.cols == B.rows
    cp = np.zeros(A.rows + 1, np.int64)
    nnz = lib().oracle_spgemm_symbolic(A.rows, B.cols, A.ptr, A.idx, B.ptr, B.idx, cp)
    cj = np.empty(nnz, np.int32); cx = np.empty(nnz, np.float32)
    lib().oracle_spgemm_numeric_f32(A.rows, B.cols, A.ptr, A.idx, np.ascontiguousarray(A.val, np.float32),
                                    B.ptr, B.idx, np.ascontiguousarray(B.val, np.float32), cp, cj, cx)
    return CSR(A.rows, B.cols, cp, cj, cx)


def spgemm_from_coo(rows, cols, I, J, V, transpose_b: bool, timing: dict | None = None):
    """The CLI's two products (/root/reference/spgemm.cu:782-792): A*A, or A*A^T when
    ``transpose_b``.  Returns (A, B, C) as CSR."""
    A = coo_to_csr(rows, cols, I, J, V)
    B = transpose(A) if transpose_b else A
    return A, B, spgemm(A, B, timing)
