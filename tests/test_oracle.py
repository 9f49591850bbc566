"""CPU tests: the host oracle against its pins (config-1 known answers from SURVEY.md
section 8c, scipy.sparse as an independent implementation) and the numpy restatement of
the tiled format against the same numbers."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import host, tiles
from pem_spgemm_b200 import synth


def _scipy(rows, cols, I, J, V):
    return sp.csr_matrix((V, (I, J)), shape=(rows, cols))


def test_config1_known_answers():
    rows, cols, I, J, V = synth.laplacian2d(256)
    assert rows == 65_536 and I.size == 326_656
    A, B, C = host.spgemm_from_coo(rows, cols, I, J, V, transpose_b=False)
    assert host.flop(A, B) == 1_629_192
    assert C.nnz == 846_852
    assert C.val.sum() == 1032.0
    assert np.abs(C.val).sum() == 4_178_952.0
    r, c, v = C.to_coo()
    assert (r[0], c[0], v[0]) == (0, 0, 18.0)
    interior = 128 * 256 + 128
    row = slice(C.ptr[interior], C.ptr[interior + 1])
    assert C.val[row][C.idx[row] == interior][0] == 20.0


def test_config1_tile_known_answers():
    rows, cols, I, J, V = synth.laplacian2d(256)
    T = tiles.tile_format(rows, cols, I, J, V)
    assert T.cnt == 19_936 and T.tile_rows == 4_096
    P = tiles.tiled_product(T, T, keep_empty=True)
    assert P.c_tile_col.size == 50_532
    assert P.pairs_a.size == 97_512
    nnz_t = np.diff(P.c_tile_nnz_ptr)
    assert int((nnz_t > 0).sum()) == 43_364
    assert int(P.c_tile_nnz_ptr[-1]) == 846_852
    Q = tiles.tiled_product(T, T, keep_empty=False)
    assert Q.c_tile_col.size == 43_364
    assert int(Q.c_tile_nnz_ptr[-1]) == 846_852


@pytest.mark.parametrize("shape,nnz,seed", [((37, 37), 200, 1), ((300, 300), 4000, 2),
                                            ((1000, 1000), 3000, 3), ((16, 16), 256, 4)])
def test_oracle_vs_scipy_square(shape, nnz, seed):
    rows, cols, I, J, V = synth.random_sparse(*shape, nnz, seed=seed, integer_values=True)
    A, B, C = host.spgemm_from_coo(rows, cols, I, J, V, transpose_b=False)
    S = _scipy(rows, cols, I, J, V)
    # structural product: scipy on the pattern (values 1) never cancels
    pat = sp.csr_matrix((np.ones_like(V), (I, J)), shape=(rows, cols))
    R = (pat @ pat).tocsr(); R.sort_indices()
    assert np.array_equal(C.ptr, R.indptr.astype(np.int64))
    assert np.array_equal(C.idx, R.indices.astype(np.int32))
    D = (S @ S).toarray()
    r, c, v = C.to_coo()
    assert np.array_equal(v, D[r, c])        # small integers: exact in any order
    assert host.flop(A, B) == int(R.nnz and (pat @ pat).sum())


def test_oracle_vs_scipy_aat_float():
    rows, cols, I, J, V = synth.random_sparse(120, 700, 3000, seed=7)
    A, B, C = host.spgemm_from_coo(rows, cols, I, J, V, transpose_b=True)
    assert (B.rows, B.cols) == (700, 120)
    S = _scipy(rows, cols, I, J, V)
    D = (S @ S.T).toarray()
    r, c, v = C.to_coo()
    np.testing.assert_allclose(v, D[r, c], rtol=1e-12, atol=1e-300)
    pat = sp.csr_matrix((np.ones_like(V), (I, J)), shape=(rows, cols))
    assert C.nnz == (pat @ pat.T).nnz


def test_structural_zero_is_kept():
    # [1 1; 1 -1] squared has cancelled off-diagonal entries that must stay in C
    I = np.array([0, 0, 1, 1], np.int32); J = np.array([0, 1, 0, 1], np.int32)
    V = np.array([1.0, 1.0, 1.0, -1.0])
    _, _, C = host.spgemm_from_coo(2, 2, I, J, V, transpose_b=False)
    assert C.nnz == 4 and np.array_equal(C.val, [2.0, 0.0, 0.0, 2.0])


def test_duplicate_and_range_errors():
    I = np.array([0, 0], np.int32); J = np.array([1, 1], np.int32); V = np.ones(2)
    with pytest.raises(ValueError):
        host.coo_to_csr(2, 2, I, J, V)
    with pytest.raises(ValueError):
        host.coo_to_csr(2, 2, np.array([2], np.int32), np.array([0], np.int32), np.ones(1))


@pytest.mark.parametrize("k", [2, 3, 4, 5])
def test_tiles_agree_with_oracle_small_configs(k):
    """The tile-level restatement and the Gustavson oracle describe the same C."""
    name, tb, (rows, cols, I, J, V) = synth.config(k, small=True)
    A, B, C = host.spgemm_from_coo(rows, cols, I, J, V, transpose_b=tb)
    TA = tiles.tile_format(rows, cols, I, J, V)
    TB = tiles.tile_format(rows, cols, I, J, V, transpose=tb)
    for keep in (True, False):
        P = tiles.tiled_product(TA, TB, keep_empty=keep)
        r, c = tiles.product_to_coo(P)
        ro, co, _ = C.to_coo()
        assert np.array_equal(r, ro) and np.array_equal(c, co)


def test_tile_format_layout_small():
    rows, cols, I, J, V = synth.random_sparse(50, 70, 400, seed=11)
    T = tiles.tile_format(rows, cols, I, J, V)
    assert T.tile_rows == 4 and T.tile_cols == 5
    # rebuild COO from the tiled arrays
    t = np.repeat(np.arange(T.cnt), np.diff(T.tile_nnz_ptr))
    r = (T.tile_row[t].astype(np.int64) << 4) + (T.row_col_idx >> 4)
    c = (T.tile_col[t].astype(np.int64) << 4) + (T.row_col_idx & 15)
    want = {(int(i), int(j)): float(v) for i, j, v in zip(I, J, V)}
    got = {(int(i), int(j)): float(v) for i, j, v in zip(r, c, T.vals)}
    assert want == got
    # masks / row_ptr / transposed masks are consistent
    for tt in range(T.cnt):
        sl = slice(T.tile_nnz_ptr[tt], T.tile_nnz_ptr[tt + 1])
        rr = T.row_col_idx[sl] >> 4; cc = T.row_col_idx[sl] & 15
        m = np.zeros(16, np.uint16); mt = np.zeros(16, np.uint16)
        for a, b in zip(rr, cc):
            m[int(a)] |= 1 << int(b); mt[int(b)] |= 1 << int(a)
        assert np.array_equal(m, T.masks[tt]) and np.array_equal(mt, T.masks_t[tt])
        cnt = np.array([bin(int(x)).count("1") for x in m])
        assert np.array_equal(T.row_ptr[tt], (np.cumsum(cnt) - cnt).astype(np.uint8))


def test_fp32_oracle_agrees_with_fp64_on_exactly_representable_products():
    """Integer-valued inputs: every partial sum is exact in both precisions, so the fp32 instantiation of the
    numeric pass must reproduce the fp64 one after rounding; structure is shared."""
    rows, cols, I, J, V = synth.random_sparse(300, 300, 4000, seed=12, integer_values=True)
    A = host.coo_to_csr(rows, cols, I, J, V)
    C64, C32 = host.spgemm(A, A), host.spgemm_f32(A, A)
    assert C32.val.dtype == np.float32
    assert np.array_equal(C64.ptr, C32.ptr) and np.array_equal(C64.idx, C32.idx)
    assert np.array_equal(C64.val.astype(np.float32), C32.val)
