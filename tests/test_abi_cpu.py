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


def test_mtx_reader_edge_files(tmp_path):
    """The reader maps the file: no terminator behind the last byte, CRLF line ends, blank lines, an empty file,
    a header-only file."""
    e = tmp_path / "e.mtx"
    e.write_bytes(b"")
    with pytest.raises(pem.PemError):
        pem.mtx_read(str(e))
    n = tmp_path / "n.mtx"
    n.write_bytes(b"%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1 1.5\n2 2 -2.25")   # no final newline
    r, c, I, J, V, _ = pem.mtx_read(str(n))
    assert (r, c, list(I), list(J), list(V)) == (2, 2, [0, 1], [0, 1], [1.5, -2.25])
    w = tmp_path / "w.mtx"
    w.write_bytes(b"%%MatrixMarket matrix coordinate integer general\r\n% c\r\n\r\n3 2 2\r\n\r\n3 1 7\r\n1 2 +4\r\n")
    r, c, I, J, V, _ = pem.mtx_read(str(w))
    assert (r, c, list(I), list(J), list(V)) == (3, 2, [2, 0], [0, 1], [7.0, 4.0])
    z = tmp_path / "z.mtx"
    z.write_bytes(b"%%MatrixMarket matrix coordinate real general\n4 5 0\n")
    r, c, I, J, V, _ = pem.mtx_read(str(z))
    assert (r, c, I.size) == (4, 5, 0)
    h = tmp_path / "h.mtx"
    h.write_bytes(b"%%MatrixMarket matrix coordinate real general")
    with pytest.raises(pem.PemError):
        pem.mtx_read(str(h))


def _fixed17(x):
    return "%.17f" % x          # the digits `std::fixed << std::setprecision(17)` prints (spgemm.cu:1529)


def test_dump_lines_match_the_reference_format(tmp_path):
    """pem_write_lines_*: one number per line, doubles as the reference's dump prints them."""
    vals = np.array([0.0, -0.0, 1.0, -1.5, 0.1, 1.0 / 3.0, 2.5e-17, 4.9e-18, 5e-324, 1e-20, 123456789.123456789, 1e22,
                     -7.000000000000001, 0.5, 0.49999999999999994, 1e15 + 0.3, 2.0 ** 53, 2.0 ** 53 + 2, 1.7976931348623157e308,
                     65536.000000000001, 1e-5, 3.0000000000000004e-5, 999999.99999999988], np.float64)
    p = str(tmp_path / "v.txt")
    pem.write_lines(p, vals)
    assert open(p).read() == "".join(_fixed17(x) + "\n" for x in vals)
    ints = np.array([0, 1, -1, 2147483647, -2147483648, 10, 99, 100, 123456789], np.int32)
    q = str(tmp_path / "i.txt")
    pem.write_lines(q, ints)
    assert open(q).read() == "".join(f"{int(x)}\n" for x in ints)
    pem.write_lines(q, ints[:3], append=True)                      # a second panel continues the file
    assert open(q).read() == "".join(f"{int(x)}\n" for x in list(ints) + list(ints[:3]))
    pem.write_lines(q, ints[:2])                                   # and a fresh dump truncates it
    assert open(q).read() == "0\n1\n"
    pem.write_lines(q, ints[:0])
    assert open(q).read() == ""


def test_dump_lines_parallel_slices_keep_order(tmp_path):
    """More lines than one slice: every host thread formats a slice and writes it at its own offset."""
    rng = np.random.default_rng(11)
    n = 700_000
    vals = np.concatenate([rng.uniform(-1, 1, n // 2), rng.standard_normal(n // 4) * 1e6,
                           rng.uniform(0, 1, n // 4) * 10.0 ** rng.integers(-25, 12, n // 4)])
    p = str(tmp_path / "v.txt")
    pem.write_lines(p, vals)
    assert open(p).read() == "".join(_fixed17(x) + "\n" for x in vals)
    pem.write_lines(p, vals[:100_000], append=True)
    got = open(p).read().split("\n")
    assert len(got) == n + 100_000 + 1 and got[n] == _fixed17(vals[0]) and got[-2] == _fixed17(vals[99_999])
    ints = rng.integers(-2 ** 31, 2 ** 31, n, dtype=np.int64).astype(np.int32)
    q = str(tmp_path / "i.txt")
    pem.write_lines(q, ints)
    assert np.array_equal(np.loadtxt(q, dtype=np.int64), ints.astype(np.int64))


def test_mtx_reader_fuzz_against_a_python_parser(tmp_path):
    """Seeded fuzz of the reader: every field / symmetry combination, random spacing (blanks, tabs, CRLF), comment and
    blank lines between entries, signs and exponents in every spelling from_chars / strtod accept; compared with a
    line-by-line Python parse of the same text."""
    rng = np.random.default_rng(2024)

    def num(x):
        style = rng.integers(0, 5)
        if style == 0:
            return repr(float(x))
        if style == 1:
            return "%.17e" % x
        if style == 2:
            return ("+" if x >= 0 else "") + "%.6f" % x
        if style == 3:
            return "%dE0" % int(x) if float(x).is_integer() else "%.10g" % x
        return "%.3E" % x

    for case in range(40):
        field = ["real", "integer", "pattern", "complex"][case % 4]
        symm = ["general", "symmetric", "skew-symmetric"][(case // 4) % 3]
        rows = int(rng.integers(1, 40))
        cols = rows if symm != "general" else int(rng.integers(1, 40))
        cells = [(i, j) for i in range(rows) for j in range(cols)
                 if symm == "general" or j < i or (j == i and symm == "symmetric")]
        take = rng.permutation(len(cells))[: int(rng.integers(0, min(len(cells), 60) + 1))]
        sep = lambda: "".join(rng.choice([" ", "\t", "  "], size=int(rng.integers(1, 3))))
        eol = "\r\n" if case % 5 == 0 else "\n"
        lines, want = [], {}
        for t in take:
            i, j = cells[int(t)]
            v = float(rng.integers(-50, 50)) if field == "integer" else float(np.round(rng.normal() * 10.0 ** int(rng.integers(-8, 8)), 12))
            if field == "pattern":
                txt, v = "", 1.0
            elif field == "integer":
                txt = sep() + ("+%d" % v if v > 0 and rng.random() < 0.3 else "%d" % v)
            elif field == "complex":
                txt = sep() + num(v) + sep() + num(rng.normal())
                v = float(txt.split()[0])
            else:
                txt = sep() + num(v)
                v = float(txt.split()[0])
            lead = sep() if rng.random() < 0.2 else ""
            lines.append(f"{lead}{i + 1}{sep()}{j + 1}{txt}" + (sep() if rng.random() < 0.2 else ""))
            if rng.random() < 0.1:
                lines.append("% a comment between entries")
            if rng.random() < 0.1:
                lines.append("")
            want[(i, j)] = v
            if symm != "general" and i != j:
                want[(j, i)] = -v if symm == "skew-symmetric" else v
        text = f"%%MatrixMarket matrix coordinate {field} {symm}{eol}% generated{eol}{eol}{rows} {cols} {len(take)}{eol}"
        text += "".join(l + eol for l in lines)
        if case % 7 == 0 and lines:
            text = text[: -len(eol)]                      # no terminator behind the last entry
        p = tmp_path / f"f{case}.mtx"
        p.write_bytes(text.encode())
        r, c, I, J, V, sym = pem.mtx_read(str(p))
        got = {(int(a), int(b)): float(v) for a, b, v in zip(I, J, V)}
        assert (r, c) == (rows, cols) and I.size == len(want), (case, field, symm)
        assert got == want, (case, field, symm)
        assert sym == (symm == "symmetric")


def test_mtx_large_parallel_parse_with_comment_and_blank_lines(tmp_path):
    """Multi-threaded reader on a body that holds comment and blank lines: every chunk's room is the count of its
    newline bytes, so the chunks behind such lines have to move down when the gaps are closed."""
    rows, cols, I, J, V = synth.random_sparse(40_000, 30_000, 200_000, seed=12)
    p = str(tmp_path / "big.mtx")
    pem.mtx_write(p, rows, cols, I, J, V)
    text = open(p).read().split("\n")
    head, body = text[:2], text[2:]
    assert body[-1] == ""
    body = body[:-1]
    rng = np.random.default_rng(3)
    out = []
    for k, line in enumerate(body):
        out.append(line)
        if k % 997 == 0:
            out.append("% comment in the body")
        if k % 1499 == 0:
            out.append("" if rng.random() < 0.5 else "  \t ")
    q = tmp_path / "big2.mtx"
    q.write_text("\n".join(head + out))                    # and no newline behind the last entry
    assert os.path.getsize(q) > (1 << 20)
    r2, c2, I2, J2, V2, _ = pem.mtx_read(str(q))
    assert (r2, c2) == (rows, cols)
    assert np.array_equal(I, I2) and np.array_equal(J, J2) and np.array_equal(V, V2)
    # one entry too few / too many for the size line is still an error
    bad = tmp_path / "big3.mtx"
    bad.write_text("\n".join(head + out[:-1]))
    with pytest.raises(pem.PemError):
        pem.mtx_read(str(bad))
