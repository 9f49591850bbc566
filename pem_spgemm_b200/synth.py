"""Synthetic input matrices for the five BASELINE.json configs (SuiteSparse is not
available offline).  Every generator returns ``(rows, cols, I, J, V)`` with int32
0-based coordinates, float64 values, no duplicate (i, j) and a deterministic
(seeded) content, so the oracle and the CUDA engine see identical inputs.

The reference takes its inputs as Matrix-Market triplets in file order
(/root/reference/spgemm.cu:43-110); entries are therefore NOT required to be
sorted and the generators do not promise any order beyond determinism.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "laplacian2d", "webbase_like", "lp_like", "cage_like", "rmat", "random_sparse",
    "config", "CONFIG_NAMES",
]


def _dedup(rows: int, cols: int, I: np.ndarray, J: np.ndarray):
    """Drop duplicate coordinates (reference quirk 3: duplicates are undefined there)."""
    key = I.astype(np.int64) * np.int64(cols) + J.astype(np.int64)
    key = np.unique(key)
    return (key // cols).astype(np.int32), (key % cols).astype(np.int32)


def _hash_values(I: np.ndarray, J: np.ndarray) -> np.ndarray:
    """Deterministic value in (0, 1] from a 64-bit mix of (i, j)."""
    x = (I.astype(np.uint64) << np.uint64(32)) | J.astype(np.uint64)
    x ^= x >> np.uint64(33)
    x *= np.uint64(0xFF51AFD7ED558CCD)
    x ^= x >> np.uint64(33)
    x *= np.uint64(0xC4CEB9FE1A85EC53)
    x ^= x >> np.uint64(33)
    return ((x >> np.uint64(11)).astype(np.float64) + 1.0) / float(1 << 53)


def laplacian2d(g: int = 256):
    """Config 1: 2-D 5-point Laplacian on a g x g grid, natural ordering, 4 / -1."""
    n = g * g
    idx = np.arange(n, dtype=np.int64)
    x, y = idx % g, idx // g
    Is, Js, Vs = [idx], [idx], [np.full(n, 4.0)]
    for m, off in ((x > 0, -1), (x < g - 1, 1), (y > 0, -g), (y < g - 1, g)):
        Is.append(idx[m]); Js.append(idx[m] + off); Vs.append(np.full(int(m.sum()), -1.0))
    I = np.concatenate(Is).astype(np.int32)
    J = np.concatenate(Js).astype(np.int32)
    V = np.concatenate(Vs)
    return n, n, I, J, V


def webbase_like(n: int = 1_000_005, seed: int = 2, target_nnz: int = 3_650_000,
                 max_deg: int = 4700, local_frac: float = 0.7, hub_exp: float = 2.2):
    """Config 2: webbase-1M-shaped power-law matrix.

    Row degrees follow a Zipf(2.2) law capped at ``max_deg`` with the heavy rows
    placed at low indices; 70 % of a row's entries are local (col = row +- Geom(0.05)),
    30 % are global and skewed towards hub columns (col = floor(n * u^2.2)).  The
    defaults give nnz = 3,091,482, flop = 92,217,721, nnz(C) = 67,736,942 (real webbase-1M: 3.1M / 6.95e7 /
    5.11e7), i.e. a COO dump of C well above 1.5 GiB as the reference's README notes.
    """
    rng = np.random.default_rng(seed)
    deg = np.minimum(rng.zipf(2.2, size=n), max_deg).astype(np.int64)
    # heavy rows correlated with low indices: sort the top 1 % to the front
    heavy = np.argsort(-deg, kind="stable")[: n // 100]
    perm = np.arange(n)
    rest = np.setdiff1d(perm, heavy, assume_unique=True)
    order = np.concatenate([heavy, rest])
    deg = deg[order]
    scale = target_nnz / float(deg.sum())
    if scale > 1.0:
        deg = np.minimum(np.floor(deg * scale + rng.random(n)).astype(np.int64), max_deg)
    I = np.repeat(np.arange(n, dtype=np.int64), deg)
    m = I.size
    local = rng.random(m) < local_frac
    step = rng.geometric(0.05, size=m) * rng.choice(np.array([-1, 1]), size=m)
    Jl = np.clip(I + step, 0, n - 1)
    Jg = np.minimum((n * rng.random(m) ** hub_exp).astype(np.int64), n - 1)
    J = np.where(local, Jl, Jg)
    I32, J32 = _dedup(n, n, I, J)
    V = np.random.default_rng(seed + 1000).uniform(-1.0, 1.0, size=I32.size)
    return n, n, I32, J32, V


def lp_like(m: int = 100_000, n: int = 1_000_000, per_row: int = 100, seed: int = 3):
    """Config 3: rectangular LP-shaped matrix for C = A * A^T (the ``[1]`` flag).

    90 % of a row's columns fall within +-5000 of ``10 * row`` (a staircase), 10 %
    are uniform over all columns.
    """
    rng = np.random.default_rng(seed)
    I = np.repeat(np.arange(m, dtype=np.int64), per_row)
    t = I.size
    centre = (I * (n // m)).astype(np.int64)
    near = np.clip(centre + rng.integers(-5000, 5001, size=t), 0, n - 1)
    far = rng.integers(0, n, size=t)
    J = np.where(rng.random(t) < 0.9, near, far)
    I32, J32 = _dedup(m, n, I, J)
    V = np.random.default_rng(seed + 1000).uniform(-1.0, 1.0, size=I32.size)
    return m, n, I32, J32, V


def cage_like(nx: int = 172, ny: int = 173, nz: int = 173):
    """Config 4: cage15-shaped banded matrix: 3-D 19-point stencil (offsets with
    |dx|+|dy|+|dz| <= 2, each in {-1,0,1}) on an nx*ny*nz grid, x fastest; values are
    a deterministic hash of (i, j) in (0, 1]."""
    n = nx * ny * nz
    idx = np.arange(n, dtype=np.int64)
    x = idx % nx
    y = (idx // nx) % ny
    z = idx // (nx * ny)
    Is, Js = [], []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if abs(dx) + abs(dy) + abs(dz) > 2:
                    continue
                ok = ((x + dx >= 0) & (x + dx < nx) & (y + dy >= 0) & (y + dy < ny)
                      & (z + dz >= 0) & (z + dz < nz))
                src = idx[ok]
                Is.append(src.astype(np.int32))
                Js.append((src + dx + dy * nx + dz * nx * ny).astype(np.int32))
    I = np.concatenate(Is)
    J = np.concatenate(Js)
    return n, n, I, J, _hash_values(I, J)


def rmat(scale: int = 22, edge_factor: int = 16, a: float = 0.35, b: float = 0.25,
         c: float = 0.25, seed: int = 22):
    """Config 5: R-MAT graph, 2^scale vertices, edge_factor * 2^scale generated edges,
    de-duplicated.  Default (a,b,c,d) = (0.35,0.25,0.25,0.15)."""
    rng = np.random.default_rng(seed)
    n = 1 << scale
    m = edge_factor * n
    I = np.zeros(m, dtype=np.int64)
    J = np.zeros(m, dtype=np.int64)
    ab, abc = a + b, a + b + c
    for _ in range(scale):
        u = rng.random(m)
        I = (I << 1) | (u >= ab)
        J = (J << 1) | (((u >= a) & (u < ab)) | (u >= abc))
    I32, J32 = _dedup(n, n, I, J)
    V = np.random.default_rng(seed + 1000).uniform(-1.0, 1.0, size=I32.size)
    return n, n, I32, J32, V


def random_sparse(rows: int, cols: int, nnz: int, seed: int = 0, shuffle: bool = True,
                  integer_values: bool = False):
    """Small uniform random matrix for unit tests (ragged dims, empty rows allowed)."""
    rng = np.random.default_rng(seed)
    if rows == 0 or cols == 0 or nnz == 0:
        z = np.zeros(0, dtype=np.int32)
        return rows, cols, z, z.copy(), np.zeros(0)
    I = rng.integers(0, rows, size=nnz)
    J = rng.integers(0, cols, size=nnz)
    I32, J32 = _dedup(rows, cols, I, J)
    if integer_values:
        V = rng.integers(-4, 5, size=I32.size).astype(np.float64)
    else:
        V = rng.uniform(-1.0, 1.0, size=I32.size)
    if shuffle:
        p = rng.permutation(I32.size)
        I32, J32, V = I32[p], J32[p], V[p]
    return rows, cols, I32, J32, V


CONFIG_NAMES = {
    1: "laplace2d_256",
    2: "webbase1m_like",
    3: "lp_like_aat",
    4: "cage15_like",
    5: "rmat22_like",
}


def config(k: int, small: bool = False):
    """Return ``(name, transpose_b, (rows, cols, I, J, V))`` for BASELINE.json config k.

    ``small=True`` returns a structurally similar but much smaller instance that the
    host oracle finishes in seconds (used by the parity tests).
    """
    if k == 1:
        return CONFIG_NAMES[1], False, laplacian2d(64 if small else 256)
    if k == 2:
        if small:
            return CONFIG_NAMES[2], False, webbase_like(n=20_003, target_nnz=60_000, max_deg=300)
        return CONFIG_NAMES[2], False, webbase_like()
    if k == 3:
        if small:
            return CONFIG_NAMES[3], True, lp_like(m=2_000, n=20_000, per_row=30)
        return CONFIG_NAMES[3], True, lp_like()
    if k == 4:
        if small:
            return CONFIG_NAMES[4], False, cage_like(28, 29, 29)
        return CONFIG_NAMES[4], False, cage_like()
    if k == 5:
        if small:
            return CONFIG_NAMES[5], False, rmat(scale=12, edge_factor=8)
        return CONFIG_NAMES[5], False, rmat()
    raise ValueError(f"unknown config {k}")
