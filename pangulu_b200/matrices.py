"""Deterministic synthetic inputs of the BASELINE.json configurations.

Every generator returns a scipy CSC matrix with sorted indices and a full diagonal.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def _stencil(shape, offsets, diag_value, off_value, dtype):
    dims = len(shape)
    n = int(np.prod(shape))
    idx = np.arange(n, dtype=np.int64).reshape(shape)
    rows = [np.arange(n, dtype=np.int64)]
    cols = [np.arange(n, dtype=np.int64)]
    vals = [np.full(n, diag_value, dtype=dtype)]
    for off in offsets:
        src = [slice(max(0, -o), shape[d] - max(0, o)) for d, o in enumerate(off)]
        dst = [slice(max(0, o), shape[d] - max(0, -o)) for d, o in enumerate(off)]
        r = idx[tuple(src)].ravel()
        c = idx[tuple(dst)].ravel()
        rows.append(r)
        cols.append(c)
        vals.append(np.full(r.size, off_value, dtype=dtype))
    A = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n), dtype=dtype)
    A.sort_indices()
    return A


def laplace2d(nx: int, ny: int | None = None, dtype=np.float64):
    """C2: 5-point Laplacian, diagonal 4, off-diagonals -1."""
    ny = nx if ny is None else ny
    offs = [(1, 0), (-1, 0), (0, 1), (0, -1)]
    return _stencil((nx, ny), offs, 4.0, -1.0, dtype)


def poisson3d(nx: int, ny: int | None = None, nz: int | None = None, dtype=np.float64):
    """C3: 7-point Poisson, diagonal 6, off-diagonals -1."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    offs = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]
    return _stencil((nx, ny, nz), offs, 6.0, -1.0, dtype)


def stencil27(nx: int, dtype=np.float64):
    """C4: 27-point stencil, diagonal 26, off-diagonals -1."""
    offs = [(a, b, c) for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1) if (a, b, c) != (0, 0, 0)]
    return _stencil((nx, nx, nx), offs, 26.0, -1.0, dtype)


def circuit_like(n: int, seed: int = 20240607, alpha: float = 2.1, dmin: int = 2, dmax: int = 1000, dtype=np.complex128):
    """C5: unsymmetric random sparse matrix with power-law row degrees, strictly diagonally dominant."""
    rng = np.random.Generator(np.random.PCG64(seed))
    u = rng.random(n)
    deg = np.minimum(dmax, np.floor(dmin * (1.0 - u) ** (-1.0 / (alpha - 1.0)))).astype(np.int64)
    deg = np.minimum(deg, n - 1)
    rows = np.repeat(np.arange(n, dtype=np.int64), deg)
    cols = rng.integers(0, n, size=rows.size, dtype=np.int64)
    keep = rows != cols
    rows, cols = rows[keep], cols[keep]
    dt = np.dtype(dtype)
    if dt.kind == "c":
        vals = (rng.uniform(-1, 1, rows.size) * (1.0 + 1j * rng.uniform(-1, 1, rows.size))).astype(dt)
    else:
        vals = rng.uniform(-1, 1, rows.size).astype(dt)
    A = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
    A.sum_duplicates()
    rowsum = np.asarray(abs(A).sum(axis=1)).ravel()
    A = (A + sp.diags((1.0 + rowsum).astype(dt))).tocsc()
    A.sort_indices()
    return A.astype(dt)


def trefethen(path: str, dtype=np.float64):
    """C1: a MatrixMarket file such as the reference's examples/Trefethen_20b.mtx (symmetric storage expanded)."""
    import scipy.io

    A = sp.csc_matrix(scipy.io.mmread(path)).astype(dtype)
    A.sort_indices()
    return A
