"""Synthetic count matrices of the shapes BASELINE.json names (SURVEY.md §8d recipe).

Poisson-lognormal counts with planted cell types: gene base rates exp(N(-1.9, 1.6^2)), K
types, each gene differentially expressed in a type with probability 0.10 (log fold change
N(0,1)), profiles on the simplex, uniform type assignment, depth exp(N(log D, 0.35^2)),
counts ~ Poisson(depth_i * profile).  D is bisected so the zero fraction hits ``sparsity``.
Gene rates are floored so QC (>=15 cells/gene, >=200 genes/cell at the benchmark shapes,
/root/reference/src/scLENS.jl:160-162) is the identity.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def _profiles(M: int, K: int, rng: np.random.Generator, de_prob: float, lfc_sd: float) -> np.ndarray:
    base = np.exp(rng.normal(-1.9, 1.6, size=M))
    de = rng.random((K, M)) < de_prob
    lfc = rng.normal(0.0, lfc_sd, size=(K, M)) * de
    prof = base[None, :] * np.exp(lfc)
    return prof / prof.sum(axis=1, keepdims=True)


def make_counts(N: int, M: int, seed: int = 0, K: int = 8, sparsity: float = 0.92,
                min_cells_per_gene: int = 40, block: int = 2048,
                de_prob: float = 0.10, lfc_sd: float = 1.0) -> sp.csc_matrix:
    """N cells x M genes, Float32 integer-valued counts, canonical CSC."""
    rng = np.random.default_rng(seed)
    prof = _profiles(M, K, rng, de_prob, lfc_sd)
    types = rng.integers(0, K, size=N)
    ldepth = rng.normal(0.0, 0.35, size=N)

    sub = rng.choice(N, size=min(N, 512), replace=False)

    def zero_frac(D, p):
        lam = (D * np.exp(ldepth[sub]))[:, None] * p[types[sub]]
        return float(np.exp(-lam).mean())

    def floored(D):
        # floor so that every gene is expected in >= min_cells_per_gene cells
        pmin = -np.log1p(-min(0.5, min_cells_per_gene / N)) / D
        p = np.maximum(prof, pmin)
        return p / p.sum(axis=1, keepdims=True)

    lo, hi = 1.0, 1e7
    for _ in range(60):
        mid = np.sqrt(lo * hi)
        if zero_frac(mid, floored(mid)) > sparsity:
            lo = mid
        else:
            hi = mid
    D = np.sqrt(lo * hi)
    p = floored(D)

    rows, cols, vals = [], [], []
    for s in range(0, N, block):
        e = min(N, s + block)
        lam = (D * np.exp(ldepth[s:e]))[:, None] * p[types[s:e]]
        c = rng.poisson(lam)
        r_, c_ = np.nonzero(c)
        rows.append((r_ + s).astype(np.int32))
        cols.append(c_.astype(np.int32))
        vals.append(c[r_, c_].astype(np.float32))
    X = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(N, M), dtype=np.float32).tocsc()
    X.sort_indices()
    # guarantee no empty gene / cell (tiny shapes): put a single count where needed
    empty_g = np.nonzero(np.diff(X.indptr) == 0)[0]
    empty_c = np.nonzero(np.asarray((X != 0).sum(axis=1)).ravel() == 0)[0]
    if len(empty_g) or len(empty_c):
        X = X.tolil()
        for g in empty_g:
            X[rng.integers(0, N), g] = 1.0
        for c in empty_c:
            X[c, rng.integers(0, M)] = 1.0
        X = X.tocsc()
        X.sort_indices()
    return X


def qc_is_identity(X: sp.csc_matrix, min_genes_per_cell=200, min_cells_per_gene=15) -> bool:
    cells_per_gene = np.diff(X.indptr)
    genes_per_cell = np.asarray((X != 0).sum(axis=1)).ravel()
    return bool((cells_per_gene >= min_cells_per_gene).all() and
                (genes_per_cell >= min_genes_per_cell).all())
