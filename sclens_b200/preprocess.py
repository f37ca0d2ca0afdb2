"""Host-side QC mirror of scLENS.preprocess (/root/reference/src/scLENS.jl:160-236).

QC is one-off host integer work in the reference and stays host work here (SURVEY.md §2:
out of scope for acceleration, but its index outputs are part of the parity contract, row
P0): sparse column/row reductions in scipy, Float32 sums and means as in the reference."""
from __future__ import annotations

import re

import numpy as np
import scipy.sparse as sp


def preprocess(tmp_df, min_tp_c=0, min_tp_g=0, max_tp_c=np.inf, max_tp_g=np.inf, min_genes_per_cell=200,
               max_genes_per_cell=0, min_cells_per_gene=15, mito_percent=5.0, ribo_percent=0.0, verbose=True):
    """DataFrame (col 0 'cell', then genes) -> filtered DataFrame with sparse Float32 gene
    columns sorted by mean expression, or None ("There is no high quality cells and genes")."""
    import pandas as pd
    cell_name = tmp_df.iloc[:, 0].to_numpy()
    gene_name = np.asarray(tmp_df.columns[1:], dtype=object)
    body = tmp_df.iloc[:, 1:]
    if all(isinstance(t, pd.SparseDtype) for t in body.dtypes):
        X = sp.csc_matrix(body.sparse.to_coo(), dtype=np.float32)
    else:
        X = sp.csc_matrix(body.to_numpy(dtype=np.float32))
    X.eliminate_zeros()
    if verbose:
        print("Inp_spec")
        print(f"data size: {tmp_df.shape}, sparsity: {1 - X.nnz / (X.shape[0] * X.shape[1])}")
    res = qc_indices(X, gene_name, min_tp_c, min_tp_g, max_tp_c, max_tp_g, min_genes_per_cell, max_genes_per_cell,
                     min_cells_per_gene, mito_percent, ribo_percent)
    if res is None:
        if verbose:
            print("There is no high quality cells and genes")
        return None
    fc_idx, gene_idx = res
    out = X[fc_idx][:, gene_idx].tocsc()
    o_df = pd.DataFrame.sparse.from_spmatrix(out, columns=gene_name[gene_idx])
    o_df.insert(0, "cell", cell_name[fc_idx])
    if verbose:
        print(f"After filtering>> data size: {o_df.shape}, sparsity: {1 - out.nnz / (out.shape[0] * out.shape[1])}")
    return o_df


def qc_indices(X: sp.csc_matrix, gene_name, min_tp_c=0, min_tp_g=0, max_tp_c=np.inf, max_tp_g=np.inf,
               min_genes_per_cell=200, max_genes_per_cell=0, min_cells_per_gene=15, mito_percent=5.0,
               ribo_percent=0.0):
    """(fc_idx mask over cells, gene positions in output order) of :183-225."""
    X = sp.csc_matrix(X, dtype=np.float32)
    f32 = np.float32
    cells_per_gene = np.diff(X.indptr)
    gene_sum = np.asarray(X.sum(axis=0), dtype=f32).ravel()
    fg = (gene_sum > min_tp_g) & (gene_sum < max_tp_g) & (cells_per_gene >= min_cells_per_gene)
    Xr = X.tocsr()
    genes_per_cell = np.diff(Xr.indptr)
    cell_sum = np.asarray(Xr.sum(axis=1), dtype=f32).ravel()
    keep = (cell_sum > min_tp_c) & (cell_sum < max_tp_c) & (genes_per_cell >= min_genes_per_cell)

    def frac_below(pattern, percent):
        if percent == 0:
            return np.ones(X.shape[0], dtype=bool)
        sel = np.array([re.match(pattern, str(g), flags=re.I) is not None for g in gene_name])
        part = np.asarray(X[:, sel].sum(axis=1), dtype=f32).ravel() if sel.any() else np.zeros(X.shape[0], f32)
        with np.errstate(invalid="ignore", divide="ignore"):
            return (part / cell_sum).astype(np.float64) < percent / 100

    keep &= frac_below(r"^mt-.", mito_percent) & frac_below(r"^RP[SL].", ribo_percent)
    if max_genes_per_cell != 0:
        keep &= genes_per_cell < max_genes_per_cell
    if not (keep.any() and fg.any()):
        return None
    sub = X[keep][:, fg].tocsc()
    col_sum = np.asarray(sub.sum(axis=0), dtype=f32).ravel()
    nn = col_sum != 0
    mean_ = (col_sum[nn] / f32(sub.shape[0])).astype(f32)
    order = np.argsort(mean_, kind="stable")
    return keep, np.nonzero(fg)[0][nn][order]
