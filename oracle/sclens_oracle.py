"""CPU oracle for the scLENS.sclens() signal-detection path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``sclens_b200/`` may import this module;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs do.

PARITY UNPINNED.  The reference (/root/reference, Julia 1.11.1 + CUDA.jl 5.5.2) has no
tests, fixtures or golden vectors for this path, and neither Julia nor the missing input
blob (data/Z8eq.csv.gz) exist in this image, so the restatement below cannot be run
against the reference itself.  It is pinned only by soft known answers (see
tests/test_oracle.py): structural checks on the one committed output out/pca.csv, and the
algebraic identities every function must satisfy.

Each function restates one reference function of ``src/scLENS.jl`` (cited as ``:NNN``).
All random draws are explicit inputs/outputs ("draw injection"): the
reference draws from Julia's unseeded task-local RNG, so oracle and CUDA path are fed the
same ``Draws`` bundle.  Indices are 0-based here; the reference is 1-based.

Two numeric modes mirror the reference's ``device_`` keyword:
  * ``"cpu"``     - Float64 Gram (syrk) and Float64 LAPACK eigensolve   (:345-359, :384)
  * ``"gpu-ref"`` - operands rounded to Float32, Float32 GEMM + eigensolve (:335-343, :377)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp


# --------------------------------------------------------------------------------------
# ingest
# --------------------------------------------------------------------------------------
def as_csc(X) -> sp.csc_matrix:
    """df2sparr (:90-120): SparseMatrixCSC{Float32,UInt32}, canonical (col,row) order."""
    X = sp.csc_matrix(X, dtype=np.float32)
    X.sum_duplicates()
    X.sort_indices()
    return X


def findnz(X: sp.csc_matrix) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """findnz (:664): triplets in CSC order (column-major, rows ascending inside a column)."""
    X = as_csc(X)
    col = np.repeat(np.arange(X.shape[1], dtype=np.int64), np.diff(X.indptr))
    return X.indices.astype(np.int64), col, X.data.copy()


# --------------------------------------------------------------------------------------
# draws
# --------------------------------------------------------------------------------------
@dataclass
class Draws:
    """Every random quantity sclens() consumes, in consumption order."""
    z_idx1: np.ndarray = None            # zero-candidate rows  (:668-673), 0-based
    z_idx2: np.ndarray = None            # zero-candidate cols
    null_perm: np.ndarray = None         # shuffle(nz_val) as a gather permutation (:275)
    null_rows: np.ndarray = None         # row_i, aligned with the CSC order of nz_col (:247)
    p_th: float = None                   # noise baseline (:709-712)
    search_sple: List[np.ndarray] = field(default_factory=list)   # per search step (:731)
    perturb_sple: List[np.ndarray] = field(default_factory=list)  # per replicate   (:772)


def draw_zero_candidates(X: sp.csc_matrix, rng: np.random.Generator):
    """:668-673.  nnz uniform (i,j) pairs, minus the non-zero set, unique in order of first
    occurrence (Julia ``setdiff``)."""
    X = as_csc(X)
    N, M = X.shape
    nnz = X.nnz
    i = rng.integers(0, N, size=nnz, dtype=np.int64)
    j = rng.integers(0, M, size=nnz, dtype=np.int64)
    key = j * N + i
    nz_row, nz_col, _ = findnz(X)
    nz_key = nz_col * N + nz_row
    keep = ~np.isin(key, nz_key)
    key = key[keep]
    _, first = np.unique(key, return_index=True)
    first.sort()
    key = key[first]
    return (key % N).astype(np.uint32), (key // N).astype(np.uint32)


def draw_null(X: sp.csc_matrix, rng: np.random.Generator, gene_order=None):
    """Draws of random_nz/_random_matrix (:261-289, :239-248).

    ``null_perm`` is the value shuffle (:275).  ``null_rows`` concatenates, gene by gene in
    ``gene_order``, a without-replacement sample of rows of size count(gene) (:244-247).
    Julia iterates a ``Dict`` (hash order, not reproducible outside Julia, Appendix A3);
    ``gene_order=None`` uses sorted order, which lines the chunks up with the column-sorted
    ``nz_col`` (the evident intent).  Any other order reproduces the reference's accidental
    misalignment, under which a column can receive duplicate rows that ``sparse`` sums."""
    X = as_csc(X)
    N, M = X.shape
    counts = np.diff(X.indptr)
    perm = rng.permutation(X.nnz).astype(np.uint32)
    order = np.arange(M) if gene_order is None else np.asarray(gene_order)
    chunks = [rng.choice(N, size=int(counts[g]), replace=False) for g in order if counts[g] > 0]
    rows = np.concatenate(chunks).astype(np.uint32) if chunks else np.zeros(0, np.uint32)
    return perm, rows


def build_null(X: sp.csc_matrix, perm: np.ndarray, rows: np.ndarray) -> sp.csc_matrix:
    """sparse(row_i, nz_col, shuffled nz_val) (:248, :275): duplicates summed, size =
    (max(row_i)+1, max(nz_col)+1) exactly as ``sparse(I,J,V)`` without dimensions."""
    X = as_csc(X)
    _, nz_col, nz_val = findnz(X)
    vals = nz_val[perm.astype(np.int64)]
    shape = (int(rows.max()) + 1, int(nz_col.max()) + 1) if len(rows) else (0, 0)
    R = sp.coo_matrix((vals, (rows.astype(np.int64), nz_col)), shape=shape, dtype=np.float32)
    R = R.tocsc()          # sums duplicates
    R.sort_indices()
    return R


def draw_noise_baseline(nm: int, rng: np.random.Generator, n_rep: int = 5000) -> float:
    """:709-712  p_th = mean over 5000 draws of max |N(0, 1/nm)| over nm samples."""
    sd = math.sqrt(1.0 / nm)
    acc = 0.0
    for _ in range(n_rep):
        acc += float(np.max(np.abs(rng.normal(0.0, sd, size=nm))))
    return acc / n_rep


def julia_round(x: float) -> int:
    """Int(round(x)): ties to even (Appendix A12)."""
    return int(np.rint(x))


# --------------------------------------------------------------------------------------
# normalisation
# --------------------------------------------------------------------------------------
def normalize_main(X: sp.csc_matrix):
    """The inline Float64 normalisation of sclens (:677-696).  Returns (dense N x M f64,
    rec_vals)."""
    X = as_csc(X)
    N, M = X.shape
    rec = {}
    tgc = np.asarray(X.sum(axis=1), dtype=np.float32).ravel().astype(np.float64)   # :678
    rec["TGC"] = tgc
    n_mat = sp.diags(1.0 / tgc) @ X.astype(np.float64)                              # :679
    mat2 = n_mat.tocsc()
    mat2.data = np.log1p(mat2.data)                                                 # :681
    mean2 = np.asarray(mat2.mean(axis=0)).ravel()                                   # :682
    dense2 = mat2.toarray()
    std2 = dense2.std(axis=0, ddof=1)                                               # :683 corrected
    rec["mat2_mean"] = mean2.reshape(1, M)
    rec["mat2_std"] = std2.reshape(1, M)
    mat3 = dense2 * (1.0 / std2)[None, :]                                           # :685
    mup = mat3.mean(axis=0)                                                         # :686
    l2X = np.sqrt((mat3 ** 2).sum(axis=1))                                          # :688
    l2mu = np.linalg.norm(mup)                                                      # :689
    l2norm = np.sqrt(l2X ** 2 - 2.0 * (mat3 @ mup) + l2mu ** 2)                     # :690
    rec["norm_tgc"] = l2norm
    mat4 = (mat3 - mup[None, :]) / (l2norm / l2norm.mean())[:, None]                # :693
    cent = mat4.mean(axis=0)                                                        # :695
    rec["cent_"] = cent.reshape(1, M)
    return mat4 - cent[None, :], rec


def logn_scale_pre_scale(X: sp.csc_matrix) -> np.ndarray:
    """logn_scale(pre_scale(X)) for centering=="mean" (:650-652): proj_l (:607, Float32),
    log1p (Float32), zscore_with_l2 (:596-605; Float32 std, Float64 from the ``1.`` literal
    on), scaled_gdata "cent" (:300-305)."""
    X = as_csc(X)
    N, M = X.shape
    rs = np.asarray(X.sum(axis=1), dtype=np.float32).ravel()
    inv = (np.float32(1.0) / rs).astype(np.float32)
    P = X.tocoo()
    y = np.log1p((inv[P.row] * P.data).astype(np.float32)).astype(np.float32)
    dense = np.zeros((N, M), dtype=np.float32)
    dense[P.row, P.col] = y
    std_ = dense.astype(np.float64).std(axis=0, ddof=1).astype(np.float32)          # :597
    Xn = dense.astype(np.float64) * (1.0 / std_.astype(np.float64))[None, :]        # :598
    mu = Xn.mean(axis=0)                                                            # :599
    l2X = np.sqrt((Xn ** 2).sum(axis=1))                                            # :601
    l2mu = np.linalg.norm(mu)                                                       # :602
    l2n = np.sqrt(l2X ** 2 - 2.0 * (Xn @ mu) + l2mu ** 2)                           # :603
    out = (Xn - mu[None, :]) / (l2n / l2n.mean())[:, None]                          # :604
    return out - out.mean(axis=0)[None, :]                                          # :305


def logn_scale_pre_scale_median(X: sp.csc_matrix) -> np.ndarray:
    """logn_scale(pre_scale(X)) for centering=="median" (:653-654):
    norm_l(scaled_gdata(Matrix{Float32}(log1p.(proj_l(X))), position_="median")) - dense Float32 throughout:
    per-gene median over ALL cells (`mapslices(median, X, dims=1)` :298) and corrected std (:307), (X - median) / std
    (:326), then every row rescaled to the mean row norm (norm_l :608).  No re-centring afterwards."""
    X = as_csc(X)
    N, M = X.shape
    rs = np.asarray(X.sum(axis=1), dtype=np.float32).ravel()
    inv = (np.float32(1.0) / rs).astype(np.float32)
    P = X.tocoo()
    dense = np.zeros((N, M), dtype=np.float32)
    dense[P.row, P.col] = np.log1p((inv[P.row] * P.data).astype(np.float32)).astype(np.float32)
    med = np.median(dense, axis=0).astype(np.float32)                               # :298
    std_ = dense.astype(np.float64).std(axis=0, ddof=1).astype(np.float32)          # :307
    W = ((dense - med[None, :]) / std_[None, :]).astype(np.float32)                 # :326
    l2 = np.sqrt((W.astype(np.float64) ** 2).sum(axis=1)).astype(np.float32)        # :608
    return (W / l2[:, None] * np.float32(l2.astype(np.float64).mean())).astype(np.float32)


# --------------------------------------------------------------------------------------
# Gram + eigen
# --------------------------------------------------------------------------------------
def wishart_matrix(X: np.ndarray, dims: int, mode: str) -> np.ndarray:
    """_wishart_matrix (:332-361).  BOTH branches divide by size(X,2) (Appendix A9)."""
    if mode == "gpu-ref":
        Xf = np.asarray(X, dtype=np.float32)
        Y = Xf.T @ Xf if dims == 2 else Xf @ Xf.T
        return (Y / np.float32(X.shape[1])).astype(np.float32)
    Xd = np.asarray(X, dtype=np.float64)
    Y = Xd.T @ Xd if dims == 2 else Xd @ Xd.T
    return Y / X.shape[1]


def get_eigen(Y: np.ndarray, mode: str, eigvals_only: bool = False):
    """_get_eigen (:375-387): all eigenpairs, ascending."""
    Y = np.asarray(Y, dtype=np.float32 if mode == "gpu-ref" else np.float64)
    if eigvals_only:
        return sla.eigh(Y, eigvals_only=True, driver="evd", check_finite=False), None
    return sla.eigh(Y, driver="evd", check_finite=False)


def corr_mat(X: np.ndarray, Y: np.ndarray, mode: str) -> np.ndarray:
    """corr_mat (:363-373)."""
    if mode == "gpu-ref":
        return np.asarray(X, np.float32).T @ np.asarray(Y, np.float32)
    return X.T @ Y


# --------------------------------------------------------------------------------------
# Marchenko-Pastur / Tracy-Widom
# --------------------------------------------------------------------------------------
def mp_parameters(L: np.ndarray) -> dict:
    """_mp_parameters (:390-408), in the dtype of L."""
    L = np.asarray(L)
    m1 = L.mean(dtype=L.dtype)
    m2 = (L * L).mean(dtype=L.dtype)
    gamma = m2 / (m1 * m1) - 1
    t = L.dtype.type
    sg = np.sqrt(gamma)
    return {"moment_1": m1, "moment_2": m2, "gamma": gamma,
            "b_plus": m1 * (t(1) + sg) ** 2, "b_minus": m1 * (t(1) - sg) ** 2,
            "s": m1, "sigma": m2}


def mp_calculation(L: np.ndarray, Lr: np.ndarray, eta=1, eps=1e-6, max_iter=10000):
    """_mp_calculation (:424-459): fixed-point refinement of the MP edges."""
    L = np.asarray(L)
    p = mp_parameters(Lr)
    b_plus, b_minus = p["b_plus"], p["b_minus"]
    Lu = L[(b_minus < L) & (L < b_plus)]
    q = mp_parameters(Lu)
    new_b_plus, new_b_minus = q["b_plus"], q["b_minus"]
    it = 0
    while True:
        loss = (1 - new_b_plus / b_plus) ** 2
        it += 1
        if loss <= eps or it == max_iter:
            break
        gradient = new_b_plus - b_plus
        new_b_plus = b_plus + eta * gradient
        Lu = L[(new_b_minus < L) & (L < new_b_plus)]
        b_plus, b_minus = new_b_plus, new_b_minus
        q = mp_parameters(Lu)
        new_b_plus, new_b_minus = q["b_plus"], q["b_minus"]
    return L[(new_b_minus < L) & (L < new_b_plus)], new_b_plus, new_b_minus, it


def tw(L: np.ndarray, L_mp: np.ndarray):
    """_tw (:461-467)."""
    gamma = mp_parameters(L_mp)["gamma"]
    p = len(L) / gamma
    sigma = 1 / p ** (2 / 3) * gamma ** (5 / 6) * (1 + np.sqrt(gamma)) ** (4 / 3)
    lam = np.asarray(L_mp).mean(dtype=np.asarray(L_mp).dtype) * (1 + np.sqrt(gamma)) ** 2 + sigma
    return lam, gamma, p, sigma


def mp_pdf(x: np.ndarray, L: np.ndarray) -> np.ndarray:
    """_mp_pdf/_marchenko_pastur (:411-422)."""
    y = mp_parameters(np.asarray(L, dtype=np.float64))
    x = np.asarray(x, dtype=np.float64)
    out = np.zeros_like(x)
    ok = (y["b_minus"] < x) & (x < y["b_plus"])
    xo = x[ok]
    out[ok] = np.sqrt((y["b_plus"] - xo) * (xo - y["b_minus"])) / (2 * y["s"] * np.pi * y["gamma"] * xo)
    return out


def mp_check(test_L: np.ndarray, p_val: float = 0.05) -> dict:
    """mp_check (:469-487).  histcounts is NaNStatistics (Appendix A18): bins [e_i, e_i+1)."""
    test_L = np.asarray(test_L, dtype=np.float64)
    bin_x = np.linspace(test_L.min() - 1, test_L.max() + 1, 100)
    count, _ = np.histogram(test_L, bins=bin_x)
    cdf = np.cumsum(count / count.sum())
    centers = (bin_x[1:] + bin_x[:-1]) / 2
    c2 = np.cumsum(mp_pdf(centers, test_L))
    nc2 = c2 / c2.max()
    D = float(np.max(np.abs(cdf - nc2)))
    c_a = math.sqrt(-0.5 * math.log(p_val))
    m = n = len(cdf)
    return {"ks_static": D, "pass": bool(D <= c_a * math.sqrt((m + n) / m / n))}


# --------------------------------------------------------------------------------------
# eigvec / sigev
# --------------------------------------------------------------------------------------
def _colnorm(A: np.ndarray) -> np.ndarray:
    return A / np.linalg.norm(A, axis=0, keepdims=True)


def positive_mask(L: np.ndarray) -> np.ndarray:
    """``L .> 0`` (:495, :515) read in exact arithmetic.  A column-centred N x M matrix with
    N <= M has rank <= N-1, so its cell Gram has one eigenvalue that is exactly zero; in floating
    point it comes out as +-1e-16 (Float64) or +-1e-7 (Float32) and the reference keeps or drops
    that vector by the sign of rounding noise.  Oracle and CUDA path both use the exact-arithmetic
    answer: an eigenvalue below eps_rel * max(L) is not positive (deviations)."""
    L = np.asarray(L)
    eps_rel = 1e-5 if L.dtype == np.float32 else 1e-10
    return L > eps_rel * L.max()


def get_eigvec(X: np.ndarray, mode: str, keep_null: bool = False):
    """get_eigvec (:489-524): keep L>0, sort descending (stable), back-project if N>M.

    ``keep_null=True`` keeps the complete eigenbasis.  It is used for the reference basis Vr2 of the
    sparsity search only (:717-721): there the reference's ``L .> 0`` keeps each exactly-null
    direction (real data has several: z_data_785's binarised matrix has rank N-5) by the sign of
    rounding noise, and every dropped one makes the matching small-eigenvalue vectors of the
    perturbed matrices look maximally delocalised (d ~ 0.01), which stops the search after five
    steps.  Keeping them all is the outcome the reference produces when those eigenvalues round
    positive, and is the only one under which d_arr measures what it is meant to measure."""
    N, M = X.shape
    if keep_null:
        assert N <= M, "the search passes the transposed matrix when N > M (:718)"
        Y = wishart_matrix(X, 1, mode)
        L, V = get_eigen(Y, mode)
        idx = np.argsort(-L, kind="stable")
        return L[idx], V[:, idx]
    if N > M:
        Y = wishart_matrix(X, 2, mode)
        L, V = get_eigen(Y, mode)
        pos = positive_mask(L)
        L, V = L[pos], V[:, pos]
        idx = np.argsort(-L, kind="stable")
        nL, nVs = L[idx], V[:, idx]
        mul_X = nVs * np.sqrt(1.0 / nL)[None, :]
        if mode == "gpu-ref":
            prod = np.asarray(X, np.float32) @ mul_X.astype(np.float32)
        else:
            prod = X @ mul_X
        return nL, _colnorm(prod)
    Y = wishart_matrix(X, 1, mode)
    L, V = get_eigen(Y, mode)
    pos = positive_mask(L)
    L, V = L[pos], V[:, pos]
    idx = np.argsort(-L, kind="stable")
    return L[idx], V[:, idx]


def get_sigev(X: np.ndarray, Xr: np.ndarray, mode: str):
    """get_sigev (:526-594).  Returns nL, nV, L, L_mp, lambda_c, b_min, b_plus, n_mp_iter.
    The reference's noise-vector back-projection (:557-564) is dead work (its result is
    never read after :704) and is not restated."""
    n, m = X.shape
    dims = 2 if n > m else 1
    Y = wishart_matrix(X, dims, mode)
    L, V = get_eigen(Y, mode)
    Yr = wishart_matrix(Xr, dims, mode)
    Lr, _ = get_eigen(Yr, mode, eigvals_only=True)
    L_mp, b_plus, b_min, n_it = mp_calculation(L, Lr[:-1])                           # :537
    lambda_c, _, _, _ = tw(L, L_mp)                                                  # :538
    sel = L > lambda_c                                                               # :541 strict
    sel_L, sel_V = L[sel], V[:, sel]
    idx = np.argsort(-sel_L, kind="stable")
    nL, nVs = sel_L[idx], sel_V[:, idx]
    if n > m:
        mul_X = nVs * np.sqrt(1.0 / nL)[None, :]                                     # :556
        if mode == "gpu-ref":
            prod = np.asarray(X, np.float32) @ mul_X.astype(np.float32)
        else:
            prod = X @ mul_X
        nVs = _colnorm(prod) if prod.shape[1] else prod                              # :558
    return nL, nVs, L, L_mp, lambda_c, b_min, b_plus, n_it


# --------------------------------------------------------------------------------------
# perturbed matrices
# --------------------------------------------------------------------------------------
def perturbed_matrix(X: sp.csc_matrix, z1, z2, sple, binarise: bool) -> sp.csc_matrix:
    """sparse(vcat(nz_row, z_idx1[sple]), vcat(nz_col, z_idx2[sple]), vcat(vals, ones), N, M)
    (:735 binarised, :774 counts)."""
    X = as_csc(X)
    N, M = X.shape
    r, c, v = findnz(X)
    if binarise:
        v = np.ones_like(v)
    sple = np.asarray(sple, dtype=np.int64)
    rr = np.concatenate([r, z1[sple].astype(np.int64)])
    cc = np.concatenate([c, z2[sple].astype(np.int64)])
    vv = np.concatenate([v, np.ones(len(sple), np.float32)])
    P = sp.coo_matrix((vv, (rr, cc)), shape=(N, M), dtype=np.float32).tocsc()
    P.sort_indices()
    return P


def robustness_scores(nV: np.ndarray, nV_set: List[np.ndarray], th: float):
    """:786-806."""
    th_ = math.cos(math.radians(th))
    n_perturb = len(nV_set)
    a_b = np.stack([np.argmax(np.abs(nV.T @ j), axis=1) for j in nV_set], axis=1)     # :788
    sub = [nV_set[s][:, a_b[:, s]] for s in range(n_perturb)]                        # :790
    b_vec = []
    for i in range(n_perturb):
        for j in range(i + 1, n_perturb):
            b_vec.append(np.max(np.abs(sub[i].T @ sub[j]), axis=1))                   # :793
    b_ = np.stack(b_vec, axis=1)
    q1 = np.quantile(b_, 0.25, axis=1)
    q3 = np.quantile(b_, 0.75, axis=1)
    iqr = q3 - q1
    m_score = np.zeros(b_.shape[0])
    sd_score = np.zeros(b_.shape[0])
    for s in range(b_.shape[0]):
        row = b_[s]
        f = row[(q1[s] - 1.5 * iqr[s] <= row) & (row <= q3[s] + 1.5 * iqr[s])]        # :800
        m_score[s] = np.median(f)
        sd_score[s] = np.std(f, ddof=1) if len(f) > 1 else np.nan
    sig_id = np.nonzero(m_score > th_)[0]                                             # :806
    return {"b_": b_, "rob_score": m_score, "m_scores": m_score, "sd_scores": sd_score,
            "a_b": a_b}, sig_id


# --------------------------------------------------------------------------------------
# driver
# --------------------------------------------------------------------------------------
def sclens(X, draws: Optional[Draws] = None, rng: Optional[np.random.Generator] = None,
           mode: str = "cpu", th: float = 60, p_step: float = 0.001, n_perturb: int = 20,
           null_gene_order=None, n_baseline: int = 5000, verbose: bool = False, centering: str = "mean"):
    """sclens (:649-832), centering="mean" (default) or "median" (:653-654: no rec_vals, every matrix through
    logn_scale_pre_scale_median).  ``mode`` = "cpu" | "gpu-ref".

    Draws missing from ``draws`` are generated from ``rng`` and recorded, so the returned
    bundle can be replayed through the CUDA path.  Returns (result dict, Draws, info)."""
    X = as_csc(X)
    N, M = X.shape
    draws = draws or Draws()
    rng = rng or np.random.default_rng(0)
    info = {}
    nz_row, nz_col, nz_val = findnz(X)
    if draws.z_idx1 is None:
        draws.z_idx1, draws.z_idx2 = draw_zero_candidates(X, rng)
    z1, z2 = draws.z_idx1, draws.z_idx2

    if centering == "median":
        logn = logn_scale_pre_scale_median
        scaled_X, rec_vals = logn(X).astype(np.float64), {}                           # :696-698 (the else branch)
    else:
        logn = logn_scale_pre_scale
        scaled_X, rec_vals = normalize_main(X)                                        # :677-696
    if draws.null_perm is None:
        draws.null_perm, draws.null_rows = draw_null(X, rng, null_gene_order)
    X_r = build_null(X, draws.null_perm, draws.null_rows)                             # :701
    info["null"] = X_r
    nL, nV, L, L_mp, lambda_c, b_min, b_plus, n_it = get_sigev(
        scaled_X, logn(X_r), mode)                                    # :704
    info.update(b_min=b_min, b_plus=b_plus, mp_iters=n_it)
    mpc = mp_check(L_mp)                                                              # :706
    nm = min(N, M)
    if draws.p_th is None:
        draws.p_th = draw_noise_baseline(nm, rng, n_baseline)
    p_th = draws.p_th

    res_base = {"L": L, "L_mp": L_mp, "lambda": lambda_c}
    min_s = nV.shape[1]
    if min_s == 0:                                                                    # :780-784
        return res_base, draws, info

    # sparsity search (:715-762)
    p_ = 0.999
    Xb = X.copy()
    Xb.data = np.ones_like(Xb.data)
    nb = logn(Xb)
    Vr2 = get_eigvec(nb.T if N > M else nb, mode, keep_null=True)[1]                  # :717-721
    n_2 = julia_round(Vr2.shape[1] / 2)                                               # :722
    tank = np.zeros((5, 0))
    tank_n = 5
    step = 0
    trace = []
    while True:
        nnzidx = julia_round((1 - p_) * M * N)                                        # :726
        if len(z1) < nnzidx:
            p_ += p_step
            break
        if step < len(draws.search_sple):
            sple = draws.search_sple[step]
        else:
            sple = rng.choice(len(z1), size=nnzidx, replace=False).astype(np.uint32)  # :731
            draws.search_sple.append(sple)
        assert len(sple) == nnzidx
        Pb = perturbed_matrix(X, z1, z2, sple, binarise=True)
        npb = logn(Pb)
        nV_2 = get_eigvec(npb.T if N > M else npb, mode)[1]                           # :733-739
        W = nV_2[:, nV_2.shape[1] - n_2 - 1:]                                         # end-n_2:end
        d_arr = np.nanmax(np.abs(corr_mat(Vr2, W, mode)), axis=0)                     # :742
        tmp_A = np.sort(d_arr)
        tank = np.hstack([tank, tmp_A[:5].reshape(5, 1).astype(np.float64)])          # :748
        ppj = tank[1, :] if tank.shape[1] < tank_n else tank[1, -tank_n:]
        trace.append((p_, nnzidx, float(ppj[-1])))
        if verbose:
            print(p_, float(ppj[-1]))
        step += 1
        if (np.sum(ppj < p_th) > tank_n - 1) or (p_ < 0.9):                           # :756
            p_ += (tank_n - 1) * p_step
            break
        p_ -= p_step
    info.update(p_sel=p_, n_search=step, search_trace=trace)

    # perturbations (:767-778)
    min_pc = int(math.ceil(min_s * 1.5))
    nV_set, nL_set = [], []
    n_add = julia_round((1 - p_) * M * N)                                             # :772
    for r in range(n_perturb):
        if r < len(draws.perturb_sple):
            sple = draws.perturb_sple[r]
        else:
            sple = rng.choice(len(z1), size=n_add, replace=False).astype(np.uint32)
            draws.perturb_sple.append(sple)
        tmp_X = perturbed_matrix(X, z1, z2, sple, binarise=False)                     # :774
        tL, tV = get_eigvec(logn(tmp_X), mode)                        # :775
        k = min(min_pc, tV.shape[1])
        nV_set.append(np.asarray(tV[:, :k], dtype=np.float64))
        nL_set.append(tL[:k])
    info.update(nV_set=nV_set, nL_set=nL_set, min_pc=min_pc, n_add=n_add)

    rob, sig_id = robustness_scores(np.asarray(nV, np.float64), nV_set, th)           # :786-806
    sq = np.sqrt(np.asarray(nL, np.float64))
    Xout0 = nV * sq[None, :]                                                          # :810
    Xout1 = nV[:, sig_id] * sq[sig_id][None, :]                                       # :811
    if mode == "gpu-ref":
        g = np.asarray(nV, np.float32).T @ scaled_X.astype(np.float32)
    else:
        g = nV.T @ scaled_X
    gene_basis = (1.0 / sq)[:, None] * g / math.sqrt(M)                               # :813-819
    res = dict(res_base)
    res.update(pca=Xout0, pca_n1=Xout1, sig_id=sig_id, robustness_scores=rob,
               signal_evec=nV, signal_ev=nL, gene_basis=gene_basis,
               **{"pass": mpc["pass"]}, ks_static=mpc["ks_static"], rec_vals=rec_vals)
    return res, draws, info


# --------------------------------------------------------------------------------------
# downstream consumer restated ahead of its device implementation
# --------------------------------------------------------------------------------------
def get_denoised(result: dict, mode: str = "gpu-ref") -> np.ndarray:
    """get_denoised_df (:889-931) on a result dict of :func:`sclens` (or of sclens_b200.sclens): the
    robust-signal reconstruction pca_n1 * gene_basis[sig_id, :] * sqrt(M) (:890-911, a Float32 GEMM on
    both reference paths) pushed back through the recorded normalisation (:913-926):
    + cent_, * norm_tgc / mean(norm_tgc), * mat2_std + mat2_mean, exp(.) - 1 clamped at 0, rows scaled to
    sum 1, * mean(TGC).  Returns the N x M Float64 matrix of the output DataFrame (without the cell column)."""
    gene_basis = np.asarray(result["gene_basis"])
    sig_id = np.asarray(result["sig_id"], dtype=np.int64)
    g_mat = gene_basis[sig_id, :]                                                     # :890
    pca_n1 = result["pca_n1"]
    if hasattr(pca_n1, "iloc"):                                                       # DataFrame(cell, x1..xr)
        pca_n1 = pca_n1.iloc[:, 1:].to_numpy()
    Xout0 = np.asarray(pca_n1, dtype=np.float32)                                      # :891
    M = gene_basis.shape[1]
    sqrtM = math.sqrt(M)                                                              # sqrt(size(gene_basis, 2))
    # the product lands in a Float32 array on both paths; `.* sqrt(size(...))` multiplies by a Float64 scalar, so
    # d_mean is the Float64 product of the Float32 GEMM result and sqrt(M)
    if mode == "gpu-ref":                                                             # :893-896 (cu() is Float32)
        d_mean = (Xout0 @ np.asarray(g_mat, dtype=np.float32)).astype(np.float32).astype(np.float64) * sqrtM
    else:                                                                             # :903-905: mul! into a Float32 array
        d_mean = (Xout0.astype(np.float64) @ np.asarray(g_mat, dtype=np.float64)).astype(np.float32).astype(np.float64) * sqrtM
    rec = result["rec_vals"]
    TGC = np.asarray(rec["TGC"], dtype=np.float64).ravel()
    mat2_mean = np.asarray(rec["mat2_mean"], dtype=np.float64).reshape(1, -1)
    mat2_std = np.asarray(rec["mat2_std"], dtype=np.float64).reshape(1, -1)
    norm_tgc = np.asarray(rec["norm_tgc"], dtype=np.float64).ravel()
    cent_ = np.asarray(rec["cent_"], dtype=np.float64).reshape(1, -1)
    r_mat1 = d_mean + cent_                                        # :921
    r_mat2 = r_mat1 * (norm_tgc / norm_tgc.mean())[:, None]                           # :922
    r_mat3 = r_mat2 * mat2_std + mat2_mean                                            # :923
    r_mat4 = np.exp(r_mat3) - 1.0                                                     # :924
    r_mat4[r_mat4 < -0.0] = 0.0                                                       # :925
    r_mat4 /= r_mat4.sum(axis=1, keepdims=True)                                       # :926
    return r_mat4 * TGC.mean()                                                        # :927


# --------------------------------------------------------------------------------------
# QC (parity row P0; host work in the reference too)
# --------------------------------------------------------------------------------------
def preprocess(X, gene_name, min_tp_c=0, min_tp_g=0, max_tp_c=np.inf, max_tp_g=np.inf, min_genes_per_cell=200,
               max_genes_per_cell=0, min_cells_per_gene=15, mito_percent=5.0, ribo_percent=0.0):
    """preprocess (:160-236) restated literally on a dense Float32 matrix.  Returns
    (fc_idx, gene_idx) where gene_idx = positions of the surviving genes in output order
    (fg_idx -> nn_idx -> stable sortperm of the Float32 per-gene mean, :218-225), or None."""
    import re
    A = np.asarray(X.todense() if sp.issparse(X) else X, dtype=np.float32)
    n_cell_counts = (A != 0).sum(axis=0)                                             # :184
    n_cell_counts_sum = A.sum(axis=0, dtype=np.float32)                              # :185
    fg_idx = (n_cell_counts_sum > min_tp_g) & (n_cell_counts_sum < max_tp_g) & (n_cell_counts >= min_cells_per_gene)
    n_gene_counts = (A != 0).sum(axis=1)                                             # :191
    n_gene_counts_sum = A.sum(axis=1, dtype=np.float32)                              # :192
    b1 = n_gene_counts_sum > min_tp_c
    b2 = n_gene_counts_sum < max_tp_c
    b3 = n_gene_counts >= min_genes_per_cell
    mito = np.array([re.match(r"^mt-.", g, flags=re.I) is not None for g in gene_name])   # :196
    ribo = np.array([re.match(r"^RP[SL].", g, flags=re.I) is not None for g in gene_name])
    with np.errstate(invalid="ignore", divide="ignore"):
        if mito_percent == 0:
            b4 = np.ones_like(b1)
        else:
            ratio = A[:, mito].sum(axis=1, dtype=np.float32) / n_gene_counts_sum      # Float32 ratio
            b4 = ratio.astype(np.float64) < mito_percent / 100                        # :201 strict
        if ribo_percent == 0:
            b5 = np.ones_like(b1)
        else:
            ratio = A[:, ribo].sum(axis=1, dtype=np.float32) / n_gene_counts_sum
            b5 = ratio.astype(np.float64) < ribo_percent / 100
    b6 = np.ones_like(b1) if max_genes_per_cell == 0 else n_gene_counts < max_genes_per_cell
    fc_idx = b1 & b2 & b3 & b4 & b5 & b6                                             # :216
    if not (fc_idx.any() and fg_idx.any()):
        return None
    oo = A[fc_idx][:, fg_idx]
    nn_idx = oo.sum(axis=0, dtype=np.float32) != 0                                   # :220
    oo = oo[:, nn_idx]
    mean_ = (oo.sum(axis=0, dtype=np.float32) / np.float32(oo.shape[0])).astype(np.float32)
    s_idx = np.argsort(mean_, kind="stable")                                         # :224
    gene_idx = np.nonzero(fg_idx)[0][nn_idx][s_idx]
    return fc_idx, gene_idx
