"""CPU tests of the oracle (test infrastructure): soft known answers, golden fixtures, and the
algebraic identities each restated function must satisfy.  The reference has no tests or golden
vectors for this path (parity unpinned, see oracle header); the pins here are
  * tests/golden/z8eq_pca_summary.json - structure of the reference's own committed output out/pca.csv
  * tests/golden/z785_*.npz            - a real bundled dataset after QC + oracle outputs on it
  * RNG-independent known answers recorded in SURVEY.md 8c for z_data_785 (shape, nnz, 11 signals, lambda_c)
"""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import sclens_oracle as orc
from sclens_b200.synth import make_counts

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def small_run():
    X = make_counts(360, 520, seed=2, K=4, de_prob=0.3, lfc_sd=1.5)
    res, draws, info = orc.sclens(X, rng=np.random.default_rng(1), mode="cpu", n_perturb=4, n_baseline=100)
    return X, res, draws, info


def test_reference_output_structure_is_reproduced(small_run):
    """out/pca.csv (the reference's only committed result): columns of :pca_n1 are orthogonal, have zero
    mean, and their sums of squares are the robust signal eigenvalues (all above lambda_c).  The oracle's
    output has exactly that structure."""
    gold = json.load(open(os.path.join(GOLD, "z8eq_pca_summary.json")))
    assert gold["n_cells"] == 3960 and gold["n_robust"] == 9
    assert gold["max_abs_offdiag_cos"] < 1e-5 and gold["max_abs_col_mean"] < 1e-6
    assert np.all(np.diff(gold["col_sumsq"]) < 0)
    X, res, _, _ = small_run
    P = np.asarray(res["pca"], np.float64)
    G = P.T @ P
    np.testing.assert_allclose(np.diag(G), res["signal_ev"], rtol=1e-9)
    d = np.sqrt(np.diag(G))
    assert np.max(np.abs(G / d[:, None] / d[None, :] - np.eye(len(d)))) < 1e-9
    assert np.max(np.abs(P.mean(axis=0))) < 1e-9
    assert np.all(np.asarray(res["signal_ev"]) > res["lambda"])
    assert np.all(np.diff(res["signal_ev"]) < 0)


def test_z785_known_answers_and_golden():
    z = np.load(os.path.join(GOLD, "z785_qc.npz"))
    X = sp.csc_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(z["shape"]))
    # RNG-independent known answers: post-QC shape and nnz
    assert X.shape == (777, 4782) and X.nnz == 392801
    g = np.load(os.path.join(GOLD, "z785_oracle.npz"))
    rng = np.random.default_rng(int(g["seed"]))
    draws = orc.Draws()
    draws.z_idx1, draws.z_idx2 = orc.draw_zero_candidates(X, rng)
    scaled, _ = orc.normalize_main(X)
    draws.null_perm, draws.null_rows = orc.draw_null(X, rng)
    Xr = orc.build_null(X, draws.null_perm, draws.null_rows)
    np.testing.assert_array_equal(Xr.indptr, g["null_indptr"])
    assert int(np.sum(Xr.indices.astype(np.int64) * 31 + Xr.data.astype(np.int64))) == int(g["null_checksum"])
    nL, nV, L, L_mp, lam, b_min, b_plus, _ = orc.get_sigev(scaled, orc.logn_scale_pre_scale(Xr), "gpu-ref")
    assert len(nL) == 11                                  # SURVEY.md 8c known answer
    assert abs(float(lam) - 1.80686) < 2e-3               # SURVEY.md 8c known answer (RNG-dependent to ~1e-3)
    assert abs(float(lam) - float(g["lambda_c"])) < 1e-5 * float(lam)
    np.testing.assert_allclose(nL, g["signal_ev"], rtol=2e-5)
    np.testing.assert_allclose(L[-100:], g["L"][-100:], rtol=2e-5)
    assert len(L_mp) == int(g["n_Lmp"])


def test_normalisation_identities_and_sparse_derivation():
    X = make_counts(300, 420, seed=5, K=3)
    N, M = X.shape
    out, rec = orc.normalize_main(X)
    assert np.max(np.abs(out.mean(axis=0))) < 1e-12                      # final re-centring (:696)
    np.testing.assert_allclose(out, orc.logn_scale_pre_scale(X), atol=2e-5)   # generic f32/f64 path == inline f64 path
    # everything from the sparse matrix only (what the CUDA kernels implement)
    r = rec["TGC"]
    C = X.tocoo()
    y = np.log1p(C.data.astype(np.float64) * (1.0 / r[C.row]))
    ybar = np.bincount(C.col, weights=y, minlength=M) / N
    cnt = np.bincount(C.col, minlength=M)
    s2 = np.bincount(C.col, weights=(y - ybar[C.col]) ** 2, minlength=M) + (N - cnt) * ybar ** 2
    sigma = np.sqrt(s2 / (N - 1))
    np.testing.assert_allclose(ybar, rec["mat2_mean"].ravel(), rtol=1e-12)
    np.testing.assert_allclose(sigma, rec["mat2_std"].ravel(), rtol=1e-11)
    z = y / sigma[C.col]
    mu = ybar / sigma
    l2 = np.sqrt(np.bincount(C.row, weights=z * z, minlength=N) - 2 * np.bincount(C.row, weights=z * mu[C.col], minlength=N) + mu @ mu)
    np.testing.assert_allclose(l2, rec["norm_tgc"], rtol=1e-11)
    inv_s = l2.mean() / l2
    cent = (np.bincount(C.col, weights=z * inv_s[C.row], minlength=M) - mu * inv_s.sum()) / N
    np.testing.assert_allclose(cent, rec["cent_"].ravel(), rtol=1e-8, atol=1e-15)
    dense = -mu[None, :] * inv_s[:, None] - cent[None, :]
    dense[C.row, C.col] = (z - mu[C.col]) * inv_s[C.row] - cent[C.col]
    np.testing.assert_allclose(dense, out, atol=1e-12)


def test_wishart_divides_by_size2_in_both_branches():
    rng = np.random.default_rng(0)
    A = rng.standard_normal((30, 50))
    np.testing.assert_allclose(orc.wishart_matrix(A, 1, "cpu"), A @ A.T / 50)
    np.testing.assert_allclose(orc.wishart_matrix(A, 2, "cpu"), A.T @ A / 50)       # Appendix A9: still / size(X,2)
    assert orc.wishart_matrix(A, 1, "gpu-ref").dtype == np.float32


def test_mp_fit_on_exact_wishart_noise():
    rng = np.random.default_rng(3)
    n, K = 400, 1000
    L = np.linalg.eigvalsh(np.cov(rng.standard_normal((n, K)), bias=True))
    Lr = np.linalg.eigvalsh(np.cov(rng.standard_normal((n, K)), bias=True))
    L_mp, b_plus, b_min, it = orc.mp_calculation(L, Lr[:-1])
    g = n / K
    assert abs(b_plus - (1 + np.sqrt(g)) ** 2) < 0.08 and abs(b_min - (1 - np.sqrt(g)) ** 2) < 0.05
    lam, gamma, p, sigma = orc.tw(L, L_mp)
    assert lam > L.max() * 0.97 and (L > lam).sum() <= 1          # pure noise: (almost) no signal
    Ls = L.copy()
    Ls[-2:] = [4.0, 6.0]
    L_mp2, _, _, _ = orc.mp_calculation(Ls, Lr[:-1])
    assert (Ls > orc.tw(Ls, L_mp2)[0]).sum() == 2
    chk = orc.mp_check(L_mp)
    assert chk["pass"] and 0 <= chk["ks_static"] < 0.2
    p_ = orc.mp_parameters(L.astype(np.float32))
    assert p_["gamma"].dtype == np.float32                        # dtype of L is preserved (Appendix A7/A8)


def test_draw_semantics():
    X = make_counts(200, 300, seed=9, K=3)
    N, M = X.shape
    rng = np.random.default_rng(0)
    z1, z2 = orc.draw_zero_candidates(X, rng)
    key = z2.astype(np.int64) * N + z1
    assert len(np.unique(key)) == len(key)                        # setdiff: unique (Appendix A4)
    assert not np.any(np.asarray(X[z1, z2]).ravel())              # ... and none of them is a non-zero
    assert 0.8 * X.nnz < len(z1) <= X.nnz
    perm, rows = orc.draw_null(X, rng)
    R = orc.build_null(X, perm, rows)
    np.testing.assert_array_equal(np.diff(R.indptr), np.diff(X.indptr))   # aligned order: counts per gene preserved
    np.testing.assert_array_equal(np.sort(R.data), np.sort(X.data))       # shuffle: same multiset of values
    perm2, rows2 = orc.draw_null(X, rng, gene_order=rng.permutation(M))   # Dict-order emulation: duplicates summed
    R2 = orc.build_null(X, perm2, rows2)
    assert R2.nnz < X.nnz and abs(R2.data.sum() - X.data.sum()) < 1e-3
    assert orc.julia_round(2.5) == 2 and orc.julia_round(3.5) == 4 and orc.julia_round(0.4999) == 0   # Appendix A12
    P = orc.perturbed_matrix(X, z1, z2, np.arange(50), binarise=True)
    assert P.nnz == X.nnz + 50 and set(np.unique(P.data)) == {1.0}


def test_robustness_scores_limits():
    rng = np.random.default_rng(1)
    N, k = 300, 3
    nV = np.linalg.qr(rng.standard_normal((N, k)))[0]
    same = [np.concatenate([nV[:, ::-1], np.linalg.qr(rng.standard_normal((N, 2)))[0]], axis=1) for _ in range(5)]
    rob, sig = orc.robustness_scores(nV, same, 60)
    np.testing.assert_allclose(rob["m_scores"], 1.0, atol=1e-12)
    np.testing.assert_array_equal(sig, [0, 1, 2])
    noise = [np.linalg.qr(rng.standard_normal((N, 5)))[0] for _ in range(5)]
    rob2, sig2 = orc.robustness_scores(nV, noise, 60)
    assert rob2["m_scores"].max() < 0.5 and len(sig2) == 0
    assert rob["b_"].shape == (k, 10)


def test_search_and_outputs_consistency(small_run):
    X, res, draws, info = small_run
    N, M = X.shape
    assert 0.9 <= info["p_sel"] <= 0.999 and info["n_search"] >= 5        # stop rule needs 5 consecutive (A16)
    assert info["n_add"] == orc.julia_round((1 - info["p_sel"]) * M * N)
    assert info["min_pc"] == int(np.ceil(1.5 * len(res["signal_ev"])))
    assert all(len(s) == orc.julia_round((1 - t[0]) * M * N) for s, t in zip(draws.search_sple, info["search_trace"]))
    k = len(res["signal_ev"])
    assert res["gene_basis"].shape == (k, M) and res["pca"].shape == (N, k)
    # replaying the recorded draws reproduces the run exactly
    res2, _, info2 = orc.sclens(X, draws=draws, mode="cpu", n_perturb=4)
    assert info2["p_sel"] == info["p_sel"] and np.array_equal(res2["sig_id"], res["sig_id"])
    np.testing.assert_array_equal(res2["L"], res["L"])


def test_preprocess_matches_host_mirror():
    from sclens_b200.preprocess import qc_indices
    rng = np.random.default_rng(4)
    X = make_counts(400, 600, seed=6, K=3, sparsity=0.5).tolil()
    genes = np.array([f"g{j}" for j in range(600)], dtype=object)
    genes[:5] = ["MT-CO1", "mt-Nd1", "Mt-x", "MTOR", "RPS3"]
    X[:20, :5] = 400.0            # cells dominated by mitochondrial counts are dropped (strict <, Float32 ratio)
    X[30:45, :] = 0               # empty cells
    X[:, 100:110] = 0             # empty genes
    X[50:60, 120] = 3.0
    X = sp.csc_matrix(X, dtype=np.float32)
    a = orc.preprocess(X, genes)
    b = qc_indices(X, genes)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])
    assert not a[0][:20].any() and not a[0][30:45].any() and a[0].sum() > 300
    assert not np.isin(np.arange(100, 110), a[1]).any()


def test_denoise_inverts_the_normalisation():
    """get_denoised_df (:889-931) with the complete eigenbasis as "signals": pca * gene_basis * sqrt(M) is then the
    normalised matrix itself, so the recorded normalisation is undone exactly and the output must be
    x_ij / r_i * mean(TGC) (the clamp at 0 and the row rescaling are identities on it)."""
    X = make_counts(120, 90, seed=5, K=3, de_prob=0.3, lfc_sd=1.5)
    scaled, rec = orc.normalize_main(X)
    N, M = X.shape
    U, sv, Vt = np.linalg.svd(scaled, full_matrices=False)
    keep = sv > 1e-9 * sv[0]
    nV, nL = U[:, keep], (sv[keep] ** 2) / M                   # eigenpairs of scaled * scaled' / M  (:339-343)
    sq = np.sqrt(nL)
    result = {"pca_n1": nV * sq[None, :], "sig_id": np.arange(keep.sum()),
              "gene_basis": (1.0 / sq)[:, None] * (nV.T @ scaled) / np.sqrt(M), "rec_vals": rec}     # :810-819
    out = orc.get_denoised(result, mode="cpu")
    dense = np.asarray(X.todense(), dtype=np.float64)
    want = dense / dense.sum(axis=1, keepdims=True) * rec["TGC"].mean()
    assert out.shape == (N, M)
    np.testing.assert_allclose(out.sum(axis=1), rec["TGC"].mean(), rtol=1e-12)
    # Float32 reconstruction (:891-896) of values whose exp(.) - 1 is ~1e-3: a few 1e-4 relative on the non-zeros
    nz = dense > 0
    np.testing.assert_allclose(out[nz], want[nz], rtol=5e-3)
    assert np.abs(out[~nz]).max() < 1e-3 * want[nz].min()


def test_denoise_on_a_full_run_is_a_probability_profile_per_cell(small_run):
    X, res, _, _ = small_run
    assert len(res["sig_id"]) > 0
    out = orc.get_denoised(res, mode="gpu-ref")
    assert out.shape == X.shape and np.isfinite(out).all() and (out >= 0).all()
    np.testing.assert_allclose(out.sum(axis=1), res["rec_vals"]["TGC"].mean(), rtol=1e-10)
    # the two reference arithmetic paths (:893-896 Float32 GEMM, :903-905 Float64 operands) agree to Float32 accuracy
    np.testing.assert_allclose(out, orc.get_denoised(res, mode="cpu"), rtol=1e-3, atol=1e-6 * out.max())


def test_median_centering_restatement_properties():
    """centering="median" (:653-654, :294-299, :608): undoing the row scaling leaves every gene with median 0 and corrected
    std 1 over all cells; every row has the same l2 norm; genes expressed in at most half the cells keep their zeros at
    the common background -0/std = 0 before scaling (their median is an implicit zero)."""
    from sclens_b200.synth import make_counts
    X = make_counts(301, 420, seed=5, K=3, de_prob=0.3, lfc_sd=1.5).tolil()
    rng = np.random.default_rng(0)
    for j in range(6):                                   # a few genes expressed in most cells: non-zero medians
        X[:, j] = rng.poisson(4.0, size=(301, 1)).astype(np.float32)
    X = sp.csc_matrix(X, dtype=np.float32)
    X.eliminate_zeros()
    out = orc.logn_scale_pre_scale_median(X).astype(np.float64)
    l2 = np.sqrt((out ** 2).sum(axis=1))
    np.testing.assert_allclose(l2, l2.mean(), rtol=1e-5)
    # rows were scaled by mean(l2_raw)/l2_raw; the raw matrix is out * l2_raw / mean(l2_raw): recover it up to that factor
    rs = np.asarray(X.sum(axis=1)).ravel()
    Y = np.zeros(X.shape)
    C = X.tocoo()
    Y[C.row, C.col] = np.log1p(C.data / rs[C.row])
    W = (Y - np.median(Y, axis=0)) / Y.std(axis=0, ddof=1)
    want = W / np.sqrt((W ** 2).sum(axis=1))[:, None] * np.sqrt((W ** 2).sum(axis=1)).mean()
    np.testing.assert_allclose(out, want, rtol=2e-4, atol=2e-5)
    assert (np.median(Y, axis=0)[:6] > 0).all() and (np.median(Y, axis=0)[6:] == 0).sum() > 300
    res, draws, info = orc.sclens(X, rng=np.random.default_rng(3), mode="cpu", n_perturb=3, n_baseline=100, centering="median")
    assert res["rec_vals"] == {} and len(res["signal_ev"]) >= 1 and 0.9 <= info["p_sel"] <= 0.999
