"""-m gpu parity of get_denoised_df (src/scLENS.jl:889-931): the fused device kernel behind scl_op_denoise against the
oracle's restatement, on result dictionaries built from the oracle's normalisation (so the test needs no eigensolver)."""
import numpy as np
import pytest

from oracle import sclens_oracle as orc
from sclens_b200 import get_denoised_df
from sclens_b200.synth import make_counts

pytestmark = pytest.mark.gpu


def make_result(N, M, r, seed):
    X = make_counts(N, M, seed=seed, K=4, de_prob=0.3, lfc_sd=1.5)
    scaled, rec = orc.normalize_main(X)
    U, sv, _ = np.linalg.svd(scaled, full_matrices=False)
    k = r + 3                                           # k "signals", r of them "robust"
    nV = U[:, :k].astype(np.float32)
    nL = ((sv[:k] ** 2) / M).astype(np.float32)
    sq = np.sqrt(nL)
    sig_id = np.sort(np.random.default_rng(seed).choice(k, size=r, replace=False))
    gene_basis = ((1.0 / sq)[:, None] * (nV.T.astype(np.float64) @ scaled) / np.sqrt(M)).astype(np.float32)   # :813-819
    return {"pca_n1": (nV[:, sig_id] * sq[sig_id][None, :]).astype(np.float32), "sig_id": sig_id, "gene_basis": gene_basis,
            "rec_vals": rec, "cell_id": np.array([f"c{i}" for i in range(N)]), "gene_id": np.array([f"g{j}" for j in range(M)])}


@pytest.mark.parametrize("N,M,r", [(301, 423, 5), (64, 1000, 1), (1500, 37, 9)])   # ragged sizes: N % 32 != 0, M % 8 != 0
def test_denoise_matches_oracle(N, M, r):
    res = make_result(N, M, r, seed=N + r)
    want = orc.get_denoised(res, mode="gpu-ref")
    odf = get_denoised_df(res)
    assert list(odf.columns[:3]) == ["cell", "g0", "g1"] and odf.shape == (N, M + 1)
    got = odf.iloc[:, 1:].to_numpy()
    assert got.dtype == np.float64
    # Float32 inner products accumulated in a different order than numpy's (:893-896), then exp(.) - 1 of values ~1e-3
    np.testing.assert_allclose(got, want, rtol=2e-4, atol=1e-6 * want.max())
    np.testing.assert_allclose(got.sum(axis=1), res["rec_vals"]["TGC"].mean(), rtol=1e-12)
    got32 = get_denoised_df(res, dtype=np.float32).iloc[:, 1:].to_numpy()
    assert got32.dtype == np.float32
    np.testing.assert_allclose(got32, got, rtol=1e-6, atol=1e-7 * want.max())
