"""-m gpu: scLENS.preprocess (:160-236) on the device against the oracle's literal restatement - the QC masks, the drop of
genes left empty and the stable sort by Float32 mean must agree index for index."""
import ctypes as C

import numpy as np
import pandas as pd
import pytest
import scipy.sparse as sp

from oracle import sclens_oracle as orc
from sclens_b200 import Handle, SclError, _lib
from sclens_b200._lib import ptr
from sclens_b200.api import preprocess, qc_indices_device
from sclens_b200.synth import make_counts

pytestmark = pytest.mark.gpu


def raw_counts(N, M, seed):
    X = make_counts(N, M, seed=seed, K=3, sparsity=0.5).tolil()
    genes = np.array([f"g{j}" for j in range(M)], dtype=object)
    genes[:6] = ["MT-CO1", "mt-Nd1", "Mt-x", "MTOR", "RPS3", "rpl11"]
    X[:20, :5] = 400.0            # cells dominated by mitochondrial counts are dropped (strict <, Float32 ratio)
    X[30:45, :] = 0               # empty cells
    X[:, 100:110] = 0             # empty genes
    X[50:60, 120] = 3.0           # a gene expressed in fewer than min_cells_per_gene cells
    X[60:75, 130] = 2.0           # exactly 15 cells ... of which some are dropped below: empty-gene drop after filtering
    X[60:75, 5:400] = 0
    return sp.csc_matrix(X, dtype=np.float32), genes


@pytest.mark.parametrize("kw", [{}, {"mito_percent": 0.0}, {"ribo_percent": 30.0, "max_genes_per_cell": 330},
                                {"min_tp_c": 500, "max_tp_g": 5000, "min_cells_per_gene": 40}])
def test_device_qc_matches_oracle_index_for_index(kw):
    X, genes = raw_counts(400, 600, seed=6)
    want = orc.preprocess(X, genes, **kw)
    with Handle() as h:
        got = qc_indices_device(X, genes, handle=h, **kw)
        assert (want is None) == (got is None)
        fc_idx, gene_idx, out = got
        np.testing.assert_array_equal(fc_idx, np.nonzero(want[0])[0])          # kept cells
        np.testing.assert_array_equal(gene_idx, want[1])                       # kept genes, in output order
        ref = X[np.nonzero(want[0])[0]][:, want[1]].tocsc()
        ref.sort_indices()
        np.testing.assert_array_equal(out.indptr, ref.indptr)                  # the filtered matrix, bit for bit
        np.testing.assert_array_equal(out.indices, ref.indices)
        np.testing.assert_array_equal(out.data, ref.data)
        means = np.asarray(out.sum(axis=0)).ravel() / out.shape[0]
        assert np.all(np.diff(means.astype(np.float32)) >= 0)                  # sorted by mean expression (:224)
        # the filtered matrix is the handle's counts now: the signal stage runs on it without another upload
        si = h.run_signal()
        assert (si.N, si.M) == out.shape


def test_device_preprocess_dataframe_and_empty_result():
    X, genes = raw_counts(300, 500, seed=7)
    df = pd.DataFrame.sparse.from_spmatrix(X, columns=genes)
    df.insert(0, "cell", [f"c{i}" for i in range(X.shape[0])])
    from sclens_b200.preprocess import preprocess as host_preprocess
    a = preprocess(df, verbose=False)
    b = host_preprocess(df, verbose=False)
    assert list(a.columns) == list(b.columns) and list(a["cell"]) == list(b["cell"])
    np.testing.assert_array_equal(a.iloc[:, 1:].sparse.to_coo().toarray(), b.iloc[:, 1:].sparse.to_coo().toarray())
    assert preprocess(df, min_genes_per_cell=10 ** 6, verbose=False) is None   # "There is no high quality cells and genes"


def test_uploads_are_validated():
    """Host-supplied structure is checked once on upload (monotone column pointers, rows in range and strictly increasing,
    positive values; candidates inside the grid) and refused with SCL_ERR_INVALID instead of being indexed with."""
    X = make_counts(300, 400, seed=3)
    with Handle() as h:
        h.set_counts(X)

        def upload(indptr, indices, data):
            return h.lib.scl_set_counts_csc(h.h, 300, 400, len(data), ptr(_lib.as_u32(indptr), C.c_uint32),
                                            ptr(_lib.as_u32(indices), C.c_uint32), ptr(_lib.as_f32(data), C.c_float), 0)

        bad = X.indices.copy(); bad[5] = 300
        assert upload(X.indptr, bad, X.data) == -1 and b"outside" in h.lib.scl_last_error(h.h)
        bad = X.indices.copy(); bad[[0, 1]] = bad[[1, 0]]
        assert upload(X.indptr, bad, X.data) == -1 and b"increasing" in h.lib.scl_last_error(h.h)
        bad = X.indptr.copy(); bad[3], bad[4] = bad[4], bad[3]
        assert upload(bad, X.indices, X.data) == -1
        bad = X.data.copy(); bad[7] = 0.0
        assert upload(X.indptr, X.indices, bad) == -1 and b"positive" in h.lib.scl_last_error(h.h)
        assert upload(X.indptr, X.indices, X.data) == 0
        z1, z2 = np.array([1, 300], np.uint32), np.array([2, 3], np.uint32)
        with pytest.raises(SclError):
            h.set_zero_candidates(z1, z2)
