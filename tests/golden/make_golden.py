"""Generates the committed golden fixtures.  Runs HERE only (needs /root/reference); the tests
read the fixtures, never /root/reference.

  z8eq_pca_summary.json  structural facts of the reference's own committed output out/pca.csv
                         (written by example.jl:36 from :pca_n1) - the only golden artefact the
                         reference holds for this path
  z785_qc.npz            post-QC counts of data/Real_Zheng_data/z_data_785.csv.gz (CSC) - a real
                         input small enough to commit
  z785_oracle.npz        oracle outputs on it with seeded draws (signal stage + robustness)
"""
import json
import os
import sys

import numpy as np
import pandas as pd
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import sclens_oracle as orc  # noqa: E402

REF = "/root/reference"


def pca_summary():
    df = pd.read_csv(f"{REF}/out/pca.csv")
    A = df.iloc[:, 1:].to_numpy(np.float64)
    G = A.T @ A
    d = np.sqrt(np.diag(G))
    C = G / d[:, None] / d[None, :]
    out = {"n_cells": int(A.shape[0]), "n_robust": int(A.shape[1]),
           "col_sumsq": np.diag(G).tolist(), "max_abs_offdiag_cos": float(np.max(np.abs(C - np.eye(len(d))))),
           "max_abs_col_mean": float(np.max(np.abs(A.mean(axis=0)))),
           "source": "out/pca.csv of Mathbiomed/scLENS (example.jl:36)"}
    json.dump(out, open(os.path.join(HERE, "z8eq_pca_summary.json"), "w"), indent=1)
    print(out)


def z785():
    df = pd.read_csv(f"{REF}/data/Real_Zheng_data/z_data_785.csv.gz")
    genes = np.asarray(df.columns[1:], dtype=object)
    X = sp.csc_matrix(df.iloc[:, 1:].to_numpy(np.float32))
    fc, gidx = orc.preprocess(X, genes)
    Xq = X[fc][:, gidx].tocsc()
    Xq.sort_indices()
    print("post-QC", Xq.shape, Xq.nnz)
    np.savez_compressed(os.path.join(HERE, "z785_qc.npz"), indptr=Xq.indptr.astype(np.uint32),
                        indices=Xq.indices.astype(np.uint32), data=Xq.data.astype(np.float32), shape=np.array(Xq.shape),
                        fc_idx=np.nonzero(fc)[0].astype(np.int32), gene_idx=gidx.astype(np.int32),
                        raw_shape=np.array(X.shape))
    res, draws, info = orc.sclens(Xq, rng=np.random.default_rng(785), mode="gpu-ref", n_perturb=6, n_baseline=500)
    print("n_signal", len(res["signal_ev"]), "lambda_c", res["lambda"], "p_sel", info["p_sel"], "n_search", info["n_search"],
          "sig", res["sig_id"], "m", res["robustness_scores"]["m_scores"])
    tr = np.array(info["search_trace"])
    np.savez_compressed(os.path.join(HERE, "z785_oracle.npz"), L=np.asarray(res["L"], np.float32),
                        n_Lmp=len(res["L_mp"]), lambda_c=float(res["lambda"]), signal_ev=np.asarray(res["signal_ev"], np.float32),
                        signal_evec=np.asarray(res["signal_evec"], np.float32), p_sel=info["p_sel"], n_search=info["n_search"],
                        trace=tr, m_scores=res["robustness_scores"]["m_scores"], sig_id=res["sig_id"], p_th=draws.p_th,
                        b_plus=float(info["b_plus"]), b_min=float(info["b_min"]), seed=785,
                        null_indptr=info["null"].indptr.astype(np.uint32), null_nnz=info["null"].nnz,
                        null_checksum=int(np.sum(info["null"].indices.astype(np.int64) * 31 + info["null"].data.astype(np.int64))))


if __name__ == "__main__":
    pca_summary()
    z785()
