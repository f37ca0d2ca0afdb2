"""CPU check of the algebra behind the two-stage eigensolver (oracle/two_stage_ref.py restates, step for step, what
sclens_b200/csrc/sy2sb.cu, sb2st.cu and backtrans.cu compute): CholeskyQR + Householder reconstruction panels, bulge
chasing, and the blocked order in which the stage-2 reflectors are applied to the eigenvectors."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import two_stage_ref as ts  # noqa: E402


@pytest.mark.parametrize("n,b,g", [(50, 8, 4), (67, 8, 8), (100, 16, 16), (33, 4, 3)])
def test_two_stage_restatement(n, b, g):
    rng = np.random.default_rng(n)
    X = rng.standard_normal((n, 3 * n))
    A = X @ X.T / (3 * n)
    w0 = np.linalg.eigvalsh(A)
    Bd, r1 = ts.sy2sb(A, b)
    i, j = np.indices(A.shape)
    assert np.abs(Bd[np.abs(i - j) > b]).max() == 0.0
    assert np.abs(np.linalg.eigvalsh(Bd) - w0).max() < 1e-12
    d, e, r2 = ts.sb2st(Bd, b)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    assert np.abs(np.linalg.eigvalsh(T) - w0).max() < 1e-12
    eye = np.eye(n)
    # blocks of g sweeps x one chase level, groups descending, levels ascending == the reflectors one by one
    assert np.abs(ts.apply_q2(r2, n, b, g, eye) - ts.apply_q2_plain(r2, eye)).max() < 1e-13
    w, Z = ts.eigh_two_stage(A, b, g)
    assert np.abs(A @ Z - Z * w).max() < 1e-12 and np.abs(Z.T @ Z - eye).max() < 1e-12


def test_panel_reconstruction_is_a_householder_qr():
    rng = np.random.default_rng(3)
    P = rng.standard_normal((90, 8)) * np.exp(rng.normal(0, 2, size=8))
    V, T, Rp = ts.panel_cholqr_hr(P)
    H = np.eye(90) - V @ T @ V.T
    assert np.abs(H.T @ H - np.eye(90)).max() < 1e-12
    R = H.T @ P
    assert np.abs(R[8:]).max() < 1e-10 * np.abs(P).max() and np.abs(R[:8] - Rp).max() < 1e-10 * np.abs(P).max()
    assert np.abs(np.tril(V[:8], -1) + np.eye(8) - V[:8]).max() == 0.0


@pytest.mark.parametrize("n,b,p", [(100, 8, 4), (131, 8, 8), (70, 4, 3)])
def test_q1_block_reflectors_equal_panel_by_panel(n, b, p):
    """The recursion behind the tcgen05 stage-1 back-transformation (p panels as one block reflector, couplings from the Gram matrix
    of the block) gives what the panels give one by one - including a last, short panel and a first block with fewer panels."""
    rng = np.random.default_rng(n + p)
    X = rng.standard_normal((n, 2 * n))
    A = X @ X.T / (2 * n)
    _, r1 = ts.sy2sb(A, b)
    Z = rng.standard_normal((n, 17))
    assert np.abs(ts.apply_q1_blocked(r1, Z, p) - ts.apply_q1(r1, Z)).max() < 1e-12
