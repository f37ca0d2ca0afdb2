"""CPU restatement (numpy, Float64 or Float32) of the two-stage symmetric eigensolver of sclens_b200/csrc/twostage*.cu.

TEST INFRASTRUCTURE ONLY - imported by tests/ (and nothing else).  It restates, step for step, the ALGEBRA the CUDA kernels
implement, so the scheme can be checked on a CPU (tests/test_two_stage_cpu.py) before and independently of the kernels:

  stage 1  dense -> band (half bandwidth b): per panel CholeskyQR with a Float64 Gram matrix, Householder reconstruction
           (Ballard et al., "Reconstructing Householder vectors from TSQR": LU of Q1 - S), T from V'V, two-sided update
           A22 -= V W' + W V'
  stage 2  band -> tridiagonal by bulge chasing (Householder, column by column), reflectors kept
  back     Z = Q1 Q2 E: Q2 applied in blocks of g sweeps x one chase level (the order proven below), Q1 panel by panel or p
           panels at a time as one block reflector (apply_q1_blocked: the recursion of backtrans.cu::apply_q1_umma)

The reference's solver is `eigen(Symmetric(...))` / CUSOLVER syevd (src/scLENS.jl:375-387); this file has no counterpart
there - it documents the replacement's arithmetic.
"""
import numpy as np


def house(x):
    """LAPACK larfg: (v with v[0] = 1, tau, beta) with (I - tau v v') x = beta e1."""
    x = np.asarray(x)
    alpha = x[0]
    xn = np.linalg.norm(x[1:]) if x.size > 1 else 0.0
    v = np.zeros_like(x)
    v[0] = 1
    if xn == 0:
        return v, x.dtype.type(0), alpha
    beta = -np.copysign(np.hypot(alpha, xn), alpha)
    tau = (beta - alpha) / beta
    v[1:] = x[1:] / (alpha - beta)
    return v, x.dtype.type(tau), x.dtype.type(beta)


def t_from_v(V):
    """T of I - V T V' = H_0 H_1 ... (forward, columnwise) for unit lower trapezoidal V: inv(T) = striu(V'V) + diag(V'V)/2."""
    G = V.astype(np.float64).T @ V.astype(np.float64)
    Ti = np.triu(G, 1) + np.diag(0.5 * np.diag(G))
    return np.linalg.inv(Ti)


def panel_cholqr_hr(P):
    """m x b panel (m >= b, full column rank) -> (V unit lower trapezoidal m x b, T b x b, R' b x b upper) with
    (I - V T V')' P = [R'; 0] up to the panel's conditioning times the working precision."""
    m, b = P.shape
    P64 = P.astype(np.float64)
    G = P64.T @ P64
    R = np.linalg.cholesky(G).T                      # P = Q R
    Q1 = np.linalg.solve(R.T, P64[:b].T).T           # top b x b block of Q = P R^-1
    # LU with signs of Q1 - S: L unit lower, U upper, S_jj = -sign(pivot before the shift)
    X = Q1.copy()
    S = np.zeros(b)
    L = np.eye(b)
    U = np.zeros((b, b))
    for j in range(b):
        S[j] = -1.0 if X[j, j] >= 0 else 1.0
        X[j, j] -= S[j]
        U[j, j:] = X[j, j:]
        L[j + 1:, j] = X[j + 1:, j] / X[j, j]
        X[j + 1:, j + 1:] -= np.outer(L[j + 1:, j], U[j, j + 1:])
    # rows below the top block: V2 = Q2 U^-1 = P2 (U R)^-1
    Mi = np.linalg.inv(U @ R)
    V = np.empty((m, b))
    V[:b] = L
    V[b:] = P64[b:] @ Mi
    V = V.astype(P.dtype)
    T = t_from_v(V)
    Rp = S[:, None] * R
    return V, T.astype(P.dtype), Rp.astype(P.dtype)


def sy2sb(A, b):
    """A symmetric n x n -> (band matrix as a dense n x n symmetric array with half bandwidth b, list of (row0, V, T))."""
    A = A.copy()
    n = A.shape[0]
    refl = []
    k = 0
    while True:
        c0 = k * b
        r0 = c0 + b
        m = n - r0
        if m <= 1:
            break
        P = A[r0:, c0:r0]
        if m >= b:
            V, T, Rp = panel_cholqr_hr(P)
            Rfull = np.zeros_like(P)
            Rfull[:b] = Rp
        else:
            # short tail: plain Householder QR of the m x b block
            Pw = P.copy()
            V = np.zeros((m, m - 1), dtype=A.dtype)
            for j in range(m - 1):
                v, tau, beta = house(Pw[j:, j])
                Pw[j:, j:] -= tau * np.outer(v, v @ Pw[j:, j:])
                V[j:, j] = v
            T = t_from_v(V).astype(A.dtype)
            Rfull = np.triu(Pw)
        A[r0:, c0:r0] = Rfull
        A[c0:r0, r0:] = Rfull.T
        A22 = A[r0:, r0:]
        Z = (A22 @ V) @ T
        S = T.T @ (V.T @ Z)
        W = Z - 0.5 * V @ S
        A22 -= V @ W.T + W @ V.T
        refl.append((r0, V, T))
        k += 1
    return A, refl


def sb2st(B, b):
    """Bulge chasing on a dense copy of the band matrix.  Returns (d, e, reflectors) with reflectors[(s, k)] = (row0, v, tau)."""
    B = B.copy()
    n = B.shape[0]
    refl = {}
    for s in range(n - 2):
        # step 0: annihilate column s below the subdiagonal
        r0 = s + 1
        L = min(b, n - r0)
        if L < 2:
            break
        v, tau, beta = house(B[r0:r0 + L, s])
        B[r0:r0 + L, s] = 0
        B[r0, s] = beta
        B[s, r0:r0 + L] = B[r0:r0 + L, s]
        refl[(s, 0)] = (r0, v, tau)
        D = B[r0:r0 + L, r0:r0 + L]
        w = tau * (D @ v)
        w -= 0.5 * tau * (w @ v) * v
        D -= np.outer(v, w) + np.outer(w, v)
        k = 1
        while True:
            r1 = r0 + L            # first row of the block below
            L1 = min(b, n - r1)
            if L1 < 1:
                break
            Bk = B[r1:r1 + L1, r0:r0 + L]
            Bk -= tau * np.outer(Bk @ v, v)          # right apply the previous reflector: creates the bulge
            if L1 >= 2:
                v1, tau1, beta1 = house(Bk[:, 0])
                Bk[:, 0] = 0
                Bk[0, 0] = beta1
                Bk[:, 1:] -= tau1 * np.outer(v1, v1 @ Bk[:, 1:])
            B[r0:r0 + L, r1:r1 + L1] = Bk.T
            if L1 < 2:
                break
            refl[(s, k)] = (r1, v1, tau1)
            D = B[r1:r1 + L1, r1:r1 + L1]
            w = tau1 * (D @ v1)
            w -= 0.5 * tau1 * (w @ v1) * v1
            D -= np.outer(v1, w) + np.outer(w, v1)
            r0, L, v, tau = r1, L1, v1, tau1
            k += 1
    return np.diag(B).copy(), np.diag(B, -1).copy(), refl


def apply_q2(refl, n, b, g, Z):
    """Z <- Q2 Z with Q2 = product of the bulge-chasing reflectors in generation order (sweep-major), applied in blocks of g
    consecutive sweeps at one chase level: groups descending, levels ascending inside a group, each block as I - V T V'."""
    Z = Z.copy()
    n_sweeps = max((s for s, _ in refl), default=-1) + 1
    n_groups = (n_sweeps + g - 1) // g
    for G in range(n_groups - 1, -1, -1):
        s0, s1 = G * g, min(n_sweeps, G * g + g)
        k = 0
        while True:
            members = [(s, refl[(s, k)]) for s in range(s0, s1) if (s, k) in refl]
            if not members:
                break
            rlo = min(r for _, (r, v, t) in members)
            rhi = max(r + v.size for _, (r, v, t) in members)
            V = np.zeros((rhi - rlo, len(members)), dtype=Z.dtype)
            for c, (s, (r, v, t)) in enumerate(members):
                V[r - rlo:r - rlo + v.size, c] = v
            # T from the taus (forward larft)
            T = np.zeros((len(members), len(members)))
            for c, (s, (r, v, t)) in enumerate(members):
                T[c, c] = t
                if c:
                    T[:c, c] = -t * (T[:c, :c] @ (V[:, :c].T.astype(np.float64) @ V[:, c].astype(np.float64)))
            T = T.astype(Z.dtype)
            Z[rlo:rhi] -= V @ (T @ (V.T @ Z[rlo:rhi]))
            k += 1
    return Z


def apply_q2_plain(refl, Z):
    """generation-order product applied to Z one reflector at a time (the definition apply_q2 must agree with)."""
    Z = Z.copy()
    for key in sorted(refl.keys(), reverse=True):
        r, v, t = refl[key]
        Z[r:r + v.size] -= t * np.outer(v, v @ Z[r:r + v.size])
    return Z


def apply_q1(refl1, Z):
    Z = Z.copy()
    for r0, V, T in reversed(refl1):
        Z[r0:] -= V @ (T @ (V.T @ Z[r0:]))
    return Z


def apply_q1_blocked(refl1, Z, p=8):
    """Z <- Q1 Z with p consecutive panels applied as ONE block reflector, as backtrans.cu::apply_q1_umma does on the tcgen05 GEMM:
    H_a H_a+1 ... H_a+p-1 Z = Z - sum_j (V_j T_j) X_j with X_j = V_j' Z - sum_{l > j} (V_j' V_l T_l) X_l, j descending.  The large
    products are Y = V_blk' Z and Z -= (VT)_blk X; the couplings S_jl = (V_blk' V_blk)_jl T_l come from the Gram matrix of the block.
    Panels are padded with zeros above their first row so the block shares the row range of its first panel."""
    Z = Z.copy()
    n = Z.shape[0]
    for a in range((len(refl1) - 1) // p * p, -1, -p):
        blk = refl1[a:a + p]
        r0 = blk[0][0]
        w = blk[0][1].shape[1]
        Vb = np.zeros((n - r0, w * len(blk)), dtype=Z.dtype)
        VTb = np.zeros_like(Vb)
        for j, (rj, V, T) in enumerate(blk):
            Vb[rj - r0:, w * j:w * j + V.shape[1]] = V
            VTb[rj - r0:, w * j:w * j + V.shape[1]] = V @ T
        Y = Vb.T @ Z[r0:]
        G = Vb.T @ Vb
        X = Y.copy()
        for j in range(len(blk) - 2, -1, -1):
            for l in range(j + 1, len(blk)):
                Tl = np.zeros((w, w), dtype=Z.dtype)
                Tl[:blk[l][2].shape[0], :blk[l][2].shape[1]] = blk[l][2]
                S = G[w * j:w * j + w, w * l:w * l + w] @ Tl
                X[w * j:w * j + w] -= S @ X[w * l:w * l + w]
        Z[r0:] -= VTb @ X
    return Z


def eigh_two_stage(A, b=8, g=4):
    Bd, refl1 = sy2sb(A, b)
    d, e, refl2 = sb2st(Bd, b)
    Tm = np.diag(d.astype(np.float64)) + np.diag(e.astype(np.float64), 1) + np.diag(e.astype(np.float64), -1)
    w, E = np.linalg.eigh(Tm)
    Z = apply_q2(refl2, A.shape[0], b, g, E.astype(A.dtype))
    Z = apply_q1(refl1, Z)
    return w.astype(A.dtype), Z
