"""Every own-kernel stage of the sclens() path once, at the shapes of one workload, WITHOUT the cuSOLVER eigensolver:
the command the ncu launch list is taken from.  (cuSOLVER's syevd launches ~10^5 small kernels per call; under ncu each
intercepted launch costs ~1.5 ms, so the full bench.py pass cannot be profiled in any reasonable time - measured: > 600 s
even for the 2000 x 3000 workload.  The eigensolver's share of a step is in bench.py's CUDA-event stage table instead.)
Usage: python scripts/own_kernels.py [B|C|small]"""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402

from bench import WORKLOADS, make_counts_fast  # noqa: E402
from sclens_b200 import Handle, _lib  # noqa: E402
from sclens_b200._lib import ptr  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "B"
N, M, seed = WORKLOADS[wl]
dev = torch.device("cuda", 0)
X = make_counts_fast(N, M, seed, device=dev)
nm, K = min(N, M), max(N, M)
rng = np.random.default_rng(0)


def stamp(name, t0):
    torch.cuda.synchronize()
    print(f"{name:28s} {1e3 * (time.perf_counter() - t0):9.2f} ms (host wall, incl. copies)", flush=True)


with Handle(seed=seed) as h:
    t0 = time.perf_counter(); h.set_counts(X); stamp("upload + CSR mirror", t0)
    for layout in ((1, 0) if N <= M else (0, 1)):          # Gram layout first, then the back-projection layout
        a, b, by = C.c_double(), C.c_double(), C.c_double()
        t0 = time.perf_counter()
        h._ck(h.lib.scl_bench_normalize(h.h, layout, 0, 1, C.byref(a), C.byref(b), C.byref(by)))
        stamp(f"normalise layout {layout}", t0)
        print(f"   stats {a.value:.3f} ms, writer {b.value:.3f} ms = {by.value / b.value / 1e6:.0f} GB/s")
    ms, cs = C.c_double(), C.c_double()
    t0 = time.perf_counter(); h._ck(h.lib.scl_bench_gram(h.h, nm, K, 0, 0, 1, C.byref(ms), C.byref(cs))); stamp("Gram (syrk)", t0)
    print(f"   {ms.value:.3f} ms = {nm * (nm + 1.0) * K / ms.value / 1e9:.0f} TFLOP/s (algorithmic n(n+1)K)")

    # perturbation merge (:735/:774): ~1 % of the grid, drawn from true zero positions
    n_try = int(0.011 * N * M)
    lin = np.unique(rng.integers(0, N * M, size=n_try, dtype=np.int64))
    r, c = (lin % N).astype(np.int64), (lin // N).astype(np.int64)
    nzlin = np.repeat(np.arange(M, dtype=np.int64), np.diff(X.indptr)) * N + X.indices.astype(np.int64)
    keep = ~np.isin(lin, nzlin, assume_unique=True)
    ar, ac = _lib.as_u32(r[keep]), _lib.as_u32(c[keep])
    colptr, rowval, val = _lib.as_u32(X.indptr), _lib.as_u32(X.indices), _lib.as_f32(X.data)
    n_out = X.nnz + len(ar)
    oc, orow, ov = np.empty(M + 1, np.uint32), np.empty(n_out, np.uint32), np.empty(n_out, np.float32)
    for binarise in (1, 0):
        t0 = time.perf_counter()
        h._ck(h.lib.scl_op_perturb_merge(h.h, N, M, X.nnz, ptr(colptr, C.c_uint32), ptr(rowval, C.c_uint32), ptr(val, C.c_float),
                                         len(ar), ptr(ar, C.c_uint32), ptr(ac, C.c_uint32), binarise, ptr(oc, C.c_uint32),
                                         ptr(orow, C.c_uint32), ptr(ov, C.c_float)))
        stamp(f"perturb merge (+{len(ar)})", t0)
    assert int(oc[-1]) == n_out
    # null matrix, device draws (:261-289)
    nnz = C.c_int64()
    t0 = time.perf_counter()
    h._ck(h.lib.scl_op_permute_null(h.h, N, M, X.nnz, ptr(colptr, C.c_uint32), ptr(rowval, C.c_uint32), ptr(val, C.c_float),
                                    None, None, C.byref(nnz), None, None, None))
    stamp("null matrix (device draws)", t0)

    # subspace iteration, corr column maxima, scores on synthetic operands of the workload's sizes
    n = nm
    A = torch.randn(n, 2 * n if n <= 12000 else n, device=dev, dtype=torch.float16)
    G = (A.float() @ A.float().T) / A.shape[1]
    del A
    U = torch.linalg.qr(torch.randn(n, 8, device=dev))[0]
    G = G + (U * torch.tensor([30., 20, 12, 8, 6, 5, 4.5, 4.2], device=dev)) @ U.T
    Gh = G.cpu().numpy()
    k = 11
    L, V = np.empty(k, np.float32), np.empty((k, n), np.float32)
    it = C.c_int32()
    t0 = time.perf_counter()
    h._ck(h.lib.scl_op_topk_subspace(h.h, n, ptr(Gh, C.c_float), k, ptr(L, C.c_float), ptr(V, C.c_float), C.byref(it)))
    stamp(f"top-{k} subspace ({it.value} GEMMs)", t0)
    nw = n // 2 + 1
    Q = torch.linalg.qr(torch.randn(n, n, device=dev))[0].T.contiguous().cpu().numpy()
    d = np.empty(nw, np.float32)
    t0 = time.perf_counter()
    h._ck(h.lib.scl_op_corr_colabsmax(h.h, n, n, nw, ptr(Q, C.c_float), ptr(Q[:nw].copy(), C.c_float), ptr(d, C.c_float)))
    stamp("corr column abs-max", t0)
    kk, mp, P = 7, 11, 20
    nV = np.linalg.qr(rng.standard_normal((N, kk)))[0].T.astype(np.float32).copy()
    sets = np.stack([np.linalg.qr(rng.standard_normal((N, mp)))[0].T for _ in range(P)]).astype(np.float32)
    b_ = np.empty((P * (P - 1) // 2, kk), np.float32)
    m, sd = np.empty(kk), np.empty(kk)
    sig, nr = np.empty(kk, np.int32), C.c_int32()
    t0 = time.perf_counter()
    h._ck(h.lib.scl_op_scores(h.h, N, kk, mp, P, ptr(nV, C.c_float), ptr(sets, C.c_float), 60.0, ptr(b_, C.c_float),
                              ptr(m, C.c_double), ptr(sd, C.c_double), ptr(sig, C.c_int32), C.byref(nr)))
    stamp("robustness scores", t0)
print("OWN_KERNELS_DONE")
