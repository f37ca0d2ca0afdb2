"""Where does the eigenvalue error at 10k x 20k come from?  Separates operand rounding (binary16 hi), the
tensor core's truncating accumulation (as a function of the promotion chunk length, with the exact Float64
diagonal or the tensor core's own), and the FP32 eigensolver (run under gpurun).
Usage: python scripts/diag_eig_error.py [N M] [chunks, comma separated]"""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402

from sclens_b200 import Handle, _lib  # noqa: E402
from sclens_b200._lib import ptr  # noqa: E402
from sclens_b200.synth import make_counts  # noqa: E402

N, M = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (10000, 20000)
chunks = [int(c) for c in sys.argv[3].split(",")] if len(sys.argv) > 3 else [64, 32, 16, 8]
X = make_counts(N, M, seed=0)
colptr, rowval, val = _lib.as_u32(X.indptr), _lib.as_u32(X.indices), _lib.as_f32(X.data)
tall = N > M
rows, K = (M, N) if tall else (N, M)
ld = (K + 7) // 8 * 8
hi = np.zeros((rows, ld), np.uint16)
lo = np.zeros((rows, ld), np.uint16)
with Handle(seed=0) as h:
    h.set_counts(X)
    si = h.run_signal()
    h._ck(h.lib.scl_op_normalize(h.h, N, M, X.nnz, ptr(colptr, C.c_uint32), ptr(rowval, C.c_uint32),
                                 ptr(val, C.c_float), 0 if tall else 1, ld, ptr(hi, C.c_uint16), ptr(lo, C.c_uint16),
                                 None, None, None, None, None))
Ahi = torch.from_numpy(hi.view(np.float16)).cuda().double()
A = Ahi + torch.from_numpy(lo.view(np.float16)).cuda().double()
G64 = A @ A.T / M
Lref = torch.linalg.eigvalsh(G64).cpu().numpy()
keep = Lref >= si.b_minus
Ghi64 = Ahi @ Ahi.T / M
del A
offmask = ~torch.eye(rows, dtype=torch.bool, device="cuda")


def rep(name, L):
    L = np.asarray(L, dtype=np.float64)
    rel = np.abs(L[keep] - Lref[keep]) / Lref[keep]
    i = int(np.argmax(rel))
    srel = (L[keep] - Lref[keep]) / Lref[keep]
    q = [0, len(rel) // 4, len(rel) // 2, 3 * len(rel) // 4, len(rel) - 8, len(rel) - 1]
    print(f"  {name:46s} max rel {rel.max():.2e} at lambda={Lref[keep][i]:.4f} (idx {i}/{keep.sum()}), median {np.median(rel):.1e}; "
          f"signed rel at idx {q}: " + " ".join(f"{srel[j]:+.1e}" for j in q), flush=True)


print("b_minus", si.b_minus, "b_plus", si.b_plus, "lambda_c", si.lambda_c, "Lmax", Lref[-1], "n_signal", si.n_signal)
rep("hi-only operand, f64 gram + f64 eig", torch.linalg.eigvalsh(Ghi64).cpu().numpy())
G32 = G64.float()
rep("f64 gram -> f32, torch f32 eigvalsh", torch.linalg.eigvalsh(G32).double().cpu().numpy())
with Handle(seed=0) as h:
    Lo = np.empty(rows, np.float32)
    ms = C.c_double()
    h._ck(h.lib.scl_op_syevd(h.h, rows, ptr(G32.cpu().numpy(), C.c_float), ptr(Lo, C.c_float), None, C.byref(ms)))
    rep("f64 gram -> f32, cuSOLVER Ssyevd ('N')", Lo)
del G32

for chunk in chunks:
    print(f"chunk_kb = {chunk} ({4 * chunk} MMAs per tensor-core accumulation chain)")
    with Handle(seed=0, gram_chunk_kb=chunk) as h:
        G = np.empty((rows, rows), np.float32)
        h._ck(h.lib.scl_op_gram(h.h, rows, K, ld, ptr(hi, C.c_uint16), None, 1.0 / M, ptr(G, C.c_float)))
        Go = torch.from_numpy(G).cuda().double()
        d = Go - Ghi64
        slope_off = float((d[offmask] * Ghi64[offmask]).sum() / (Ghi64[offmask] ** 2).sum())
        slope_diag = float((d.diagonal() * Ghi64.diagonal()).sum() / (Ghi64.diagonal() ** 2).sum())
        resid = d - slope_off * Ghi64
        print(f"  tensor-core Gram vs f64 Gram of the same operand: off-diagonal shrink {slope_off:+.3e}, diagonal shrink "
              f"{slope_diag:+.3e}, mean off-diagonal offset {float(d[offmask].mean()):+.3e}, rms residual off-diag after removing the "
              f"shrink {float(resid[offmask].pow(2).mean().sqrt()):.2e} (rms off-diag {float(Ghi64[offmask].pow(2).mean().sqrt()):.2e}, "
              f"mean |off-diag| {float(Ghi64[offmask].abs().mean()):.2e})", flush=True)
        rep("tensor-core diagonal, f64 eig", torch.linalg.eigvalsh(Go).cpu().numpy())
        Ge = Go.clone()
        Ge.diagonal().copy_(Ghi64.diagonal())
        rep("exact diagonal, f64 eig", torch.linalg.eigvalsh(Ge).cpu().numpy())
        Gc2 = Go / (1.0 + slope_off)
        Gc2.diagonal().copy_(Ghi64.diagonal())
        rep("exact diag, off-diag / (1+off-diag shrink), f64 eig", torch.linalg.eigvalsh(Gc2).cpu().numpy())
        del Go, Ge, Gc2, d, resid
    for tcd in (0, 2):
        with Handle(seed=0, gram_chunk_kb=chunk, gram_tc_diag=tcd) as h:
            h.set_counts(X)
            s2 = h.run_signal()
            rep(f"pipeline L, mode={tcd} (0 calibrated, 2 exact diag only), n_signal={s2.n_signal}", h.L())
