"""Where does the eigenvalue error at 10k x 20k come from?  Separates operand rounding (binary16 hi),
tensor-core accumulation, and the FP32 eigensolver (run under gpurun)."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402

from sclens_b200 import Handle, _lib  # noqa: E402
from sclens_b200._lib import ptr  # noqa: E402
from sclens_b200.synth import make_counts  # noqa: E402

N, M = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (10000, 20000)
X = make_counts(N, M, seed=0)
with Handle(seed=0) as h:
    h.set_counts(X)
    si = h.run_signal()
    L_pipe = h.L().astype(np.float64)
    ld = M
    hi = np.empty((N, ld), np.uint16)
    lo = np.empty((N, ld), np.uint16)
    colptr, rowval, val = _lib.as_u32(X.indptr), _lib.as_u32(X.indices), _lib.as_f32(X.data)
    h._ck(h.lib.scl_op_normalize(h.h, N, M, X.nnz, ptr(colptr, C.c_uint32), ptr(rowval, C.c_uint32),
                                 ptr(val, C.c_float), 1, ld, ptr(hi, C.c_uint16), ptr(lo, C.c_uint16),
                                 None, None, None, None, None))
    G_ours = np.empty((N, N), np.float32)
    h._ck(h.lib.scl_op_gram(h.h, N, M, ld, ptr(hi, C.c_uint16), None, 1.0 / M, ptr(G_ours, C.c_float)))
    G_ours3 = np.empty((N, N), np.float32)
    h._ck(h.lib.scl_op_gram(h.h, N, M, ld, ptr(hi, C.c_uint16), ptr(lo, C.c_uint16), 1.0 / M, ptr(G_ours3, C.c_float)))
    Ahi = torch.from_numpy(hi.view(np.float16)).cuda().double()
    A = Ahi + torch.from_numpy(lo.view(np.float16)).cuda().double()
    G64 = A @ A.T / M
    Lref = torch.linalg.eigvalsh(G64).cpu().numpy()
    keep = Lref >= si.b_minus

    def rep(name, L):
        rel = np.abs(L[keep] - Lref[keep]) / Lref[keep]
        ab = np.abs(L[keep] - Lref[keep])
        i = int(np.argmax(rel))
        print(f"{name:34s} max rel {rel.max():.3e} at lambda={Lref[keep][i]:.4f} (idx {i}/{keep.sum()}), "
              f"median rel {np.median(rel):.2e}, max abs {ab.max():.3e} at lambda={Lref[keep][int(np.argmax(ab))]:.3f}", flush=True)

    print("b_minus", si.b_minus, "b_plus", si.b_plus, "lambda_c", si.lambda_c, "Lmax", Lref[-1])
    rep("pipeline L (fp16 gram + Ssyevd)", L_pipe)
    Ghi64 = Ahi @ Ahi.T / M
    rep("hi-only operand, f64 all", torch.linalg.eigvalsh(Ghi64).cpu().numpy())
    Go = torch.from_numpy(G_ours).cuda()
    print("our gram vs hi-only f64 gram: max abs", float((Go.double() - Ghi64).abs().max()), "diag max rel",
          float(((Go.double().diagonal() - Ghi64.diagonal()).abs() / Ghi64.diagonal()).max()),
          "mean signed rel offdiag", float(((Go.double() - Ghi64) * Ghi64.sign()).mean() / Ghi64.abs().mean()))
    rep("our fp16 gram, f64 eigvalsh", torch.linalg.eigvalsh(Go.double()).cpu().numpy())
    Go3 = torch.from_numpy(G_ours3).cuda()
    print("our x3 gram vs f64 gram: max abs", float((Go3.double() - G64).abs().max()))
    rep("our fp16x3 gram, f64 eigvalsh", torch.linalg.eigvalsh(Go3.double()).cpu().numpy())
    G32 = G64.float()
    rep("f64 gram -> f32, torch f32 eigvalsh", torch.linalg.eigvalsh(G32).double().cpu().numpy())
    Lo = np.empty(N, np.float32)
    ms = C.c_double()
    h._ck(h.lib.scl_op_syevd(h.h, N, ptr(G32.cpu().numpy(), C.c_float), ptr(Lo, C.c_float), None, C.byref(ms)))
    rep("f64 gram -> f32, our Ssyevd (N)", Lo.astype(np.float64))
    V = np.empty((N, N), np.float32)
    h._ck(h.lib.scl_op_syevd(h.h, N, ptr(G32.cpu().numpy(), C.c_float), ptr(Lo, C.c_float), ptr(V, C.c_float), C.byref(ms)))
    rep("f64 gram -> f32, our Ssyevd (V)", Lo.astype(np.float64))
    # Rayleigh-quotient refinement: lambda_i = v_i^T G v_i with fp32 vectors, f64 arithmetic
    Vt = torch.from_numpy(V).cuda().double()          # rows of memory = eigenvectors (column-major result)
    rq = ((Vt @ G64) * Vt).sum(dim=1) / (Vt * Vt).sum(dim=1)
    rep("Rayleigh quotients of Ssyevd vecs", rq.cpu().numpy())
