"""Why does one low-lying eigenvalue of the 10k x 10k benchmark Gram carry a 1.8e-4 relative error from every FP32
eigensolver?  Looks at that eigenpair in Float64: Rayleigh quotients of the FP32 eigenvectors, residuals,
localisation of the true eigenvector (run under gpurun)."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402

from sclens_b200 import Handle, _lib  # noqa: E402
from sclens_b200._lib import ptr  # noqa: E402
from sclens_b200.synth import make_counts  # noqa: E402

N, M = 10000, 20000
X = make_counts(N, M, seed=0)
colptr, rowval, val = _lib.as_u32(X.indptr), _lib.as_u32(X.indices), _lib.as_f32(X.data)
hi = np.zeros((N, M), np.uint16)
lo = np.zeros((N, M), np.uint16)
with Handle(seed=0) as h:
    h._ck(h.lib.scl_op_normalize(h.h, N, M, X.nnz, ptr(colptr, C.c_uint32), ptr(rowval, C.c_uint32), ptr(val, C.c_float),
                                 1, M, ptr(hi, C.c_uint16), ptr(lo, C.c_uint16), None, None, None, None, None))
    A = torch.from_numpy(hi.view(np.float16)).cuda().double() + torch.from_numpy(lo.view(np.float16)).cuda().double()
    G64 = A @ A.T / M
    del A
    Lref, Vref = torch.linalg.eigh(G64)
    G32 = G64.float()
    Lo = np.empty(N, np.float32)
    V = np.empty((N, N), np.float32)
    ms = C.c_double()
    h._ck(h.lib.scl_op_syevd(h.h, N, ptr(G32.cpu().numpy(), C.c_float), ptr(Lo, C.c_float), ptr(V, C.c_float), C.byref(ms)))
Lr = Lref.cpu().numpy()
rel = (Lo.astype(np.float64) - Lr) / np.maximum(Lr, 1e-3)
worst = np.argsort(-np.abs(rel[3:]))[:6] + 3
Vt = torch.from_numpy(V).cuda().double()                  # rows = eigenvectors
G32d = G32.double()
Z = Vt @ G32d
rq = ((Z * Vt).sum(1) / (Vt * Vt).sum(1)).cpu().numpy()
res = (Z - torch.from_numpy(rq).cuda()[:, None] * Vt).norm(dim=1).cpu().numpy()
ortho = (Vt[worst] @ Vt.T)
print("largest eigenvalue", Lr[-1], "; eigenvalue spacing near the outliers", np.diff(Lr)[worst])
for i in worst:
    vi = Vref[:, i]
    pr = float(1.0 / (vi ** 4).sum())                     # participation ratio (N = delocalised, 1 = one cell)
    big = torch.topk(vi.abs(), 3)
    o = ortho[list(worst).index(i)].clone()
    o[i] = 0
    j = int(o.abs().argmax())
    print(f"idx {i}: lambda_ref {Lr[i]:.6f} ssyevd rel {rel[i]:+.2e}  rq(f64 of f32 vec) rel {(rq[i] - Lr[i]) / Lr[i]:+.2e}  "
          f"residual |Gv-rho v| {res[i]:.2e}  participation {pr:.0f}  top |v| {big.values.cpu().numpy().round(3)} at {big.indices.cpu().numpy()}  "
          f"max non-orthogonality {float(o[j]):+.1e} with idx {j} (lambda {Lr[j]:.4f})  overlap with true vec {float((Vt[i] @ vi).abs()):.6f}")
print("median |rel| ssyevd", np.median(np.abs(rel[3:])), " median |rel| rq", np.median(np.abs((rq[3:] - Lr[3:]) / Lr[3:])),
      " max |rel| rq", np.max(np.abs((rq[3:] - Lr[3:]) / Lr[3:])), "at", int(np.argmax(np.abs((rq[3:] - Lr[3:]) / Lr[3:]))) + 3)
