"""The own HBM / tensor kernels of one pass at a workload's shapes, once each, with nothing else around them (no
eigensolver, no torch.linalg): the command the round-2 ncu captures are taken from.
Usage: python scripts/prof_own.py [C|B|small]"""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402

from bench import WORKLOADS, make_counts_fast  # noqa: E402
from sclens_b200 import Handle, _lib  # noqa: E402
from sclens_b200._lib import ptr  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "C"
N, M, seed = WORKLOADS[wl]
X = make_counts_fast(N, M, seed, device=torch.device("cuda", 0))
nm, K = min(N, M), max(N, M)
with Handle(seed=seed) as h:
    h.set_counts(X)
    for layout in ((1, 0) if N <= M else (0, 1)):
        a, b, by = C.c_double(), C.c_double(), C.c_double()
        h._ck(h.lib.scl_bench_normalize(h.h, layout, 0, 1, C.byref(a), C.byref(b), C.byref(by)))
        print(f"normalise layout {layout}: stats {a.value:.3f} ms, writer {b.value:.3f} ms = {by.value / b.value / 1e6:.0f} GB/s", flush=True)
    ms, cs = C.c_double(), C.c_double()
    h._ck(h.lib.scl_bench_gram(h.h, nm, K, 0, 0, 1, C.byref(ms), C.byref(cs)))
    print(f"Gram {ms.value:.3f} ms = {nm * (nm + 1.0) * K / ms.value / 1e9:.0f} TFLOP/s (algorithmic n(n+1)K)", flush=True)
    # zero candidates drawn on the device, then two perturbation merges of 1 % of the grid through the pipeline's own path
    n = C.c_int64()
    h._ck(h.lib.scl_op_draw_zero_candidates(h.h, seed, C.byref(n), None, None))
    n_add = int(round(0.01 * N * M))
    r, c = np.empty(n_add, np.uint32), np.empty(n_add, np.uint32)
    h._ck(h.lib.scl_op_draw_subset(h.h, n_add, 7, ptr(r, C.c_uint32), ptr(c, C.c_uint32)))
    colptr, rowval, val = _lib.as_u32(X.indptr), _lib.as_u32(X.indices), _lib.as_f32(X.data)
    n_out = X.nnz + n_add
    oc, orow, ov = np.empty(M + 1, np.uint32), np.empty(n_out, np.uint32), np.empty(n_out, np.float32)
    for binarise in (1, 0):
        h._ck(h.lib.scl_op_perturb_merge(h.h, N, M, X.nnz, ptr(colptr, C.c_uint32), ptr(rowval, C.c_uint32), ptr(val, C.c_float),
                                         n_add, ptr(r, C.c_uint32), ptr(c, C.c_uint32), binarise, ptr(oc, C.c_uint32),
                                         ptr(orow, C.c_uint32), ptr(ov, C.c_float)))
    assert int(oc[-1]) == n_out
    nnz = C.c_int64()
    h._ck(h.lib.scl_op_permute_null(h.h, N, M, X.nnz, ptr(colptr, C.c_uint32), ptr(rowval, C.c_uint32), ptr(val, C.c_float),
                                    None, None, C.byref(nnz), None, None, None))
print("PROF_OWN_DONE")
