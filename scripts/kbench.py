"""Kernel-level timings through the C ABI (scl_bench_*): the tcgen05 Gram kernel and the normalisation kernels
alone, on device-resident data.  Usage: python scripts/kbench.py gram|norm|all [B|C|small]"""
import ctypes as C
import json
import sys

sys.path.insert(0, ".")
import numpy as np  # noqa: E402
import torch  # noqa: E402,F401

from sclens_b200 import Handle  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
wl = sys.argv[2] if len(sys.argv) > 2 else "B"
SHAPES = {"B": (10000, 20000), "C": (68000, 20000), "small": (2000, 3000)}
N, M = SHAPES[wl]
peaks = {}
try:
    peaks = json.load(open("MEASURED_PEAKS.json"))
except Exception:
    pass
out = {}
with Handle(seed=0) as h:
    if what in ("gram", "all"):
        rows, K = min(N, M), max(N, M)
        chunks = [int(a) for a in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 32, 128, 1 << 20]
        for mode in (0, 1):
            for chunk in chunks:
                ms, cs = C.c_double(), C.c_double()
                h._ck(h.lib.scl_bench_gram(h.h, rows, K, mode, chunk, 3 if mode == 0 else 1, C.byref(ms), C.byref(cs)))
                tf = rows * (rows + 1.0) * K / (ms.value * 1e-3) / 1e12
                key = f"gram_{wl}_mode{mode}_chunk{chunk}"
                out[key] = {"ms": ms.value, "alg_TFLOPs": tf, "frac_sustained": tf / peaks.get("bf16_tflops_sustained", 1420.4),
                            "mean_diag": cs.value}
                print(key, out[key], flush=True)
    if what in ("norm", "all"):
        from bench import make_counts_fast
        import torch
        X = make_counts_fast(N, M, {"B": 0, "C": 1, "small": 3}[wl], device=torch.device("cuda", 0))
        h.set_counts(X)
        for layout in (0, 1):
            for with_lo in (0, 1):
                a, b, by = C.c_double(), C.c_double(), C.c_double()
                h._ck(h.lib.scl_bench_normalize(h.h, layout, with_lo, 5, C.byref(a), C.byref(b), C.byref(by)))
                key = f"norm_{wl}_layout{layout}_lo{with_lo}"
                out[key] = {"stats_ms": a.value, "densify_ms": b.value, "densify_GBs": by.value / b.value / 1e6,
                            "frac_hbm": by.value / b.value / 1e6 / peaks.get("hbm_gbs", 6549.1),
                            "stats_GBs_5pass": 5 * 8.0 * X.nnz / a.value / 1e6}
                print(key, out[key], flush=True)
open(f"gpurun_out/kbench_{what}_{wl}.json", "w").write(json.dumps(out, indent=1))
