"""ncu target: one normalisation (statistics line passes + dense writer) per writer variant at one workload's shape.
python scripts/prof_norm.py [B|C]"""
import ctypes as C
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402

from bench import WORKLOADS, make_counts_fast  # noqa: E402
from sclens_b200 import Handle  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "B"
N, M, seed = WORKLOADS[wl]
X = make_counts_fast(N, M, seed, device=torch.device("cuda", 0))
layout = 0 if N > M else 1
with Handle(seed=0) as h:
    h.set_counts(X)
    for lo in (0, 1):      # default tuning: line passes (variant 4), automatic writer
        a, b, by = C.c_double(), C.c_double(), C.c_double()
        h._ck(h.lib.scl_bench_normalize(h.h, layout, lo, 1, C.byref(a), C.byref(b), C.byref(by)))
        print(f"lo {lo}: stats {a.value:.3f} ms, writer {b.value:.3f} ms = {by.value / b.value / 1e6:.0f} GB/s", flush=True)
