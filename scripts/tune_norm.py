"""Tuning study of the normalisation kernels (statistics line passes + dense writer) on one workload, one process:
python scripts/tune_norm.py [B|C].  Writes gpurun_out/tune_norm_<wl>.json."""
import ctypes as C
import json
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402

from bench import WORKLOADS, make_counts_fast  # noqa: E402
from sclens_b200 import Handle  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "B"
N, M, seed = WORKLOADS[wl]
X = make_counts_fast(N, M, seed, device=torch.device("cuda", 0))
layout = 0 if N > M else 1
out = {}
with Handle(seed=0) as h:
    h.set_counts(X)

    def run(tag, lo=0):
        a, b, by = C.c_double(), C.c_double(), C.c_double()
        h._ck(h.lib.scl_bench_normalize(h.h, layout, lo, 5, C.byref(a), C.byref(b), C.byref(by)))
        out[tag] = {"stats_ms": a.value, "densify_ms": b.value, "densify_GBs": by.value / b.value / 1e6}
        print(tag, out[tag], flush=True)

    for variant in (4, 8):
        for heavy in (4096,):
            h.lib.scl_debug_set_tuning(variant, heavy, 0)
            run(f"stats_v{variant}_heavy{heavy}")
    h.lib.scl_debug_set_tuning(8, 4096, 0)
    for writer in (0, 2, -1):
        for lo in (0, 1):
            h.lib.scl_debug_set_tuning(-1, 0, writer)
            run(f"strips_writer{writer}_lo{lo}", lo)
open(f"gpurun_out/tune_norm_{wl}.json", "w").write(json.dumps(out, indent=1))
