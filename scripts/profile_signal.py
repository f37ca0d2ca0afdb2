"""Short single-GPU run for ncu: the signal stage (normalise, Gram, syevd, null, fit) of one workload."""
import sys
import time

sys.path.insert(0, ".")
import numpy as np  # noqa: E402

from bench import WORKLOADS, make_counts_fast  # noqa: E402
from sclens_b200 import Handle  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "B"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
N, M, seed = WORKLOADS[wl]
import torch  # noqa: E402
X = make_counts_fast(N, M, seed, device=torch.device("cuda", 0))
with Handle(seed=seed, gram_mode=int(sys.argv[3]) if len(sys.argv) > 3 else 0) as h:
    h.set_counts(X)
    for _ in range(reps):
        t0 = time.perf_counter()
        si = h.run_signal()
        p = h.profile()
        print(f"signal stage {time.perf_counter() - t0:.2f}s n_signal={si.n_signal} lambda_c={si.lambda_c:.5f} "
              f"gram_ms={p.gram_gemm_ms / max(1, p.gram_gemm_launches):.3f} TF={p.gram_alg_flops / max(1e-9, p.gram_gemm_ms) / 1e9:.1f} "
              f"densify_ms={p.densify_ms / max(1, p.densify_launches):.3f} GB/s={p.densify_alg_bytes / max(1e-9, p.densify_ms) / 1e6:.0f} "
              f"stats_ms={p.stats_ms / max(1, p.gram_gemm_launches):.3f} sparse_ms={p.sparse_ms:.2f} syevd_ms={p.syevd_ms / max(1, p.syevd_calls):.1f}", flush=True)
        h.reset_profile()
