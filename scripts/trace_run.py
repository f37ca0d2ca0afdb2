"""One pass at workload B with SCL_TRACE=1 (host-side phase times of the sparse merge) and the per-class CUDA-event totals."""
import sys
import time
sys.path.insert(0, ".")
import torch  # noqa: E402
from bench import WORKLOADS, make_counts_fast  # noqa: E402
from sclens_b200 import Handle  # noqa: E402
wl = sys.argv[1] if len(sys.argv) > 1 else "B"
N, M, seed = WORKLOADS[wl]
X = make_counts_fast(N, M, seed, device=torch.device("cuda", 0))
with Handle(seed=seed) as h:
    h.set_counts(X)
    for rep in range(2):
        h.reset_profile()
        t0 = time.perf_counter(); si = h.run_signal(); t1 = time.perf_counter()
        ri = h.run_robustness(n_perturb=int(sys.argv[2]) if len(sys.argv) > 2 else 3); t2 = time.perf_counter()
        p = h.profile()
        L = h.L()
        print("   L[:5]", L[:5], "L[-3:]", L[-3:], "n_signal", si.n_signal, "lambda_c", si.lambda_c)
        print(f"rep {rep}: signal {t1 - t0:.2f}s robustness {t2 - t1:.2f}s n_search {ri.n_search} | ms: gram {p.gram_gemm_ms:.1f} other_gemm {p.other_gemm_ms:.1f} "
              f"densify {p.densify_ms:.1f} ({p.densify_launches}) stats {p.stats_ms:.1f} sparse {p.sparse_ms:.1f} ({p.sparse_calls}) syevd {p.syevd_ms:.1f} ({p.syevd_calls}) "
              f"refine {p.refine_ms:.1f} small {p.small_ms:.1f}", flush=True)
