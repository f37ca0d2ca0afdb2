"""Own tridiagonal stage vs the library solve, on the synthetic Wishart matrix and on the data Gram matrix of workload C
(68k x 20k): python scripts/eig_study2.py [out.json]"""
import ctypes as C
import json
import sys
import time

sys.path.insert(0, ".")
import torch  # noqa: E402
from bench import WORKLOADS, make_counts_fast  # noqa: E402
from sclens_b200 import Handle  # noqa: E402

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r2_eig_study2.json"
wl = sys.argv[2] if len(sys.argv) > 2 else "C"
out = {}
N, M, seed = WORKLOADS[wl]
n = min(N, M)
dev = torch.device("cuda", 0)
with Handle(seed=seed) as h:
    def bench(mode, il=0, iu=0):
        ms = C.c_double()
        h._ck(h.lib.scl_bench_syevd(h.h, n, mode, il, iu, C.byref(ms)))
        return ms.value

    bench(1)   # warm-up
    for name, mode, il, iu in (("synthetic/Ssyevd_vectors", 0, 0, 0), ("synthetic/tri_all_vectors", 6, 0, 0),
                               ("synthetic/tri_all_vectors_2", 6, 0, 0),
                               ("synthetic/tri_smallest_half", 7, 1, n // 2 + 65), ("synthetic/tri_values_only", 8, 0, 0)):
        out[name] = bench(mode, il, iu)
        print(name, out[name], flush=True)
    t0 = time.perf_counter()
    X = make_counts_fast(N, M, seed, device=dev)
    h.set_counts(X)
    si = h.run_signal()
    print("run_signal", time.perf_counter() - t0, "n_signal", si.n_signal, flush=True)
    for name, mode, il, iu in (("data_gram/Ssyevd_vectors", 16, 0, 0), ("data_gram/Ssyevd_values_only", 17, 0, 0),
                               ("data_gram/Ssytrd_alone", 20, 0, 0), ("data_gram/tri_all_vectors", 22, 0, 0),
                               ("data_gram/tri_smallest_half", 23, 1, n // 2 + 65), ("data_gram/tri_values_only", 24, 0, 0)):
        out[name] = bench(mode, il, iu)
        print(name, out[name], flush=True)
open(path, "w").write(json.dumps(out, indent=1))
