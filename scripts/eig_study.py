"""Library eigensolver study on an n x n Wishart matrix (ctypes only): python scripts/eig_study.py [n] [out.json]

Single solves (warmed): Ssyevd with vectors / values only, Ssyevdx smallest half, Xsyevd, Ssytrd alone, Sormtr alone;
then 2 and 3 independent Ssyevd solves issued concurrently from host threads on separate streams of ONE GPU."""
import ctypes as C
import json
import sys

sys.path.insert(0, ".")
from sclens_b200 import Handle  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
path = sys.argv[2] if len(sys.argv) > 2 else f"gpurun_out/eig_study_{n}.json"
out = {"n": n}
with Handle() as h:
    singles = (("warmup_values_only", 1, 0, 0), ("Ssyevd_vectors", 0, 0, 0), ("Ssyevd_vectors_2", 0, 0, 0),
               ("Ssyevd_values_only", 1, 0, 0), ("Ssyevdx_smallest_half_vectors", 2, 1, n // 2 + 1),
               ("Ssyevdx_smallest_half_vectors_2", 2, 1, n // 2 + 1), ("Xsyevd_vectors", 3, 0, 0),
               ("Xsyevd_vectors_2", 3, 0, 0), ("Ssytrd_alone", 4, 0, 0), ("Sormtr_alone", 5, 0, 0))
    for name, mode, il, iu in singles:
        ms = C.c_double()
        try:
            h._ck(h.lib.scl_bench_syevd(h.h, n, mode, il, iu, C.byref(ms)))
            out[name] = ms.value
        except Exception as e:  # a variant the library rejects is a result too
            out[name] = str(e)
        print(name, out[name], flush=True)
    for name, k, mode in (("concurrent_1_vectors", 1, 0), ("concurrent_2_vectors", 2, 0), ("concurrent_3_vectors", 3, 0),
                          ("concurrent_2_values_only", 2, 1)):
        ms = C.c_double()
        try:
            h._ck(h.lib.scl_bench_syevd_concurrent(h.h, n, k, mode, C.byref(ms)))
            out[name] = {"wall_ms": ms.value, "ms_per_solve": ms.value / k}
        except Exception as e:
            out[name] = str(e)
        print(name, out[name], flush=True)
open(path, "w").write(json.dumps(out, indent=1))
