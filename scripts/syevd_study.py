"""Library eigensolver variants on an n x n Wishart matrix (ctypes only, no torch): python scripts/syevd_study.py [n]"""
import ctypes as C
import json
import sys

sys.path.insert(0, ".")
from sclens_b200 import Handle  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
out = {}
with Handle() as h:
    for name, mode, il, iu in (("warmup_values_only", 1, 0, 0), ("Ssyevd_vectors", 0, 0, 0), ("Ssyevd_values_only", 1, 0, 0),
                               ("Ssyevdx_smallest_half_vectors", 2, 1, n // 2 + 1), ("Xsyevd_vectors", 3, 0, 0)):
        ms = C.c_double()
        try:
            h._ck(h.lib.scl_bench_syevd(h.h, n, mode, il, iu, C.byref(ms)))
            out[name] = ms.value
        except Exception as e:  # a variant the library rejects is a result too
            out[name] = str(e)
        print(name, out[name], flush=True)
open(f"gpurun_out/syevd_study_{n}.json", "w").write(json.dumps(out, indent=1))
