"""Own tridiagonalisation (sytrd.cu) against cusolverDnSsytrd at n = 20 000 (and the solve built on each):
python scripts/eig_study3.py [n] [out.json]"""
import ctypes as C
import json
import sys

sys.path.insert(0, ".")
from sclens_b200 import Handle  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
path = sys.argv[2] if len(sys.argv) > 2 else f"gpurun_out/r2_eig_study3_{n}.json"
out = {"n": n}
with Handle() as h:
    def bench(mode, il=0, iu=0):
        ms = C.c_double()
        h._ck(h.lib.scl_bench_syevd(h.h, n, mode, il, iu, C.byref(ms)))
        return ms.value

    bench(1)
    for name, api, mode, il, iu in (("Ssytrd_alone", -1, 4, 0, 0), ("own_sytrd_alone", -1, 9, 0, 0), ("own_sytrd_alone_2", -1, 9, 0, 0),
                                    ("tri_all_vectors/Ssytrd", 12, 6, 0, 0), ("tri_all_vectors/own_sytrd", 28, 6, 0, 0),
                                    ("tri_smallest_half/own_sytrd", 28, 7, 1, n // 2 + 65),
                                    ("tri_values_only/own_sytrd", 28, 8, 0, 0), ("Ssyevd_vectors", 0, 0, 0, 0)):
        h.lib.scl_debug_set_eig_api(api)
        try:
            out[name] = bench(mode, il, iu)
        except Exception as e:
            out[name] = str(e)
        print(name, out[name], flush=True)
    h.lib.scl_debug_set_eig_api(-1)
open(path, "w").write(json.dumps(out, indent=1))
