"""SCL_TRACE=1 python scripts/sytrd_trace.py [n]: section times of the own tridiagonalisation (stderr)."""
import ctypes as C
import sys
sys.path.insert(0, ".")
from sclens_b200 import Handle  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
with Handle() as h:
    for rep in range(2):
        ms = C.c_double()
        h._ck(h.lib.scl_bench_syevd(h.h, n, 9, 0, 0, C.byref(ms)))
        print("own sytrd", n, ms.value, "ms", flush=True)
