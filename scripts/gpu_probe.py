"""GPU probe: cuSOLVER syevd timings and Gram throughput at benchmark shapes (run under gpurun)."""
import ctypes as C
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402

from sclens_b200 import Handle  # noqa: E402
from sclens_b200._lib import ptr  # noqa: E402

out = {}
h = Handle()
for n in [int(a) for a in sys.argv[1:]] or [4096, 10000]:
    rng = np.random.default_rng(0)
    A = rng.standard_normal((n, n + 64), dtype=np.float32)
    G = (torch.from_numpy(A).cuda() @ torch.from_numpy(A).cuda().T / (n + 64)).cpu().numpy()
    L = np.empty(n, np.float32)
    V = np.empty((n, n), np.float32)
    ms = C.c_double()
    for vec in (True, False):
        for rep in range(2):
            h._ck(h.lib.scl_op_syevd(h.h, n, ptr(G, C.c_float), ptr(L, C.c_float), ptr(V, C.c_float) if vec else None, C.byref(ms)))
        out[f"syevd_n{n}_{'V' if vec else 'N'}_ms"] = ms.value
        print(n, vec, ms.value, flush=True)
print(json.dumps(out))
open("gpurun_out/probe.json", "w").write(json.dumps(out))
