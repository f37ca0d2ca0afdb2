"""torchrun --nproc-per-node G scripts/multi_gpu_check.py : the G-rank cooperative sclens() pass (cell-sharded
Gram + NCCL all-reduce, speculative search waves, replicate-parallel perturbations) must reproduce the 1-GPU pass."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from sclens_b200 import Handle  # noqa: E402
from sclens_b200.synth import make_counts  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
shape = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1200, 900)
X = make_counts(*shape, seed=21, K=5, de_prob=0.3, lfc_sd=1.5)


def run(h, one_call=False):
    t0 = time.perf_counter()
    if one_call:      # scl_run_pass: the pass's eigensolves (data, null, reference basis, search steps) dealt to the ranks
        si, ri = h.run_pass(n_perturb=8)
    else:             # the two stage calls: every rank repeats the three serial solves
        si = h.run_signal()
        ri = h.run_robustness(n_perturb=8)
    dt = time.perf_counter() - t0
    b, m, sd = h.scores()
    return dict(L=h.L(), nL=h.signal_ev(), nV=h.signal_evec(), n_signal=si.n_signal, lam=si.lambda_c, p_sel=ri.p_sel,
                n_search=ri.n_search, trace=h.search_trace(), m=m, sig=h.sig_id(), dt=dt)


h = Handle(device=local, seed=7)
uid = torch.zeros(128, dtype=torch.uint8, device=dev)
if rank == 0:
    buf = (C.c_uint8 * 128)()
    assert h.lib.scl_nccl_unique_id(buf) == 0
    uid = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
dist.broadcast(uid, 0)
h.comm_init(bytes(uid.cpu().tolist()), rank, world)
h.set_counts(X)
staged = run(h)
multi = run(h, one_call=True)
multi2 = run(h, one_call=True)
dist.barrier()
ok = True
if rank == 0:
    with Handle(device=local, seed=7) as h1:
        h1.set_counts(X)
        single = run(h1)
        single = run(h1)
    print(f"world={world} shape={shape} multi {multi2['dt']:.2f}s single {single['dt']:.2f}s n_signal {multi['n_signal']} "
          f"p_sel {multi['p_sel']} n_search {multi['n_search']}", flush=True)
    for name, got in (("scl_run_pass", multi), ("scl_run_pass again", multi2), ("scl_run_signal + scl_run_robustness", staged)):
        assert got["n_signal"] == single["n_signal"] and got["n_search"] == single["n_search"] and got["p_sel"] == single["p_sel"], name
        np.testing.assert_allclose(got["L"][10:], single["L"][10:], rtol=2e-5, err_msg=name)
        np.testing.assert_allclose(got["trace"][1], single["trace"][1], rtol=2e-2, err_msg=name)
        np.testing.assert_allclose(got["m"], single["m"], atol=2e-2, err_msg=name)
        np.testing.assert_array_equal(got["sig"], single["sig"], err_msg=name)
        cos = np.abs(np.sum(got["nV"] * single["nV"], axis=0))
        assert cos.min() > 1 - 1e-5, name
    print("MULTI_GPU_CHECK_OK", flush=True)
# the shared pass against the ORACLE with injected draws (every rank regenerates the same seeded oracle run and injects its
# draws; the stop rule of the search is fed by steps that different ranks evaluated)
if len(sys.argv) <= 2:
    from oracle import sclens_oracle as orc   # test infrastructure: this script is a test, not the product path
    from sclens_b200 import SCL_GRAM_FP16X3, sclens
    Xo = make_counts(900, 640, seed=21, K=5, de_prob=0.3, lfc_sd=1.5)
    ref, draws, info = orc.sclens(Xo, rng=np.random.default_rng(7), mode="cpu", n_perturb=6, n_baseline=300)
    with Handle(device=local, seed=7, gram_mode=SCL_GRAM_FP16X3, exact_perturb=True) as ho:
        uid2 = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            buf = (C.c_uint8 * 128)()
            assert ho.lib.scl_nccl_unique_id(buf) == 0
            uid2 = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
        dist.broadcast(uid2, 0)
        out = sclens(Xo, draws=draws, n_perturb=6, gram_mode=SCL_GRAM_FP16X3, exact_perturb=True, verbose=False, handle=ho,
                     comm=(bytes(uid2.cpu().tolist()), rank, world))
    assert len(out["signal_ev"]) == len(ref["signal_ev"]), (rank, len(out["signal_ev"]), len(ref["signal_ev"]))
    np.testing.assert_allclose(out["signal_ev"], ref["signal_ev"], rtol=1e-4)
    assert out["info"]["n_search"] == info["n_search"] and out["info"]["p_sel"] == info["p_sel"], rank
    np.testing.assert_array_equal(out["sig_id"], ref["sig_id"])
    np.testing.assert_allclose(out["robustness_scores"]["m_scores"], ref["robustness_scores"]["m_scores"], atol=2e-2)
    cos = np.abs(np.sum(out["signal_evec"] * np.asarray(ref["signal_evec"]), axis=0))
    assert cos.min() > 1 - 1e-5, (rank, cos)
    if rank == 0:
        print("MULTI_GPU_ORACLE_OK", flush=True)
# every rank must hold identical results
t = torch.tensor([multi["p_sel"], float(multi["n_search"]), float(len(multi["sig"])), float(multi["m"].sum())], device=dev, dtype=torch.float64)
ref = t.clone()
dist.broadcast(ref, 0)
assert torch.equal(t, ref), (rank, t, ref)
h.close()
dist.destroy_process_group()
