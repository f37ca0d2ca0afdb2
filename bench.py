#!/usr/bin/env python
"""bench.py - sclens() throughput on synthetic data of the BASELINE.json shapes.

    python bench.py --gpus N --steps K --warmup W [--workload C|B|small] [--impl reference]

Workloads: C = 68k cells x 20k genes (BASELINE.json configs[2], the shape the metric is quoted on; default),
B = 10k x 20k (configs[1]), small = 2k x 3k (smoke).

One "step" = one complete sclens() pass (signal detection + robustness test, n_perturb=20) over
the synthetic count matrix of the workload.  Our arm:
  value  cells/s with the CSC counts already resident in HBM (library stream CUDA events)
  e2e    cells/s through the public API sclens_b200.sclens(X_host, handle=h): pinned host -> device copy of
         the CSC, the whole path, device -> host read of every result; the rank's handle (library context) is
         reused, a cold call that creates its own handle is reported beside it (B / small workloads)
  roofline   the dominant own kernel, the tcgen05 Gram: algorithmic n(n+1)K FLOP per launch over its
             CUDA-event time, against MEASURED_PEAKS.json; the HBM-bound densify kernel beside it
  cpu_baseline   the oracle (reference device_="cpu" numerics) timed on the host cores on a bounded
             sample and scaled to one sclens() pass (what was sampled is spelled out)
`--impl reference` times that CPU implementation as the reference arm.
Under torchrun (N>1) every rank owns one GPU; ranks cooperate on ONE sclens() pass ("strong").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (N cells, M genes, seed)  -- BASELINE.json configs[1], configs[2]
    "B": (10000, 20000, 0),
    "C": (68000, 20000, 1),
    "small": (2000, 3000, 3),
}
# search steps the reference arm assumes (it runs before any GPU pass exists): the count our own pass takes on the
# same synthetic matrix (deterministic per workload and seed; measured on B200, profiles/r1_bench_*.json)
NOMINAL_SEARCH_STEPS = {"B": 14, "C": 14, "small": 14}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------------------
# synthetic counts (same recipe as sclens_b200.synth.make_counts; Poisson draws on the GPU when there is one)
# --------------------------------------------------------------------------------------
def make_counts_fast(N, M, seed, device=None):
    from sclens_b200 import synth
    if device is None:
        return synth.make_counts(N, M, seed=seed)
    import torch
    rng = np.random.default_rng(seed)
    K = 8
    prof = synth._profiles(M, K, rng, 0.10, 1.0)
    types = rng.integers(0, K, size=N)
    ldepth = rng.normal(0.0, 0.35, size=N)
    sub = rng.choice(N, size=min(N, 512), replace=False)
    min_cells = 40

    def floored(D):
        pmin = -np.log1p(-min(0.5, min_cells / N)) / D
        p = np.maximum(prof, pmin)
        return p / p.sum(axis=1, keepdims=True)

    def zero_frac(D):
        lam = (D * np.exp(ldepth[sub]))[:, None] * floored(D)[types[sub]]
        return float(np.exp(-lam).mean())

    lo, hi = 1.0, 1e7
    for _ in range(60):
        mid = np.sqrt(lo * hi)
        if zero_frac(mid) > 0.92:
            lo = mid
        else:
            hi = mid
    D = float(np.sqrt(lo * hi))
    p = torch.tensor(floored(D), dtype=torch.float32, device=device)
    depth = torch.tensor(D * np.exp(ldepth), dtype=torch.float32, device=device)
    ttypes = torch.tensor(types, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    rows, cols, vals = [], [], []
    block = 4096
    for s in range(0, N, block):
        e = min(N, s + block)
        lam = depth[s:e, None] * p[ttypes[s:e]]
        c = torch.poisson(lam, generator=gen)
        nz = torch.nonzero(c)
        rows.append((nz[:, 0] + s).to(torch.int32).cpu().numpy())
        cols.append(nz[:, 1].to(torch.int32).cpu().numpy())
        vals.append(c[nz[:, 0], nz[:, 1]].cpu().numpy().astype(np.float32))
    X = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, M),
                      dtype=np.float32).tocsc()
    X.sort_indices()
    return X


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if x > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------
# CPU arm: the oracle's device_="cpu" numerics on a bounded sample, scaled to one sclens() pass
# --------------------------------------------------------------------------------------
def cpu_sample(X, n_search, n_perturb, frac=4):
    """Times each distinct stage of the CPU path on a 1/frac (linear) sample of the workload with all
    host cores and scales to the full pass by the stage's complexity and multiplicity."""
    from oracle import sclens_oracle as orc
    N, M = X.shape
    nm = min(N, M)
    ns = max(300, nm // frac)
    cores = os.cpu_count() or 1
    # sample: the first ns cells (N <= M) or genes (N > M) -> a (ns x M) / (N x ns) sub-problem
    Xs = X[:ns, :] if N <= M else X[:, :ns]
    Xs = Xs[:, np.diff(Xs.tocsc().indptr) > 0] if N <= M else Xs[np.diff(Xs.tocsr().indptr) > 0, :]
    Xs = sp.csc_matrix(Xs)
    t0 = time.perf_counter()
    dense, _ = orc.normalize_main(Xs)                                   # :677-696 (Float64, dense)
    t_norm = time.perf_counter() - t0
    t0 = time.perf_counter()
    Y = orc.wishart_matrix(dense, 1 if N <= M else 2, "cpu")            # :345-359 Float64 syrk
    t_gram = time.perf_counter() - t0
    t0 = time.perf_counter()
    L, V = orc.get_eigen(Y, "cpu")                                      # :384 LAPACK, all eigenpairs
    t_eig = time.perf_counter() - t0
    t0 = time.perf_counter()
    _ = np.abs(V.T @ V[:, : V.shape[1] // 2 + 1]).max(axis=0)           # corr_mat + nanmaximum (:742)
    t_corr = time.perf_counter() - t0
    n_s = Y.shape[0]
    lin = (N * M) / (dense.shape[0] * dense.shape[1])
    cube = (nm / n_s) ** 3
    gram_scale = (nm / n_s) ** 2 * (max(N, M) / (dense.shape[1] if N <= M else dense.shape[0]))
    n_full = 2 + 1 + n_search + n_perturb                                # get_sigev x2, Vr2, search, perturbations
    total = n_full * (t_norm * lin + t_gram * gram_scale + t_eig * cube) + n_search * t_corr * cube
    sample = (f"oracle cpu-mode stages timed on a {dense.shape[0]}x{dense.shape[1]} sub-matrix (1/{frac} of the smaller "
              f"side) with {cores} cores: normalise {t_norm:.2f}s, f64 Gram {t_gram:.2f}s, LAPACK eig n={n_s} {t_eig:.2f}s, "
              f"corr {t_corr:.2f}s; scaled by N*M, n^2*K, n^3 and the pass's stage counts "
              f"({n_full} normalise+Gram+eig incl. {n_search} search steps and {n_perturb} perturbations)")
    return total, cores, sample


# --------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed passes (default: 1 at the 68k x 20k workload, 2 otherwise)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # default = the configuration BASELINE.json's metric is quoted on (68k x 20k, fits one GPU): 65 s per pass on one
    # B200, so the default run is 3 warm-ups + 1 timed pass + 1 end-to-end call, about 6 minutes; B is configs[1]
    ap.add_argument("--workload", default=os.environ.get("SCLENS_BENCH_WORKLOAD"), choices=sorted(WORKLOADS))
    ap.add_argument("--n-perturb", type=int, default=20)
    ap.add_argument("--gram-mode", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    workload_note = None
    if args.workload is None:
        # 66 s per pass at 68k x 20k on one B200 (profiles/r1_bench_C_1gpu.json): keep the metric's own shape unless the
        # requested W + K passes (+ end-to-end call + CPU sample) would run past half an hour on ONE GPU - the rule does
        # not look at --gpus, so every point of a scaling series runs the same workload
        est = (args.warmup + (args.steps or 1) + 2) * 66.0
        args.workload = "C" if est <= 1800.0 else "B"
        if args.workload == "B":
            workload_note = (f"{args.warmup}+{args.steps} passes of the 68k x 20k workload were estimated at {est:.0f} s; "
                             "fell back to BASELINE.json configs[1] (10k x 20k) - pass --workload C to force the metric's shape")
    if args.steps is None:
        args.steps = 1 if args.workload == "C" else 2

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    N, M, seed = WORKLOADS[args.workload]
    config = {"workload": f"{args.workload}: synthetic {N} cells x {M} genes, ~92% sparse, n_perturb={args.n_perturb}, "
                          f"full MP/TW fit + sparsity search + stability", "N": N, "M": M,
              "l2": "inputs exceed L2 (dense operand >= 400 MB)", "parallelism": f"replicates+search x{world}"}
    if workload_note:
        config["workload_note"] = workload_note

    if args.impl == "reference":
        if rank != 0:
            return
        X = make_counts_fast(N, M, seed)
        vals = []
        for i in range(args.warmup + args.steps):
            total, cores, sample = cpu_sample(X, NOMINAL_SEARCH_STEPS[args.workload], args.n_perturb)
            if i >= args.warmup:
                vals.append(total)
        sec = float(np.mean(vals))
        v = N / sec
        print(json.dumps({"impl": "reference", "metric": "sclens_cells_per_s", "value": v, "unit": "cells/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "cells/s", "cores": cores, "kind": "port",
                                           "sample": sample + f"; search steps nominal={NOMINAL_SEARCH_STEPS[args.workload]}"},
                          "e2e": {"value": v, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    from sclens_b200 import Handle, sclens
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    t0 = time.perf_counter()
    X = make_counts_fast(N, M, seed, device=dev)
    log(f"[rank {rank}] synthetic counts {X.shape} nnz={X.nnz} sparsity={1 - X.nnz / (N * M):.4f} in {time.perf_counter() - t0:.1f}s")

    h = Handle(device=local_rank, gram_mode=args.gram_mode, seed=seed)

    def fresh_uid():
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            import ctypes as C
            buf = (C.c_uint8 * 128)()
            assert h.lib.scl_nccl_unique_id(buf) == 0
            uid = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
        dist.broadcast(uid, 0)
        return bytes(uid.cpu().tolist())

    if world > 1:
        h.comm_init(fresh_uid(), rank, world)
    h.set_counts(X)

    def one_pass():
        si = h.run_signal()
        ri = h.run_robustness(th=60.0, p_step=0.001, n_perturb=args.n_perturb)
        return si, ri

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(args.warmup):
        t0 = time.perf_counter()
        si, ri = one_pass()
        log(f"[rank {rank}] warmup {i}: {time.perf_counter() - t0:.2f}s n_signal={si.n_signal} n_search={ri.n_search} "
            f"p_sel={ri.p_sel:.3f} n_robust={ri.n_robust}")
    h.reset_profile()
    launches0 = h.profile().kernel_launches
    clocks = ClockSampler(local_rank)
    clocks.start()
    barrier()
    h.timer_start()
    for _ in range(args.steps):
        si, ri = one_pass()
    ms = h.timer_stop()
    barrier()
    clk = clocks.stop()
    prof = h.profile()
    launches = prof.kernel_launches - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    value = N / (ms_per_step / 1e3)

    # ---- e2e through the public API with host buffers (pinned), results read back
    e2e = None
    if args.e2e_steps > 0:
        try:
            pinned = [torch.from_numpy(a).pin_memory() for a in (X.indptr.astype(np.uint32).view(np.int32),
                                                                 X.indices.astype(np.uint32).view(np.int32), X.data)]
            Xp = sp.csc_matrix((pinned[2].numpy(), pinned[1].numpy().view(np.uint32), pinned[0].numpy().view(np.uint32)),
                               shape=X.shape)
            h2d = sum(int(p.numel() * p.element_size()) for p in pinned)
            d2h = 0

            def call(handle):
                out = sclens(Xp, n_perturb=args.n_perturb, gram_mode=args.gram_mode, verbose=False, seed=seed,
                             device=local_rank, handle=handle)
                nbytes = sum(int(np.asarray(v).nbytes) for v in (out["L"], out["L_mp"], out["signal_evec"], out["signal_ev"],
                                                                 out["gene_basis"], out["robustness_scores"]["b_"],
                                                                 out["robustness_scores"]["m_scores"],
                                                                 out["robustness_scores"]["sd_scores"], out["sig_id"]))
                nbytes += sum(int(np.asarray(v).nbytes) for v in out["rec_vals"].values())
                return nbytes, dict(out["info"]["host_ms"])

            # one cold call (N = 1 only): the call creates and destroys its own handle - cuSOLVER state and every device
            # workspace are set up inside it.  Reported beside the headline, not as the headline.
            cold_s, cold_host_ms = None, None
            if world == 1 and args.workload != "C":      # at 68k x 20k the cold call would add a minute to the default run
                barrier()
                t0 = time.perf_counter()
                _, cold_host_ms = call(None)
                barrier()
                cold_s = time.perf_counter() - t0
            # the timed calls reuse the rank's handle (the library context: workspaces, cuSOLVER state, NCCL communicator),
            # as a user running sclens() on one matrix after another would; host buffers in, every result read back
            e2e_s = 0.0
            e2e_host_ms = {}
            for _ in range(args.e2e_steps):
                barrier()
                t0 = time.perf_counter()
                d2h, e2e_host_ms = call(h)
                barrier()
                e2e_s += time.perf_counter() - t0
            te = torch.tensor([e2e_s / args.e2e_steps], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            e2e_s = float(te.item())
            e2e = {"value": N / e2e_s, "unit": "cells/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                   "seconds_per_step": e2e_s, "host_ms_last_call": e2e_host_ms,
                   "handle": "reused across calls (device workspaces, cuSOLVER state and the NCCL communicator persist)",
                   "cold_call_seconds": cold_s, "cold_call_host_ms": cold_host_ms}
        except Exception as exc:      # the device-timed headline above must survive a failure of this leg
            if world > 1:
                raise                  # ranks must stay in step: a one-sided failure cannot be patched over
            e2e = {"value": None, "unit": "cells/s", "error": f"{type(exc).__name__}: {exc}"[:300]}

    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "MEASURED_PEAKS.json (sustained bf16 / copy)" if peaks else "fallback of B200_PROFILING.md"
    # DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed `ncu --set full`
    # captures of the same kernels at this workload's shapes (profiles/r1_traffic.json; null for other workloads)
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json"))).get(args.workload, {})
    except Exception:
        pass
    gram_tf = prof.gram_alg_flops / (prof.gram_gemm_ms * 1e-3) / 1e12 if prof.gram_gemm_ms > 0 else 0.0
    dens_gbs = prof.densify_alg_bytes / (prof.densify_ms * 1e-3) / 1e9 if prof.densify_ms > 0 else 0.0
    roofline = {"kernel": "k_gemm_umma (tcgen05 Gram, syrk schedule)", "bound": "tensor", "achieved": gram_tf,
                "peak": tf_peak, "unit": "TFLOP/s", "frac": gram_tf / tf_peak, "traffic": traffic.get("k_gemm_umma"),
                "peak_source": peak_src, "launches": int(prof.gram_gemm_launches),
                "avg_ms_per_launch": prof.gram_gemm_ms / max(1, prof.gram_gemm_launches),
                "alg_flops_per_launch": prof.gram_alg_flops / max(1, prof.gram_gemm_launches),
                "hbm_kernel": {"kernel": "k_densify_tma2 (normalised dense binary16 operand: background + sparse patches composed in shared memory, TMA bulk stores)",
                               "bound": "hbm", "achieved": dens_gbs, "peak": hbm_peak, "unit": "GB/s",
                               "frac": dens_gbs / hbm_peak, "launches": int(prof.densify_launches),
                               "alg_bytes_per_launch": prof.densify_alg_bytes / max(1, prof.densify_launches),
                               "traffic": traffic.get("k_densify")}}
    stage_ms = {"gram_gemm": prof.gram_gemm_ms, "other_gemm": prof.other_gemm_ms, "densify": prof.densify_ms,
                "stats": prof.stats_ms, "sparse": prof.sparse_ms, "syevd_library": prof.syevd_ms,
                "syevd_calls": int(prof.syevd_calls), "refine_f64": prof.refine_ms, "small": prof.small_ms, "total": ms}
    line = {"metric": "sclens_cells_per_s", "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f16 operands / f32 accumulate (Gram), f32 syevd, f64 statistics",
            "data": "synthetic", "config": config, "clocks": clk, "gpu_launches": int(launches),
            "roofline": roofline, "stage_ms_over_timed_region": stage_ms,
            "phase_ms_last_step": {"signal_normalise_gram": si.t_gram_ms, "signal_syevd": si.t_syevd_ms, "null_matrix": si.t_null_ms,
                                   "mp_fit_host": si.t_fit_ms, "backproject": si.t_backproject_ms,
                                   "noise_baseline_and_zero_candidates": ri.t_baseline_ms, "search_total": ri.t_search_ms,
                                   "search_syevd": ri.t_search_syevd_ms, "perturbations": ri.t_perturb_ms,
                                   "scores": ri.t_score_ms, "gene_basis": ri.t_outputs_ms},
            "gemm_tflops": {"other_gemm_2mnk": (prof.other_gemm_flops / (prof.other_gemm_ms * 1e-3) / 1e12) if prof.other_gemm_ms > 0 else 0.0},
            "result": {"n_signal": si.n_signal, "lambda_c": si.lambda_c, "n_search": ri.n_search, "p_sel": ri.p_sel,
                       "n_robust": ri.n_robust}}
    if e2e is not None:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        try:
            total, cores, sample = cpu_sample(X, ri.n_search, args.n_perturb)
            line["cpu_baseline"] = {"value": N / total, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample,
                                    "seconds_per_pass_estimate": total}
        except Exception as exc:          # e.g. host memory: report it, keep the line
            line["cpu_baseline"] = {"value": None, "unit": "cells/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {type(exc).__name__}: {exc}"[:300]}
    print(json.dumps(line))
    h.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
