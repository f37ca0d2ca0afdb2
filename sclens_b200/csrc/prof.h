// Per-handle kernel-class timing with CUDA events on the launching stream.  Events are
// recorded asynchronously and resolved at the existing stage-end synchronisation points, so
// profiling adds no host-device synchronisation to the path.
#pragma once
#include <atomic>
#include <vector>
#include <cuda_runtime.h>

namespace scl {

enum ProfKind { PK_GRAM_GEMM = 0, PK_OTHER_GEMM, PK_DENSIFY, PK_STATS, PK_SPARSE, PK_SYEVD, PK_SMALL, PK_REFINE, PK_COMM, PK_COUNT };

struct ProfEvent {
  cudaEvent_t a, b;
  int kind;
};

struct Prof {
  double ms[PK_COUNT] = {};
  long calls[PK_COUNT] = {};
  double gram_alg_flops = 0;      // n(n+1)K per Gram launch
  double other_gemm_flops = 0;    // 2 m n k
  double densify_alg_bytes = 0;   // 8 nnz + 4(M+1) + N*M*s_out
  double sparse_alg_bytes = 0;    // 20 nnz (+12 n_add)
  double stats_alg_bytes = 0;     // 40 nnz per normalisation
  double comm_bytes = 0;          // payload of the NCCL collectives
  long stats_norms = 0;           // normalisations (one set of statistics passes each)
  std::vector<ProfEvent> pending;
  void resolve();
  void reset();
};

extern std::atomic<long long> g_kernel_launches;
inline void count_launches(int n) { g_kernel_launches.fetch_add(n, std::memory_order_relaxed); }

struct ProfScope {
  Prof* p;
  cudaStream_t st;
  ProfEvent ev;
  ProfScope(Prof* prof, cudaStream_t s, int kind) : p(prof), st(s) {
    ev.kind = kind;
    cudaEventCreate(&ev.a);
    cudaEventCreate(&ev.b);
    cudaEventRecord(ev.a, st);
  }
  ~ProfScope() {
    cudaEventRecord(ev.b, st);
    p->pending.push_back(ev);
  }
};

}  // namespace scl
