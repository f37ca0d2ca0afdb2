// Stage 2 of the two-stage eigensolver: symmetric band (half bandwidth kBand) -> tridiagonal by bulge chasing.
//
// Sweep s annihilates column s below the sub-diagonal with one Householder reflector and chases the bulge it creates down
// the band, one kBand x kBand block pair per step (the block below the previous reflector's diagonal block takes the
// reflector from the right, a new reflector clears its first column, the next diagonal block takes that one from both sides).
// Step k of sweep s touches rows/columns s + 1 + (k-1) kBand .. s + (k+1) kBand only, so sweep s may run step k as soon as
// sweep s-1 has finished step k+1: the sweeps form a pipeline, one CTA per sweep in flight, each publishing its progress in
// a flag that its successor polls (ld.acquire / st.release at gpu scope, band data through L2 only).  One persistent
// cooperative kernel (co-residency is what makes the spin-waits safe); the whole band (n x 2 kBand floats, 10 MB at
// n = 20 000) lives in L2.  The critical path is 2n dependent steps, which is why a step is kept to a handful of barriers.
#include <algorithm>
#include <cstdlib>
#include <vector>
#include "common.cuh"
#include "tmp.cuh"
#include "twostage.h"

namespace scl {
namespace {

constexpr int B = kBand;
constexpr int LD = 72;         // shared-memory stride: conflict-free for the column-of-rows and the row-of-columns passes
constexpr int kDone = 1 << 30;

struct SbArgs {
  float* AB;
  int n;
  int* prog;
  float* V2;
  long long ldv2;
  float* tau2;
  long long ldt2;
  int keep;
  unsigned long long* prof;   // [grid][8] nanoseconds of thread 0 per section (SCL_TRACE), may be null
};

__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define SB_TICK(slot)                         \
  do {                                        \
    if (a.prof && tid == 0) {                 \
      const unsigned long long t_ = now_ns(); \
      tacc[slot] += t_ - tlast;               \
      tlast = t_;                             \
    }                                         \
  } while (0)

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// LAPACK slarfg on x[0..64) (shared memory, zero beyond the reflector's length), by every warp for itself: the reflector is
// v[0] = 1, v[i] = x[i] * sc; returns tau, beta, sc in every lane
__device__ __forceinline__ void house_all(const float* x, float& tau, float& beta, float& sc) {
  const int lane = (int)threadIdx.x & 31;
  const float x0 = x[lane], x1 = x[lane + 32];
  const float alpha = __shfl_sync(0xffffffffu, x0, 0);
  float ss = (lane == 0 ? 0.f : x0 * x0) + x1 * x1;
  ss = warp_sum(ss);
  tau = 0.f; beta = alpha; sc = 0.f;
  if (ss > 0.f) {
    beta = -copysignf(sqrtf(alpha * alpha + ss), alpha);
    tau = (beta - alpha) / beta;
    sc = 1.f / (alpha - beta);
  }
}

// Thread (ri = tid / 4, q = tid % 4) keeps its part of a 64 x 64 block in registers: for the passes that sum along a row it
// owns row ri, columns q + 4j (j < 16); for the pass that sums along a column it owns column ri, rows q + 4j - the four
// partial sums of a line always sit in adjacent lanes (two shuffles), and the block changes ownership once through shared
// memory.  Loops have fixed trip counts (ragged blocks at the end of the band are zero-filled), so everything unrolls.
__global__ void __launch_bounds__(256, 1) k_sb2st(SbArgs a) {
  __shared__ __align__(16) float Bt[B][LD];   // [column][row]; Bt[0] is the column the new reflector is made from
  __shared__ float ws[B];
  const int tid = (int)threadIdx.x, lane = tid & 31;
  const int ri = tid >> 2, q = tid & 3;
  const int n = a.n;
  int seen = 0;   // thread 0: last progress value read from the predecessor
  unsigned long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = now_ns();

  auto wait_for = [&](int s, int need) {
    if (s > 0 && tid == 0) {
      while (seen < need) seen = ld_acquire(a.prog + s - 1);
    }
    __syncthreads();
  };
  // the barrier orders every thread's band stores before thread 0's release store (cumulativity)
  auto publish = [&](int s, int val) {
    __syncthreads();
    if (tid == 0) st_release(a.prog + s, val);
  };
  // diagonal block rows/cols r..r+Ld-1 from its lower triangle: dq[j] = D[q + 4j][ri]
  auto load_diag = [&](float (&dq)[16], int r, int Ld) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int c = q + 4 * j;
      const int lo = min(ri, c), hi = max(ri, c);
      dq[j] = hi < Ld ? __ldcg(a.AB + (size_t)(r + lo) * kLdab + (hi - lo)) : 0.f;
    }
  };
  // D <- H D H, H = I - tau v v' with v[x] = x == 0 ? 1 : Bt[0][x] * sc; lower triangle written straight back to the band
  auto two_sided = [&](const float (&dq)[16], const float (&vc)[16], float vi, float tau, float sc, int r, int Ld) {
    float w = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) w = fmaf(dq[j], vc[j], w);
    w += __shfl_xor_sync(0xffffffffu, w, 1);
    w += __shfl_xor_sync(0xffffffffu, w, 2);
    if (q == 0) ws[ri] = tau * w;
    __syncthreads();
    const float vl0 = lane == 0 ? 1.f : Bt[0][lane] * sc, vl1 = Bt[0][lane + 32] * sc;
    float al = ws[lane] * vl0 + ws[lane + 32] * vl1;
    al = warp_sum(al);
    const float h = 0.5f * tau * al;
    const float wi = ws[ri] - h * vi;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int c = q + 4 * j;
      const float wc = ws[c] - h * vc[j];
      if (ri >= c && ri < Ld) __stcg(a.AB + (size_t)(r + c) * kLdab + (ri - c), dq[j] - vi * wc - wi * vc[j]);
    }
  };
  auto keep_reflector = [&](int s, int k, float tau, float sc) {
    if (a.keep) {
      if (tid < B) a.V2[(size_t)s * a.ldv2 + (size_t)k * B + tid] = tid == 0 ? 1.f : Bt[0][tid] * sc;
      if (tid == 0) a.tau2[(size_t)s * a.ldt2 + k] = tau;
    }
  };

  for (int s = (int)blockIdx.x; s < n - 2; s += (int)gridDim.x) {
    seen = 0;
    int r0 = s + 1, L = min(B, n - r0);
    float vq[16], dq[16];
    float tau, beta, sc;
    // ---- step 0: the reflector of column s, applied to its diagonal block from both sides
    wait_for(s, 2);
    if (tid < B) Bt[0][tid] = tid < L ? __ldcg(a.AB + (size_t)s * kLdab + 1 + tid) : 0.f;
    load_diag(dq, r0, L);
    __syncthreads();
    house_all(Bt[0], tau, beta, sc);
    if (tid < L) __stcg(a.AB + (size_t)s * kLdab + 1 + tid, tid == 0 ? beta : 0.f);
#pragma unroll
    for (int j = 0; j < 16; ++j) vq[j] = (q + 4 * j) == 0 ? 1.f : Bt[0][q + 4 * j] * sc;
    two_sided(dq, vq, ri == 0 ? 1.f : Bt[0][ri] * sc, tau, sc, r0, L);
    keep_reflector(s, 0, tau, sc);
    publish(s, 1);
    for (int k = 1;; ++k) {
      const int r1 = r0 + L, L1 = min(B, n - r1);
      if (L1 < 1) break;
      SB_TICK(0);
      wait_for(s, k + 2);
      SB_TICK(1);
      // ---- the block below: rows r1.., columns r0..r0+L-1 (band offset of (row i, column c): L + i - c); bq[j] = B[q + 4j][ri]
      float bq[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int c = q + 4 * j;
        bq[j] = (c < L && ri < L1) ? __ldcg(a.AB + (size_t)(r0 + c) * kLdab + (L - c) + ri) : 0.f;
      }
      if (L1 >= 2) load_diag(dq, r1, L1);
      {  // B <- B (I - tau v v')
        float x = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) x = fmaf(bq[j], vq[j], x);
        x += __shfl_xor_sync(0xffffffffu, x, 1);
        x += __shfl_xor_sync(0xffffffffu, x, 2);
        x *= tau;
#pragma unroll
        for (int j = 0; j < 16; ++j) Bt[q + 4 * j][ri] = fmaf(-x, vq[j], bq[j]);
      }
      __syncthreads();
      SB_TICK(2);
      // ---- ownership by columns: thread owns column ri, rows q + 4j
      float bc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) bc[j] = Bt[ri][q + 4 * j];
      float tau1 = 0.f, sc1 = 0.f;
      if (L1 >= 2) {
        house_all(Bt[0], tau1, beta, sc1);
        float vr[16], y = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          vr[j] = (q + 4 * j) == 0 ? 1.f : Bt[0][q + 4 * j] * sc1;
          y = fmaf(vr[j], bc[j], y);
        }
        y += __shfl_xor_sync(0xffffffffu, y, 1);
        y += __shfl_xor_sync(0xffffffffu, y, 2);
        y *= tau1;
#pragma unroll
        for (int j = 0; j < 16; ++j) bc[j] = ri == 0 ? ((q + 4 * j) == 0 ? beta : 0.f) : fmaf(-y, vr[j], bc[j]);
      }
      if (ri < L) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (q + 4 * j < L1) __stcg(a.AB + (size_t)(r0 + ri) * kLdab + (L - ri) + q + 4 * j, bc[j]);
      }
      SB_TICK(3);
      if (L1 < 2) break;
#pragma unroll
      for (int j = 0; j < 16; ++j) vq[j] = (q + 4 * j) == 0 ? 1.f : Bt[0][q + 4 * j] * sc1;
      two_sided(dq, vq, ri == 0 ? 1.f : Bt[0][ri] * sc1, tau1, sc1, r1, L1);
      keep_reflector(s, k, tau1, sc1);
      SB_TICK(4);
      publish(s, k + 1);
      SB_TICK(5);
      r0 = r1;
      L = L1;
      tau = tau1;
    }
    publish(s, kDone);
  }
  if (a.prof && tid == 0)
    for (int t = 0; t < 8; ++t) a.prof[(size_t)blockIdx.x * 8 + t] = tacc[t];
}

__global__ void k_band_to_tridiag(const float* __restrict__ AB, int n, float* __restrict__ d, float* __restrict__ e) {
  const int j = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (j < n) {
    d[j] = AB[(size_t)j * kLdab];
    if (j + 1 < n) e[j] = AB[(size_t)j * kLdab + 1];
  }
}

}  // namespace

void sb2st(float* AB, int n, float* d, float* e, bool keep, float* V2, long long ldv2, float* tau2, long long ldt2,
           cudaStream_t st) {
  Tmp<int> prog((size_t)n + 1, st);
  SCL_CUDA(cudaMemsetAsync(prog.p, 0, ((size_t)n + 1) * sizeof(int), st));
  if (n > 2) {
    int dev = 0, sms = 0, per_sm = 0;
    SCL_CUDA(cudaGetDevice(&dev));
    SCL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SCL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sb2st, 256, 0));
    SCL_REQUIRE(per_sm >= 1, "sb2st: kernel does not fit on an SM");
    // at most n / (2 kBand) sweeps can be in flight at once (each trails its predecessor by two steps)
    const int grid = std::max(1, std::min(sms, n / (2 * B) + 2));
    static const bool trace = getenv("SCL_TRACE") != nullptr;
    Tmp<unsigned long long> prof(trace ? (size_t)grid * 8 : 1, st);
    SbArgs args{AB, n, prog.p, V2, ldv2, tau2, ldt2, keep ? 1 : 0, trace ? prof.p : nullptr};
    void* params[] = {&args};
    SCL_CUDA(cudaLaunchCooperativeKernel((void*)k_sb2st, dim3(grid), dim3(256), params, 0, st));
    if (trace) {
      std::vector<unsigned long long> hp((size_t)grid * 8);
      SCL_CUDA(cudaMemcpyAsync(hp.data(), prof.p, hp.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
      SCL_CUDA(cudaStreamSynchronize(st));
      static const char* name[6] = {"(step 0 / next sweep's start)", "wait for predecessor", "loads + right apply", "reflector + left apply + store",
                                    "two-sided diagonal block", "publish (barrier + flag)"};
      for (int q = 0; q < 6; ++q) {
        double mn = 1e30, mx = 0, av = 0;
        for (int g = 0; g < grid; ++g) {
          const double v = (double)hp[(size_t)g * 8 + q] * 1e-6;
          mn = std::min(mn, v); mx = std::max(mx, v); av += v / grid;
        }
        fprintf(stderr, "[scl] sb2st n=%d grid=%d %-32s ms per CTA: min %8.2f  mean %8.2f  max %8.2f\n", n, grid, name[q], mn, av, mx);
      }
    }
  }
  k_band_to_tridiag<<<(n + 255) / 256, 256, 0, st>>>(AB, n, d, e);
  SCL_CUDA(cudaGetLastError());
  count_launches(2);
}

}  // namespace scl
