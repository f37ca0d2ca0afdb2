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

// LAPACK slarfg on x[0..L) (shared memory), by warp 0: v (v[0] = 1, zero-padded to kBand), tau, beta
__device__ __forceinline__ void house_warp(const float* x, int L, float* v, float* tau, float* beta) {
  const int lane = (int)threadIdx.x & 31;
  const float x0 = lane < L ? x[lane] : 0.f, x1 = lane + 32 < L ? x[lane + 32] : 0.f;
  const float alpha = __shfl_sync(0xffffffffu, x0, 0);
  float ss = (lane == 0 ? 0.f : x0 * x0) + x1 * x1;
  ss = warp_sum(ss);
  float t = 0.f, b = alpha, sc = 0.f;
  if (ss > 0.f) {
    b = -copysignf(sqrtf(alpha * alpha + ss), alpha);
    t = (b - alpha) / b;
    sc = 1.f / (alpha - b);
  }
  v[lane] = lane == 0 ? 1.f : x0 * sc;
  v[lane + 32] = x1 * sc;
  if (lane == 0) {
    *tau = t;
    *beta = b;
  }
}

__global__ void __launch_bounds__(256, 1) k_sb2st(SbArgs a) {
  __shared__ __align__(16) float Bt[B][LD], Dt[B][LD];   // [column][row]
  __shared__ float vbuf[2][B], xs[B], ws[B];
  __shared__ float s_tau[2], s_beta;
  const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = a.n;
  int seen = 0;   // thread 0: last progress value read from the predecessor
  unsigned long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = now_ns();

  auto wait_for = [&](int s, int need) {
    if (s > 0 && tid == 0) {
      while (seen < need) seen = ld_acquire(a.prog + s - 1);
    }
    __syncthreads();
  };
  auto publish = [&](int s, int val) {
    __threadfence();
    __syncthreads();
    if (tid == 0) st_release(a.prog + s, val);
  };
  // diagonal block rows/cols r..r+L-1 -> Dt, full symmetric
  auto load_diag = [&](int r, int L) {
    const int ii = tid & 63;
#pragma unroll 4
    for (int c = tid >> 6; c < L; c += 4) {
      const int i = c + ii;
      if (i < L) {
        const float v = __ldcg(a.AB + (size_t)(r + c) * kLdab + ii);
        Dt[c][i] = v;
        Dt[i][c] = v;
      }
    }
  };
  // D <- H D H with H = I - tau v v', written straight back to the band (lower triangle)
  auto two_sided = [&](int r, int L, const float* v, float tau) {
    {  // w = tau D v : thread (i = tid / 4, q = tid % 4) sums c = q, q + 4, ...
      const int i = tid >> 2, q = tid & 3;
      float s = 0.f;
      if (i < L)
        for (int c = q; c < L; c += 4) s = fmaf(Dt[c][i], v[c], s);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (q == 0) ws[i] = i < L ? tau * s : 0.f;
    }
    __syncthreads();
    // alpha = w'v, every warp for itself
    float al = ws[lane] * v[lane] + ws[lane + 32] * v[lane + 32];
    al = warp_sum(al);
    const float h = 0.5f * tau * al;
    const int ii = tid & 63;
#pragma unroll 4
    for (int c = tid >> 6; c < L; c += 4) {
      const int i = c + ii;
      if (i < L) {
        const float wi = ws[i] - h * v[i], wc = ws[c] - h * v[c];
        __stcg(a.AB + (size_t)(r + c) * kLdab + ii, Dt[c][i] - v[i] * wc - wi * v[c]);
      }
    }
  };
  auto keep_reflector = [&](int s, int k, const float* v, float tau) {
    if (a.keep) {
      if (tid < B) a.V2[(size_t)s * a.ldv2 + (size_t)k * B + tid] = v[tid];
      if (tid == 0) a.tau2[(size_t)s * a.ldt2 + k] = tau;
    }
  };

  for (int s = (int)blockIdx.x; s < n - 2; s += (int)gridDim.x) {
    seen = 0;
    int cur = 0;
    int r0 = s + 1, L = min(B, n - r0);
    // ---- step 0: the reflector of column s, applied to its diagonal block from both sides
    wait_for(s, 2);
    if (tid < B) {
      xs[tid] = tid < L ? __ldcg(a.AB + (size_t)s * kLdab + 1 + tid) : 0.f;
      vbuf[0][tid] = 0.f;
      vbuf[1][tid] = 0.f;
    }
    load_diag(r0, L);
    __syncthreads();
    if (warp == 0) house_warp(xs, L, vbuf[cur], &s_tau[cur], &s_beta);
    __syncthreads();
    if (tid < L) __stcg(a.AB + (size_t)s * kLdab + 1 + tid, tid == 0 ? s_beta : 0.f);
    two_sided(r0, L, vbuf[cur], s_tau[cur]);
    keep_reflector(s, 0, vbuf[cur], s_tau[cur]);
    publish(s, 1);
    for (int k = 1;; ++k) {
      const int r1 = r0 + L, L1 = min(B, n - r1);
      if (L1 < 1) break;
      const float* v = vbuf[cur];
      float* v1 = vbuf[cur ^ 1];
      const float tau = s_tau[cur];
      SB_TICK(0);
      wait_for(s, k + 2);
      SB_TICK(1);
      // ---- the block below: rows r1.., columns r0..r0+L-1 (band offset of (i, c): L + i - c)
      {
        const int i = tid & 63;
#pragma unroll 4
        for (int c = tid >> 6; c < L; c += 4)
          Bt[c][i] = i < L1 ? __ldcg(a.AB + (size_t)(r0 + c) * kLdab + (L - c) + i) : 0.f;
      }
      if (L1 >= 2) load_diag(r1, L1);
      __syncthreads();
      SB_TICK(2);
      {  // x = B v, then B -= tau x v'
        const int i = tid >> 2, q = tid & 3;
        float s1 = 0.f;
        for (int c = q; c < L; c += 4) s1 = fmaf(Bt[c][i], v[c], s1);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        if (q == 0) xs[i] = tau * s1;
      }
      __syncthreads();
      {
        const int i = tid & 63;
        const float xi = xs[i];
#pragma unroll 4
        for (int c = tid >> 6; c < L; c += 4) Bt[c][i] = fmaf(-xi, v[c], Bt[c][i]);
      }
      __syncthreads();
      SB_TICK(3);
      if (L1 >= 2) {
        // new reflector from the block's first column
        if (warp == 0) house_warp(Bt[0], L1, v1, &s_tau[cur ^ 1], &s_beta);
        __syncthreads();
        const float tau1 = s_tau[cur ^ 1];
        {  // columns 1.. : B -= tau1 v1 (v1' B); the four lanes of a column own all of its rows
          const int c = tid >> 2, q = tid & 3;
          const bool act = c >= 1 && c < L;
          float y = 0.f;
          if (act)
            for (int i = q; i < L1; i += 4) y = fmaf(v1[i], Bt[c][i], y);
          y += __shfl_xor_sync(0xffffffffu, y, 1);   // converged: every lane of the warp takes part
          y += __shfl_xor_sync(0xffffffffu, y, 2);
          y *= tau1;
          if (act) {
            for (int i = q; i < L1; i += 4) Bt[c][i] = fmaf(-y, v1[i], Bt[c][i]);
          } else if (c == 0) {
            for (int i = q; i < L1; i += 4) Bt[0][i] = i == 0 ? s_beta : 0.f;
          }
        }
        __syncthreads();
      }
      {
        const int i = tid & 63;
        if (i < L1) {
#pragma unroll 4
          for (int c = tid >> 6; c < L; c += 4) __stcg(a.AB + (size_t)(r0 + c) * kLdab + (L - c) + i, Bt[c][i]);
        }
      }
      if (L1 < 2) break;
      SB_TICK(4);
      two_sided(r1, L1, v1, s_tau[cur ^ 1]);
      keep_reflector(s, k, v1, s_tau[cur ^ 1]);
      SB_TICK(5);
      publish(s, k + 1);
      SB_TICK(6);
      r0 = r1;
      L = L1;
      cur ^= 1;
    }
    publish(s, kDone);
  }
  if (a.prof && tid == 0)
    for (int q = 0; q < 8; ++q) a.prof[(size_t)blockIdx.x * 8 + q] = tacc[q];
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
      static const char* name[7] = {"(step 0 / loop overhead)", "wait for predecessor", "load blocks", "right apply", "reflector + left apply + store",
                                    "two-sided diagonal block", "publish (fence + flag)"};
      for (int q = 0; q < 7; ++q) {
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
