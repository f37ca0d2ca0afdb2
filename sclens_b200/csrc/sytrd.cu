// Householder tridiagonalisation of a symmetric FP32 matrix (the first and by far the longest part of the full-spectrum
// solve, _get_eigen src/scLENS.jl:375-382): one persistent cooperative kernel, output in LAPACK ssytrd('L') form so that
// the library's back-transformation (Sormtr) applies unchanged.
//
// Why not the library's: cusolverDnSsytrd is a one-stage blocked reduction whose matrix-vector half streams the whole
// trailing SQUARE per column (4 n^3 / 3 bytes: 1.6 s at n = 20 000 at HBM speed, 2.1-2.4 s measured).  Here
//   * the symmetric matrix-vector product reads the LOWER TRIANGLE only - every tile contributes to y twice, once through
//     its rows and once through its columns - half the bytes;
//   * one launch: the ~60 000 dependent steps of a 20 000-column reduction are separated by a hand-rolled grid barrier
//     (one atomic + an acquire spin, ~2 us) instead of kernel boundaries;
//   * the panel's corrections are applied on the fly (W is kept without its last rank-1 term, alpha_c v_c, which needs a
//     global dot product - the term is added wherever W is read), so a column costs three barriers, not five;
//   * results are deterministic: partial sums meet in a fixed order (per-CTA row partials in shared memory, reduced by row
//     owners), no floating-point atomics.
// Blocked algorithm = LAPACK slatrd / ssytrd (lower): panel of kNB columns, rank-2k update of the trailing matrix
// A22 -= V W' + W V' (register-tiled FP32 FMA, lower triangle only) after each panel.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>
#include "common.cuh"
#include "eigen.h"
#include "tmp.cuh"

namespace scl {
namespace {

constexpr int kNB = 32;        // panel width = strip width of the matrix-vector product
constexpr int kT = 256;        // threads per CTA (one CTA per SM; 255 registers per thread for the 128 x 32 tiles held in flight)
constexpr int kW = kT / 32;
constexpr int kTile = 128;     // rank-2k update tile (8 x 8 outputs per thread)

struct SyArgs {
  float* A;          // n x n, column-major with leading dimension ld (a multiple of 4): lower triangle in, reflectors out
  int n, ld;
  float* W;          // ld x kNB, column-major: the panel's W without its alpha_c v_c terms
  float *d, *e, *tau;
  float* P;          // [grid][ld] row partials of the matrix-vector product
  float* ycol;       // n: column parts of the matrix-vector product
  double* part;      // [grid][2 kNB + 2]: panel dot products, norm / w'v partials
  float* scal;       // [0] = A(i+1, i) after the column update
  unsigned* bar;
  unsigned long long* prof;   // [grid][16] nanoseconds per section of CTA's thread 0 (timing studies; may be null)
};

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
  target += gridDim.x;
}

__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define SY_TICK(slot)                                   \
  do {                                                  \
    if (tid == 0) {                                     \
      const unsigned long long t_ = now_ns();           \
      tacc[slot] += t_ - tlast;                         \
      tlast = t_;                                       \
    }                                                   \
  } while (0)

__device__ __forceinline__ double warp_sum_d(double v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sum over the CTA (all threads call); result valid in every thread
__device__ double cta_sum(double v, double* red) {
  v = warp_sum_d(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0;
  for (int w = 0; w < kW; ++w) s += red[w];
  return s;
}

// sum of `cnt` doubles at stride `stride`, by one warp, fixed order (lane-strided partial sums, butterfly): every lane gets it
__device__ __forceinline__ double warp_gather_sum(const double* p, int cnt, size_t stride) {
  const int lane = threadIdx.x & 31;
  double s = 0;
  for (int q = lane; q < cnt; q += 32) s += __ldcg(p + (size_t)q * stride);
  return warp_sum_d(s);
}

// panel loops: all kNB column slots are issued (clamped address, value masked) so the loads of a row leave together
__device__ __forceinline__ void load_panel_row(const SyArgs& a, size_t ld, int k0, int c, int r, const float* s_alpha,
                                               float (&vr)[kNB], float (&wr)[kNB]) {
#pragma unroll
  for (int cp = 0; cp < kNB; ++cp) {
    const bool ok = cp < c;
    const float v = __ldcg(a.A + (ok ? (size_t)(k0 + cp) * ld + r : 0));
    const float w = __ldcg(a.W + (ok ? (size_t)cp * ld + r : 0));
    vr[cp] = ok ? v : 0.f;
    wr[cp] = ok ? fmaf(s_alpha[cp], v, w) : 0.f;
  }
}

__global__ void __launch_bounds__(kT, 1) k_sytrd(SyArgs a) {
  extern __shared__ __align__(16) float smem[];   // row accumulator of the matrix-vector product / update tiles
  __shared__ double s_red[kW];
  __shared__ float s_rowV[kNB], s_rowW[kNB], s_p[kNB], s_q[kNB], s_alpha[kNB];
  __shared__ float s_col[kW][32];
  __shared__ double s_dd[kT / (2 * kNB)][2 * kNB];
  __shared__ float s_pq[kW][2 * kNB];
  __shared__ double s_scal[3];   // tau, scale, -
  float* ysm = smem;
  const int n = a.n, P = (int)gridDim.x, b = (int)blockIdx.x, tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t ld = (size_t)a.ld;
  unsigned target = gridDim.x;
  unsigned long long tacc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tlast = now_ns();
  double* part1 = a.part + (size_t)P * 2 * kNB;   // [P] squared norms
  double* part3 = part1 + P;                      // [P] w'v

  for (int k0 = 0; k0 < n; k0 += kNB) {
    const int k1 = min(n, k0 + kNB), ncols = k1 - k0;
    for (int c = 0; c < ncols; ++c) {
      const int i = k0 + c, t0 = i + 1;
      // ---- alpha of the previous column of this panel; row i of V and of W (with its alpha terms)
      if (c > 0) {
        if (warp == 0) {
          const double s = warp_gather_sum(part3, P, 1);
          if (lane == 0) s_alpha[c - 1] = (float)(-0.5 * s_scal[0] * s);
        }
        __syncthreads();
        if (tid < c) {
          const float v = __ldcg(a.A + (size_t)(k0 + tid) * ld + i);
          s_rowV[tid] = v;
          s_rowW[tid] = __ldcg(a.W + (size_t)tid * ld + i) + s_alpha[tid] * v;
        }
        __syncthreads();
      }
      // ---- phase 1: A(i:n, i) -= V(i:n, :) Wf(i, :)' + Wf(i:n, :) V(i, :)'; diagonal, sub-diagonal, |x|^2
      double nrm = 0;
      for (int rb = b + P * warp; rb * 32 < n; rb += P * kW) {   // row block rb belongs to CTA rb % P, warp (rb / P) % kW
        const int r = rb * 32 + lane;
        if (r < i || r >= n) continue;
        float upd = 0.f;
        const float a0 = __ldcg(a.A + (size_t)i * ld + r);
        {
          float vr[kNB], wr[kNB];
          load_panel_row(a, ld, k0, c, r, s_alpha, vr, wr);
#pragma unroll
          for (int cp = 0; cp < kNB; ++cp) {
            if (cp < c) {
              upd = fmaf(vr[cp], s_rowW[cp], upd);
              upd = fmaf(wr[cp], s_rowV[cp], upd);
            }
          }
        }
        const float ar = a0 - upd;
        __stcg(a.A + (size_t)i * ld + r, ar);
        if (r == i) a.d[i] = ar;
        else if (r == t0) __stcg(a.scal, ar);
        else nrm += (double)ar * (double)ar;
      }
      if (i == n - 1) break;   // last diagonal entry: nothing left to reduce
      nrm = cta_sum(nrm, s_red);
      if (tid == 0) __stcg(part1 + b, nrm);
      SY_TICK(0);
      grid_barrier(a.bar, target);
      SY_TICK(1);

      // ---- phase 2: reflector scalars; y = A22 v over the lower triangle; panel dot products
      if (warp == 0) {
        const double xn2 = warp_gather_sum(part1, P, 1);
        const double ain = (double)__ldcg(a.scal);
        double beta = ain, tau = 0, scale = 0;
        if (xn2 > 0) {
          beta = -copysign(sqrt(ain * ain + xn2), ain);
          tau = (beta - ain) / beta;
          scale = 1.0 / (ain - beta);
        }
        if (lane == 0) {
          s_scal[0] = tau;
          s_scal[1] = scale;
          if (b == 0) { a.e[i] = (float)beta; a.tau[i] = (float)tau; }
        }
      }
      for (int r = (t0 & ~3) + tid; r < n; r += kT) ysm[r] = 0.f;
      __syncthreads();
      const float tau = (float)s_scal[0], scale = (float)s_scal[1];
      const float* coli = a.A + (size_t)i * ld;
      {
        // panel dots p = V' v, q = Wf' v over the rows this warp owns
        float pacc[kNB], qacc[kNB];
#pragma unroll
        for (int cp = 0; cp < kNB; ++cp) { pacc[cp] = 0.f; qacc[cp] = 0.f; }
        bool any = false;
        if (c > 0) {
          for (int rb = b + P * warp; rb * 32 < n; rb += P * kW) {
            const int r = rb * 32 + lane;
            if (r < t0 || r >= n) continue;
            any = true;
            const float vr_i = r == t0 ? 1.f : __ldcg(coli + r) * scale;
            float vr[kNB], wr[kNB];
            load_panel_row(a, ld, k0, c, r, s_alpha, vr, wr);
#pragma unroll
            for (int cp = 0; cp < kNB; ++cp) {
              pacc[cp] = fmaf(vr[cp], vr_i, pacc[cp]);
              qacc[cp] = fmaf(wr[cp], vr_i, qacc[cp]);
            }
          }
        }
        if (__any_sync(0xffffffffu, any)) {
#pragma unroll
          for (int cp = 0; cp < kNB; ++cp) {
            const float ps = warp_sum_f(pacc[cp]), qs = warp_sum_f(qacc[cp]);
            if (lane == 0) { s_pq[warp][cp] = ps; s_pq[warp][kNB + cp] = qs; }
          }
        } else {
          s_pq[warp][lane] = 0.f;
          s_pq[warp][32 + lane] = 0.f;
        }
        __syncthreads();
        if (tid < 2 * kNB) {
          double s = 0;
          for (int w = 0; w < kW; ++w) s += (double)s_pq[w][tid];
          __stcg(a.part + (size_t)b * 2 * kNB + tid, s);
        }
      }
      SY_TICK(2);
      // strips of 32 columns, dealt to the CTAs cyclically; a warp walks 32-row blocks down its strip
      {
        // strips are dealt in snake order (round 0: CTA b takes the b-th tallest, round 1: the (P-1-b)-th of the next P, ...):
        // strip heights fall linearly, so a plain cyclic deal would give the first CTAs 60 % more rows than the last
        const int s0 = t0 >> 5, smax = (n - 1) >> 5;
        for (int round = 0;; ++round) {
          const int s = s0 + round * P + ((round & 1) ? P - 1 - b : b);
          if (s0 + round * P > smax) break;   // uniform over the grid
          if (s > smax) continue;
          const int ja = max(32 * s, t0), jb = min(32 * s + 32, n);
          const int jc = 32 * s + lane;
          const float vC = (jc >= ja && jc < jb) ? (jc == t0 ? 1.f : __ldcg(coli + jc) * scale) : 0.f;
          float colacc[32];
#pragma unroll
          for (int cc = 0; cc < 32; ++cc) colacc[cc] = 0.f;
          // the strip's 32 x 32 diagonal block (lower triangle, one row per lane) goes to the last warp; below it the warps
          // walk 128-row blocks aligned to the strip, four consecutive rows per lane: one 16-byte load per column
          const bool full_cols = 32 * s + 32 <= n;
          if (warp == kW - 1) {
            const int rr = 32 * s + lane;
            const bool rvalid = rr < n && rr >= t0;
            const float vr = rvalid ? (rr == t0 ? 1.f : __ldcg(coli + rr) * scale) : 0.f;
            float av[32];
#pragma unroll
            for (int cc = 0; cc < 32; ++cc) {
              const int j = 32 * s + cc;
              const bool ok = rvalid && j >= ja && j < jb && rr >= j;
              const float t = __ldcg(a.A + (ok ? (size_t)j * ld + rr : 0));
              av[cc] = ok ? t : 0.f;
            }
            float ac = 0.f;
#pragma unroll
            for (int cc = 0; cc < 32; ++cc) {
              const float vj = __shfl_sync(0xffffffffu, vC, cc);
              ac = fmaf(av[cc], vj, ac);
              // strictly-lower entries also feed their column's sum; the diagonal entry itself is counted once, above
              colacc[cc] = fmaf(av[cc], lane > cc ? vr : 0.f, colacc[cc]);
            }
            if (rvalid) ysm[rr] += ac;
          }
          const int rbase = 32 * s + 32;
          for (int R = warp; rbase + R * 128 < n; R += kW) {
            const int r = rbase + R * 128 + 4 * lane;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            if (full_cols && rbase + R * 128 + 128 <= n) {
              // inside the matrix: 32 unconditional 16-byte loads in flight, then the FMAs.  Columns left of t0 (first
              // strip only) hold finished reflectors: finite values times v = 0.
              const float* p = a.A + (size_t)(32 * s) * ld + r;
              float4 av[32];
#pragma unroll
              for (int cc = 0; cc < 32; ++cc) av[cc] = __ldcg(reinterpret_cast<const float4*>(p + (size_t)cc * ld));
              float4 vr = __ldcg(reinterpret_cast<const float4*>(coli + r));
              vr.x *= scale; vr.y *= scale; vr.z *= scale; vr.w *= scale;
#pragma unroll
              for (int cc = 0; cc < 32; ++cc) {
                const float vj = __shfl_sync(0xffffffffu, vC, cc);
                acc[0] = fmaf(av[cc].x, vj, acc[0]);
                acc[1] = fmaf(av[cc].y, vj, acc[1]);
                acc[2] = fmaf(av[cc].z, vj, acc[2]);
                acc[3] = fmaf(av[cc].w, vj, acc[3]);
                colacc[cc] = fmaf(av[cc].x, vr.x, fmaf(av[cc].y, vr.y, fmaf(av[cc].z, vr.z, fmaf(av[cc].w, vr.w, colacc[cc]))));
              }
              float4* yp = reinterpret_cast<float4*>(ysm + r);   // one warp per row block inside a strip; strips are
              float4 y4 = *yp;                                    // separated by barriers
              y4.x += acc[0]; y4.y += acc[1]; y4.z += acc[2]; y4.w += acc[3];
              *yp = y4;
            } else {
              // the matrix edge (last rows, a partial last strip): every position is masked and read from a safe address
#pragma unroll 1
              for (int q = 0; q < 4; ++q) {
                const int rr = r + q;
                const bool rvalid = rr < n;
                const float vr = rvalid ? __ldcg(coli + rr) * scale : 0.f;
                float av[32];
#pragma unroll
                for (int cc = 0; cc < 32; ++cc) {
                  const int j = 32 * s + cc;
                  const bool ok = rvalid && j >= ja && j < jb;
                  const float t = __ldcg(a.A + (ok ? (size_t)j * ld + rr : 0));
                  av[cc] = ok ? t : 0.f;
                }
                float ac = 0.f;
#pragma unroll
                for (int cc = 0; cc < 32; ++cc) {
                  const float vj = __shfl_sync(0xffffffffu, vC, cc);
                  ac = fmaf(av[cc], vj, ac);
                  colacc[cc] = fmaf(av[cc], vr, colacc[cc]);
                }
                if (rvalid) ysm[rr] += ac;
              }
            }
          }
          float mine = 0.f;
#pragma unroll
          for (int cc = 0; cc < 32; ++cc) {
            const float t = warp_sum_f(colacc[cc]);
            if (lane == cc) mine = t;
          }
          s_col[warp][lane] = mine;
          __syncthreads();
          if (tid < 32) {
            float t = 0.f;
            for (int w = 0; w < kW; ++w) t += s_col[w][tid];
            const int j = 32 * s + tid;
            if (j >= ja && j < jb) __stcg(a.ycol + j, t);
          }
          __syncthreads();
        }
        SY_TICK(3);
        float* Pb = a.P + (size_t)b * ld;
        for (int r = t0 + tid; r < n; r += kT) __stcg(Pb + r, ysm[r]);
      }
      SY_TICK(4);
      grid_barrier(a.bar, target);
      SY_TICK(5);

      // ---- phase 3: w = tau (y - V q - Wf p) on the rows this warp owns (stored without alpha v); w'v; v into A(:, i)
      {
        // 64 dot products x 148 partials: thread (x, quarter) sums every fourth partial of dot product x (independent loads),
        // the four quarters meet in shared memory - fixed order
        const int x = tid & (2 * kNB - 1), ch = tid >> 6;
        double sx = 0;
        if ((x & (kNB - 1)) < c)
#pragma unroll 8
          for (int q = ch; q < P; q += kT / (2 * kNB)) sx += __ldcg(a.part + (size_t)q * 2 * kNB + x);
        s_dd[ch][x] = sx;
        __syncthreads();
        if (tid < 2 * kNB) {
          double t = 0;
          for (int q = 0; q < kT / (2 * kNB); ++q) t += s_dd[q][tid];
          if (tid < kNB) s_p[tid] = (float)t; else s_q[tid - kNB] = (float)t;
        }
      }
      __syncthreads();
      double dot = 0;
      for (int rb = b + P * warp; rb * 32 < n; rb += P * kW) {   // row block rb belongs to CTA rb % P, warp (rb / P) % kW
        const int r = rb * 32 + lane;
        if (r < t0 || r >= n) continue;
        // the row's partial sums of all CTAs: four independent chains, 32 loads in flight
        float y0 = __ldcg(a.ycol + r), y1 = 0.f, y2 = 0.f, y3 = 0.f;
        const float* pp = a.P + r;
        int q = 0;
#pragma unroll 8
        for (; q + 4 <= P; q += 4) {
          y0 += __ldcg(pp + (size_t)q * ld);
          y1 += __ldcg(pp + (size_t)(q + 1) * ld);
          y2 += __ldcg(pp + (size_t)(q + 2) * ld);
          y3 += __ldcg(pp + (size_t)(q + 3) * ld);
        }
        for (; q < P; ++q) y0 += __ldcg(pp + (size_t)q * ld);
        const float y = (y0 + y1) + (y2 + y3);
        const float vr_i = r == t0 ? 1.f : __ldcg(coli + r) * scale;
        float corr = 0.f;
        {
          float vr[kNB], wr[kNB];
          load_panel_row(a, ld, k0, c, r, s_alpha, vr, wr);
#pragma unroll
          for (int cp = 0; cp < kNB; ++cp) {
            if (cp < c) {
              corr = fmaf(vr[cp], s_q[cp], corr);
              corr = fmaf(wr[cp], s_p[cp], corr);
            }
          }
        }
        const float w = tau * (y - corr);
        __stcg(a.W + (size_t)c * ld + r, w);
        dot += (double)w * (double)vr_i;
        __stcg(a.A + (size_t)i * ld + r, vr_i);   // the reflector, its leading one stored explicitly while the panel is open
      }
      dot = cta_sum(dot, s_red);
      if (tid == 0) __stcg(part3 + b, dot);
      SY_TICK(6);
      grid_barrier(a.bar, target);
      SY_TICK(7);
    }
    if (k1 >= n) break;
    // ---- panel end: alpha of its last column, then A22 -= V Wf' + Wf V' on the lower triangle of rows / columns >= k1
    if (warp == 0) {
      const double s = warp_gather_sum(part3, P, 1);
      if (lane == 0) s_alpha[ncols - 1] = (float)(-0.5 * s_scal[0] * s);
    }
    __syncthreads();
    {
      const int u = tid;
      float* Vr = smem;                      // [kNB][kTile] each
      float* Wr = Vr + kNB * kTile;
      float* Vc = Wr + kNB * kTile;
      float* Wc = Vc + kNB * kTile;
      const int m = n - k1, Tn = (m + kTile - 1) / kTile, n_tiles = Tn * (Tn + 1) / 2;
      const int rx = (u & 15) * 8, cy = (u >> 4) * 8;
      for (int t = b; t < n_tiles; t += P) {
        // tile index -> (tr, tc), tc <= tr
        int tr = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
        while (tr * (tr + 1) / 2 > t) --tr;
        while ((tr + 1) * (tr + 2) / 2 <= t) ++tr;
        const int tc = t - tr * (tr + 1) / 2;
        const int r0 = k1 + tr * kTile, c0 = k1 + tc * kTile;
        // k1 is a multiple of 32 and ld of 4: 16-byte loads; rows beyond n inside the last group of four are padding
        for (int x = u; x < kNB * kTile / 4; x += kT) {
          const int k = x / (kTile / 4), o = (x % (kTile / 4)) * 4;
          float4 vr = make_float4(0.f, 0.f, 0.f, 0.f), wr = vr, vc = vr, wc = vr;
          if (k < ncols) {
            const float al = s_alpha[k];
            if (r0 + o < n) {
              vr = __ldcg(reinterpret_cast<const float4*>(a.A + (size_t)(k0 + k) * ld + r0 + o));
              wr = __ldcg(reinterpret_cast<const float4*>(a.W + (size_t)k * ld + r0 + o));
              wr.x = fmaf(al, vr.x, wr.x); wr.y = fmaf(al, vr.y, wr.y); wr.z = fmaf(al, vr.z, wr.z); wr.w = fmaf(al, vr.w, wr.w);
            }
            if (c0 + o < n) {
              vc = __ldcg(reinterpret_cast<const float4*>(a.A + (size_t)(k0 + k) * ld + c0 + o));
              wc = __ldcg(reinterpret_cast<const float4*>(a.W + (size_t)k * ld + c0 + o));
              wc.x = fmaf(al, vc.x, wc.x); wc.y = fmaf(al, vc.y, wc.y); wc.z = fmaf(al, vc.z, wc.z); wc.w = fmaf(al, vc.w, wc.w);
            }
          }
          *reinterpret_cast<float4*>(Vr + k * kTile + o) = vr;
          *reinterpret_cast<float4*>(Wr + k * kTile + o) = wr;
          *reinterpret_cast<float4*>(Vc + k * kTile + o) = vc;
          *reinterpret_cast<float4*>(Wc + k * kTile + o) = wc;
        }
        __syncthreads();
        float acc[8][8];
#pragma unroll
        for (int x = 0; x < 8; ++x)
#pragma unroll
          for (int y = 0; y < 8; ++y) acc[x][y] = 0.f;
        for (int k = 0; k < ncols; ++k) {
          float vr[8], wr[8], vc[8], wc[8];
          *reinterpret_cast<float4*>(vr) = *reinterpret_cast<const float4*>(Vr + k * kTile + rx);
          *reinterpret_cast<float4*>(vr + 4) = *reinterpret_cast<const float4*>(Vr + k * kTile + rx + 4);
          *reinterpret_cast<float4*>(wr) = *reinterpret_cast<const float4*>(Wr + k * kTile + rx);
          *reinterpret_cast<float4*>(wr + 4) = *reinterpret_cast<const float4*>(Wr + k * kTile + rx + 4);
          *reinterpret_cast<float4*>(vc) = *reinterpret_cast<const float4*>(Vc + k * kTile + cy);
          *reinterpret_cast<float4*>(vc + 4) = *reinterpret_cast<const float4*>(Vc + k * kTile + cy + 4);
          *reinterpret_cast<float4*>(wc) = *reinterpret_cast<const float4*>(Wc + k * kTile + cy);
          *reinterpret_cast<float4*>(wc + 4) = *reinterpret_cast<const float4*>(Wc + k * kTile + cy + 4);
#pragma unroll
          for (int x = 0; x < 8; ++x)
#pragma unroll
            for (int y = 0; y < 8; ++y) acc[x][y] = fmaf(vr[x], wc[y], fmaf(wr[x], vc[y], acc[x][y]));
        }
        // write-back: 8 consecutive rows per column as two 16-byte read-modify-writes when they are all below the diagonal
#pragma unroll
        for (int y = 0; y < 8; ++y) {
          const int cc = c0 + cy + y;
          if (cc >= n) continue;
          const int rr = r0 + rx;
          float* p = a.A + (size_t)cc * ld + rr;
          if (rr >= cc && rr + 8 <= n) {
            float4 lo4 = __ldcg(reinterpret_cast<const float4*>(p)), hi4 = __ldcg(reinterpret_cast<const float4*>(p + 4));
            lo4.x -= acc[0][y]; lo4.y -= acc[1][y]; lo4.z -= acc[2][y]; lo4.w -= acc[3][y];
            hi4.x -= acc[4][y]; hi4.y -= acc[5][y]; hi4.z -= acc[6][y]; hi4.w -= acc[7][y];
            __stcg(reinterpret_cast<float4*>(p), lo4);
            __stcg(reinterpret_cast<float4*>(p + 4), hi4);
          } else {
#pragma unroll
            for (int x = 0; x < 8; ++x) {
              const int r = rr + x;
              if (r < n && r >= cc) __stcg(p + x, __ldcg(p + x) - acc[x][y]);
            }
          }
        }
        __syncthreads();
      }
    }
    SY_TICK(8);
    grid_barrier(a.bar, target);
    SY_TICK(9);
  }
  if (tid == 0 && a.prof)
    for (int q = 0; q < 10; ++q) a.prof[(size_t)b * 16 + q] = tacc[q];
}

// the sub-diagonal of T goes back where ssytrd leaves it (the reflectors' leading ones were stored there while their panel was open)
__global__ void k_restore_subdiagonal(float* __restrict__ A, int n, int ld, const float* __restrict__ e) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n - 1) A[(size_t)i * ld + i + 1] = e[i];
}

}  // namespace

// A: symmetric n x n, column-major with leading dimension lda (a multiple of 4, 16-byte aligned base; the lower triangle is
// read); on return as after ssytrd('L'): d (n), e (n - 1), tau (n - 1), reflectors below the sub-diagonal.  Returns false
// when the matrix is outside what the kernel handles (the caller uses the library's Ssytrd then).
bool sytrd_lower(float* dA, int n, int lda, float* d_d, float* d_e, float* d_tau, cudaStream_t st) {
  if (n < 256 || n > 48000 || lda < n || (lda & 3) || ((uintptr_t)dA & 15)) return false;
  int dev = 0, sms = 0, coop = 0;
  SCL_CUDA(cudaGetDevice(&dev));
  SCL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  SCL_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  if (!coop) return false;
  const size_t smem = std::max((size_t)lda * sizeof(float), (size_t)4 * kNB * kTile * sizeof(float));
  if (smem > 200 * 1024) return false;
  SCL_CUDA(cudaFuncSetAttribute(k_sytrd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  SCL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sytrd, kT, smem));
  if (per_sm < 1) return false;
  const int grid = sms;
  Tmp<float> W((size_t)lda * kNB, st), P((size_t)grid * lda, st), ycol(n, st), scal(4, st);
  Tmp<double> part((size_t)grid * (2 * kNB + 2), st);
  Tmp<unsigned> bar(1, st);
  SCL_CUDA(cudaMemsetAsync(bar.p, 0, sizeof(unsigned), st));
  SCL_CUDA(cudaMemsetAsync(part.p, 0, (size_t)grid * (2 * kNB + 2) * sizeof(double), st));
  SCL_CUDA(cudaMemsetAsync(d_tau, 0, (size_t)n * sizeof(float), st));
  SCL_CUDA(cudaMemsetAsync(d_e, 0, (size_t)n * sizeof(float), st));
  static const bool trace = getenv("SCL_TRACE") != nullptr;
  Tmp<unsigned long long> prof(trace ? (size_t)grid * 16 : 1, st);
  SyArgs args{dA, n, lda, W.p, d_d, d_e, d_tau, P.p, ycol.p, part.p, scal.p, bar.p, trace ? prof.p : nullptr};
  void* params[] = {&args};
  count_launches(2);
  // a cooperative launch is all or nothing: if the device cannot hold one CTA per SM right now (another context shares
  // it), the launch is refused and the caller takes the library's tridiagonalisation instead
  if (cudaLaunchCooperativeKernel((void*)k_sytrd, dim3(grid), dim3(kT), params, smem, st) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  k_restore_subdiagonal<<<(n + 255) / 256, 256, 0, st>>>(dA, n, lda, d_e);
  SCL_CUDA(cudaGetLastError());
  if (trace) {
    std::vector<unsigned long long> hp((size_t)grid * 16);
    SCL_CUDA(cudaMemcpyAsync(hp.data(), prof.p, hp.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    SCL_CUDA(cudaStreamSynchronize(st));
    static const char* name[10] = {"phase1 (column update)", "barrier 1", "phase2 scalars + panel dots", "symv strips", "row partials out",
                                   "barrier 2", "phase3 (w)", "barrier 3", "rank-2k update", "barrier 4"};
    for (int q = 0; q < 10; ++q) {
      double mn = 1e30, mx = 0, av = 0;
      for (int g = 0; g < grid; ++g) {
        const double v = (double)hp[(size_t)g * 16 + q] * 1e-6;
        mn = std::min(mn, v); mx = std::max(mx, v); av += v / grid;
      }
      fprintf(stderr, "[scl] sytrd n=%d %-28s ms per CTA: min %9.2f  mean %9.2f  max %9.2f\n", n, name[q], mn, av, mx);
    }
  }
  return true;
}

}  // namespace scl
