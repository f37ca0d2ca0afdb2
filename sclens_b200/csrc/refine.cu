// Float64 Rayleigh-quotient refinement of the eigenvalues returned by the FP32 eigensolver.
//
// cusolverDnSsyevd (the reference's solver, src/scLENS.jl:377) is backward stable in FP32: every eigenvalue carries
// an ABSOLUTE error of order eps32 * |G| (measured 1.9e-5 on the 10k x 10k benchmark Gram, whose largest
// eigenvalue is 21), i.e. 2e-4 RELATIVE at the lower Marchenko-Pastur edge - above the 1e-4 contract although the
// matrix itself is good to 3e-5.  The Rayleigh quotient of a computed eigenvector is accurate to second order in
// the eigenvector error, so   rho_i = v_i' G v_i / v_i' v_i   evaluated in Float64 from the FP32 eigenvectors and
// the FP32 Gram matrix recovers the eigenvalues of G to ~1e-6 relative.  One n x n x n Float64 contraction per
// sclens() call (only the data matrix's spectrum is an output), ~1 % of the eigensolver's time.
//
// Kernel: classic register-tiled GEMM  Z = V * G  (G symmetric, so both operands are read K-major), 128 x 128
// output tile per CTA, 8 x 8 per thread, FP32 operands staged in shared memory and widened in registers, FP64
// accumulation; the epilogue contracts the tile of Z with the matching tile of V, so Z is never written.
#include "common.cuh"
#include "tmp.cuh"

namespace scl {

namespace {

constexpr int kTile = 128;
constexpr int kStepK = 16;
constexpr int kRqThreads = 256;

// part[i * n_tiles + tc] = sum over c in column tile tc of (V G)[i][c] * V[i][c]
__global__ void __launch_bounds__(kRqThreads)
k_rq_partial(const float* __restrict__ V, const float* __restrict__ G, int n, int n_tiles, double* __restrict__ part) {
  __shared__ __align__(16) float As[2][kStepK][kTile];   // V tile, [k][row]
  __shared__ __align__(16) float Bs[2][kStepK][kTile];   // G tile, [k][col]
  const int tid = threadIdx.x;
  const int tr = blockIdx.y, tc = blockIdx.x;
  const int row0 = tr * kTile, col0 = tc * kTile;
  const int tx = tid & 15, ty = tid >> 4;                // 16 x 16 threads
  // global -> shared: thread loads rows lr and lr + 64 (of both operands), 4 consecutive k
  const int lr = tid & 63, lk = (tid >> 6) * 4;          // 64 rows x 4 k-groups of 4
  double acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.0;

  auto load = [&](int kb, float4 (&va)[2], float4 (&vb)[2]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = row0 + lr + 64 * h, c = col0 + lr + 64 * h, k = kb + lk;
      float t[4] = {0.f, 0.f, 0.f, 0.f}, u[4] = {0.f, 0.f, 0.f, 0.f};
      if (r < n) {
        if (k + 3 < n && (n & 3) == 0) {
          const float4 q = *reinterpret_cast<const float4*>(V + (size_t)r * n + k);
          t[0] = q.x; t[1] = q.y; t[2] = q.z; t[3] = q.w;
        } else {
          for (int q = 0; q < 4; ++q) if (k + q < n) t[q] = V[(size_t)r * n + k + q];
        }
      }
      if (c < n) {
        if (k + 3 < n && (n & 3) == 0) {
          const float4 q = *reinterpret_cast<const float4*>(G + (size_t)c * n + k);
          u[0] = q.x; u[1] = q.y; u[2] = q.z; u[3] = q.w;
        } else {
          for (int q = 0; q < 4; ++q) if (k + q < n) u[q] = G[(size_t)c * n + k + q];
        }
      }
      va[h] = make_float4(t[0], t[1], t[2], t[3]);
      vb[h] = make_float4(u[0], u[1], u[2], u[3]);
    }
  };
  auto stash = [&](int buf, const float4 (&va)[2], const float4 (&vb)[2]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lr + 64 * h;
      As[buf][lk + 0][r] = va[h].x; As[buf][lk + 1][r] = va[h].y; As[buf][lk + 2][r] = va[h].z; As[buf][lk + 3][r] = va[h].w;
      Bs[buf][lk + 0][r] = vb[h].x; Bs[buf][lk + 1][r] = vb[h].y; Bs[buf][lk + 2][r] = vb[h].z; Bs[buf][lk + 3][r] = vb[h].w;
    }
  };

  float4 va[2], vb[2];
  load(0, va, vb);
  stash(0, va, vb);
  __syncthreads();
  int buf = 0;
  for (int kb = 0; kb < n; kb += kStepK, buf ^= 1) {
    const bool more = kb + kStepK < n;
    if (more) load(kb + kStepK, va, vb);
#pragma unroll
    for (int k = 0; k < kStepK; ++k) {
      // rows ty*4..+3 and 64+ty*4..+3; columns tx*4..+3 and 64+tx*4..+3 (conflict-free float4 reads)
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const double a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const double b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    if (more) stash(buf ^ 1, va, vb);
    __syncthreads();
  }
  // epilogue: contract with V[i][c] over this tile's columns, reduce over the 16 threads that share the rows
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = row0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    double s = 0.0;
    if (r < n) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = col0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
        if (c < n) s = fma(acc[i][j], (double)V[(size_t)r * n + c], s);
      }
    }
#pragma unroll
    for (int o = 8; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);   // lanes of one ty: 16 consecutive
    if (tx == 0 && r < n) part[(size_t)r * n_tiles + tc] = s;
  }
}

// w[i] = sum_t part[i][t] / |v_i|^2   (one warp per eigenvector, fixed order)
__global__ void __launch_bounds__(256) k_rq_finish(const float* __restrict__ V, const double* __restrict__ part, int n,
                                                   int n_tiles, float* __restrict__ w) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  double vv = 0.0, s = 0.0;
  for (int c = lane; c < n; c += 32) {
    const double x = (double)V[(size_t)i * n + c];
    vv = fma(x, x, vv);
  }
  for (int t = lane; t < n_tiles; t += 32) s += part[(size_t)i * n_tiles + t];
  for (int o = 16; o; o >>= 1) {
    vv += __shfl_xor_sync(0xffffffffu, vv, o);
    s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  if (lane == 0) w[i] = (float)(s / vv);
}

}  // namespace

// dV: eigenvectors as rows of memory (cuSOLVER's column-major result), dG: the matrix they belong to (symmetric)
void refine_eigenvalues(const float* dG, const float* dV, int n, float* dW, cudaStream_t st) {
  const int n_tiles = (n + kTile - 1) / kTile;
  Tmp<double> part((size_t)n * n_tiles, st);
  count_launches(2);
  k_rq_partial<<<dim3(n_tiles, n_tiles), kRqThreads, 0, st>>>(dV, dG, n, n_tiles, part.p);
  k_rq_finish<<<(n * 32 + 255) / 256, 256, 0, st>>>(dV, part.p, n, n_tiles, dW);
  SCL_CUDA(cudaGetLastError());
}

}  // namespace scl
