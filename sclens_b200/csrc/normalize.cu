// Normalisation of src/scLENS.jl:677-696 (== logn_scale(pre_scale(X)), :650-652, :596-608)
// computed from the sparse matrix only and emitted directly as the
// dense binary16 Gram operand.  Statistics are Float64 and deterministic (fixed reduction
// trees, no floating-point atomics); the dense N x M matrix is written exactly once.
//
//   r_i = sum_j x_ij                      y_ij = log1p(x_ij * (1/r_i))            (:678-681)
//   ybar_j = mean_i y, sigma_j = std_i y (corrected, zeros included)             (:682-683)
//   z_ij = y_ij / sigma_j, mu_j = ybar_j / sigma_j                               (:685-686)
//   l_i = sqrt(sum_j z^2 - 2 sum_j z mu + |mu|^2), s_i = l_i / mean(l)           (:688-693)
//   c_j = mean_i (z_ij - mu_j)/s_i                                               (:695)
//   out_ij = (z_ij - mu_j)/s_i - c_j                                             (:696)
#include "common.cuh"
#include "tmp.cuh"

namespace scl {

static constexpr int kDenseChunk = 8192;   // positions per shared-memory strip
static constexpr int kDenseThreads = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block sum (all threads receive the result).  blockDim multiple of 32, <= 1024.
__device__ __forceinline__ double block_sum(double v, double* red /* >= 32 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0;
  for (int w = 0; w < nw; ++w) s += red[w];
  return s;
}

// ---- per-cell total counts (CSR, one warp per cell) -------------------------------------
__global__ void k_row_sum(const uint32_t* __restrict__ rowptr, const float* __restrict__ rval, int N,
                          double* __restrict__ tgc) {
  const int lane = threadIdx.x & 31;
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nrows_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (; row < N; row += nrows_per_grid) {
    double s = 0;
    for (uint32_t t = rowptr[row] + lane; t < rowptr[row + 1]; t += 32) s += (double)rval[t];
    s = warp_sum(s);
    if (lane == 0) tgc[row] = s;
  }
}

// ---- per-gene mean / corrected std of y (CSC, one block per gene) ------------------------
__global__ void __launch_bounds__(128) k_gene_stats(const uint32_t* __restrict__ colptr,
                                                    const uint32_t* __restrict__ rowval,
                                                    const float* __restrict__ val, const double* __restrict__ tgc,
                                                    int N, int M, double* __restrict__ ybar,
                                                    double* __restrict__ sigma, double* __restrict__ mu,
                                                    float* __restrict__ mu_f, float* __restrict__ inv_sigma_f) {
  __shared__ double red[32];
  for (int j = blockIdx.x; j < M; j += gridDim.x) {
    const uint32_t b = colptr[j], e = colptr[j + 1];
    double s1 = 0;
    for (uint32_t t = b + threadIdx.x; t < e; t += blockDim.x)
      s1 += log1p((double)val[t] * (1.0 / tgc[rowval[t]]));
    s1 = block_sum(s1, red);
    const double m = s1 / (double)N;
    double s2 = 0;
    for (uint32_t t = b + threadIdx.x; t < e; t += blockDim.x) {
      double d = log1p((double)val[t] * (1.0 / tgc[rowval[t]])) - m;
      s2 += d * d;
    }
    s2 = block_sum(s2, red);
    if (threadIdx.x == 0) {
      double var = (s2 + (double)(N - (int)(e - b)) * m * m) / (double)(N - 1);
      double sd = sqrt(var);
      ybar[j] = m;
      sigma[j] = sd;
      mu[j] = m / sd;
      mu_f[j] = (float)(m / sd);
      inv_sigma_f[j] = (float)(1.0 / sd);
    }
  }
}

// ---- small deterministic reductions (one block) -------------------------------------------
// mode 0: out = sum v^2 ; mode 1: out = mean v ; mode 2: out = sum v
__global__ void __launch_bounds__(1024) k_reduce(const double* __restrict__ v, int n, int mode,
                                                 double* __restrict__ out) {
  __shared__ double red[32];
  double s = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double x = v[i];
    s += mode == 0 ? x * x : x;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) *out = mode == 1 ? s / (double)n : s;
}

// ---- per-cell l2 norm after mean shift (CSR, one warp per cell) ---------------------------
__global__ void k_cell_l2(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ colidx,
                          const float* __restrict__ rval, const double* __restrict__ tgc,
                          const double* __restrict__ sigma, const double* __restrict__ mu,
                          const double* __restrict__ scalars, int N, double* __restrict__ l2) {
  const int lane = threadIdx.x & 31;
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nrows_per_grid = (gridDim.x * blockDim.x) >> 5;
  const double mu2 = scalars[0];
  for (; row < N; row += nrows_per_grid) {
    const double inv_r = 1.0 / tgc[row];
    double a = 0, b = 0;
    for (uint32_t t = rowptr[row] + lane; t < rowptr[row + 1]; t += 32) {
      uint32_t c = colidx[t];
      double z = log1p((double)rval[t] * inv_r) / sigma[c];
      a += z * z;
      b += z * mu[c];
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) l2[row] = sqrt(a - 2.0 * b + mu2);
  }
}

__global__ void k_inv_s(const double* __restrict__ l2, const double* __restrict__ scalars, int N,
                        double* __restrict__ inv_s, float* __restrict__ inv_s_f) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) {
    double v = scalars[1] / l2[i];
    inv_s[i] = v;
    inv_s_f[i] = (float)v;
  }
}

// ---- per-gene centre after cell scaling (CSC, one block per gene) --------------------------
__global__ void __launch_bounds__(128) k_gene_center(const uint32_t* __restrict__ colptr,
                                                     const uint32_t* __restrict__ rowval,
                                                     const float* __restrict__ val, const double* __restrict__ tgc,
                                                     const double* __restrict__ sigma, const double* __restrict__ mu,
                                                     const double* __restrict__ inv_s,
                                                     const double* __restrict__ scalars, int N, int M,
                                                     double* __restrict__ cent, float* __restrict__ cent_f) {
  __shared__ double red[32];
  const double sum_inv_s = scalars[2];
  for (int j = blockIdx.x; j < M; j += gridDim.x) {
    const uint32_t b = colptr[j], e = colptr[j + 1];
    const double sd = sigma[j];
    double s = 0;
    for (uint32_t t = b + threadIdx.x; t < e; t += blockDim.x) {
      uint32_t r = rowval[t];
      s += log1p((double)val[t] * (1.0 / tgc[r])) / sd * inv_s[r];
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) {
      double c = (s - mu[j] * sum_inv_s) / (double)N;
      cent[j] = c;
      cent_f[j] = (float)c;
    }
  }
}

void compute_norm_stats(const SpMat& A, NormStats& S, cudaStream_t st) {
  const int N = A.N, M = A.M;
  SCL_REQUIRE(N > 1 && M > 0 && A.nnz > 0, "matrix too small to normalise");
  S.tgc.ensure(N); S.l2.ensure(N); S.inv_s.ensure(N); S.inv_s_f.ensure(N);
  S.ybar.ensure(M); S.sigma.ensure(M); S.mu.ensure(M); S.cent.ensure(M);
  S.mu_f.ensure(M); S.cent_f.ensure(M); S.inv_sigma_f.ensure(M);
  S.scalars.ensure(4);
  count_launches(8);
  const int wgrid = min((N + 7) / 8, 148 * 8);
  k_row_sum<<<wgrid, 256, 0, st>>>(A.rowptr.p, A.rval.p, N, S.tgc.p);
  k_gene_stats<<<min(M, 148 * 16), 128, 0, st>>>(A.colptr.p, A.rowval.p, A.val.p, S.tgc.p, N, M, S.ybar.p, S.sigma.p,
                                                  S.mu.p, S.mu_f.p, S.inv_sigma_f.p);
  k_reduce<<<1, 1024, 0, st>>>(S.mu.p, M, 0, S.scalars.p + 0);
  k_cell_l2<<<wgrid, 256, 0, st>>>(A.rowptr.p, A.colidx.p, A.rval.p, S.tgc.p, S.sigma.p, S.mu.p, S.scalars.p, N,
                                   S.l2.p);
  k_reduce<<<1, 1024, 0, st>>>(S.l2.p, N, 1, S.scalars.p + 1);
  k_inv_s<<<(N + 255) / 256, 256, 0, st>>>(S.l2.p, S.scalars.p, N, S.inv_s.p, S.inv_s_f.p);
  k_reduce<<<1, 1024, 0, st>>>(S.inv_s.p, N, 2, S.scalars.p + 2);
  k_gene_center<<<min(M, 148 * 16), 128, 0, st>>>(A.colptr.p, A.rowval.p, A.val.p, S.tgc.p, S.sigma.p, S.mu.p,
                                                   S.inv_s.p, S.scalars.p, N, M, S.cent.p, S.cent_f.p);
  SCL_CUDA(cudaGetLastError());
}

// ---- fused densify + normalise writer -----------------------------------------------------
// One CTA per (line, strip): fill the rank-structured background -mu_j/s_i - c_j into shared
// memory, patch the line's non-zeros, then stream the strip out with 16-byte stores in the
// Gram operand dtype.  CELL_MAJOR=false: line = gene j (CSC), positions = cells;
// CELL_MAJOR=true: line = cell i (CSR mirror), positions = genes.
template <bool CELL_MAJOR, bool WITH_LO>
__global__ void __launch_bounds__(kDenseThreads)
k_densify(const uint32_t* __restrict__ ptr, const uint32_t* __restrict__ idx, const float* __restrict__ val,
          const double* __restrict__ tgc, const double* __restrict__ sigma, const double* __restrict__ mu,
          const double* __restrict__ inv_s, const double* __restrict__ cent, const float* __restrict__ inv_s_f,
          const float* __restrict__ mu_f, const float* __restrict__ cent_f, int n_lines, int line_len, size_t ld,
          __half* __restrict__ out_hi, __half* __restrict__ out_lo, double* __restrict__ sumsq_partial, int pos0,
          int pos1) {
  __shared__ __align__(16) float tile[kDenseChunk];
  __shared__ double red_sq[kDenseThreads / 32];
  // positions [pos0, pos1) of every line are emitted (pos0 multiple of 8; the whole padded line by default)
  const int n_strips = (pos1 - pos0 + kDenseChunk - 1) / kDenseChunk;
  const long long total = (long long)n_lines * n_strips;
  for (long long w = blockIdx.x; w < total; w += gridDim.x) {
    const int line = (int)(w / n_strips);
    const int base = pos0 + (int)(w % n_strips) * kDenseChunk;
    const int span = min(kDenseChunk, pos1 - base);                  // includes the zero pad up to ld
    const int len = max(0, min(span, line_len - base));              // real positions
    // background
    if (CELL_MAJOR) {
      const float a = -inv_s_f[line];
      for (int i = threadIdx.x; i < span; i += kDenseThreads)
        tile[i] = i < len ? fmaf(a, mu_f[base + i], -cent_f[base + i]) : 0.f;
    } else {
      const float a = -mu_f[line], c = -cent_f[line];
      for (int i = threadIdx.x; i < span; i += kDenseThreads) tile[i] = i < len ? fmaf(a, inv_s_f[base + i], c) : 0.f;
    }
    __syncthreads();
    // patch non-zeros of this line that fall into the strip (line entries are sorted)
    const uint32_t sb = ptr[line], se = ptr[line + 1];
    uint32_t lo = sb, hi = se;
    while (lo < hi) {
      uint32_t mid = (lo + hi) >> 1;
      if (idx[mid] < (uint32_t)base) lo = mid + 1; else hi = mid;
    }
    if (CELL_MAJOR) {
      const double inv_r = 1.0 / tgc[line], is = inv_s[line];
      for (uint32_t t = lo + threadIdx.x; t < se; t += kDenseThreads) {
        uint32_t p = idx[t];
        if (p >= (uint32_t)(base + len)) break;
        double z = log1p((double)val[t] * inv_r) / sigma[p];
        tile[p - base] = (float)((z - mu[p]) * is - cent[p]);
      }
    } else {
      const double sd = sigma[line], m = mu[line], c = cent[line];
      for (uint32_t t = lo + threadIdx.x; t < se; t += kDenseThreads) {
        uint32_t p = idx[t];
        if (p >= (uint32_t)(base + len)) break;
        double z = log1p((double)val[t] * (1.0 / tgc[p])) / sd;
        tile[p - base] = (float)((z - m) * inv_s[p] - c);
      }
    }
    __syncthreads();
    // stream out: 8 elements (16 bytes) per store
    __half* dst_hi = out_hi + (size_t)line * ld + base;
    __half* dst_lo = WITH_LO ? out_lo + (size_t)line * ld + base : nullptr;
    double sq = 0;   // exact sum of squares of the emitted (rounded) values: the Gram diagonal
    for (int i = threadIdx.x * 8; i < span; i += kDenseThreads * 8) {
      float4 v0 = *reinterpret_cast<const float4*>(&tile[i]);
      float4 v1 = *reinterpret_cast<const float4*>(&tile[i + 4]);
      float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      __align__(16) __half h[8];
      __align__(16) __half l[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        h[q] = __float2half_rn(f[q]);
        if (WITH_LO) l[q] = __float2half_rn(f[q] - __half2float(h[q]));
        const double e = WITH_LO ? (double)__half2float(h[q]) + (double)__half2float(l[q]) : (double)__half2float(h[q]);
        sq += e * e;
      }
      *reinterpret_cast<uint4*>(dst_hi + i) = *reinterpret_cast<const uint4*>(h);
      if (WITH_LO) *reinterpret_cast<uint4*>(dst_lo + i) = *reinterpret_cast<const uint4*>(l);
    }
    if (sumsq_partial) {
      sq = warp_sum(sq);
      if ((threadIdx.x & 31) == 0) red_sq[threadIdx.x >> 5] = sq;
      __syncthreads();
      if (threadIdx.x == 0) {
        double t = 0;
        for (int q = 0; q < kDenseThreads / 32; ++q) t += red_sq[q];
        sumsq_partial[w] = t;   // [line][strip]
      }
    }
    __syncthreads();
  }
}

// G[i][i] = scale * sum over the line's strips of the exact sums of squares (fixed order: deterministic)
__global__ void k_set_diagonal(float* __restrict__ G, int n, int n_strips, const double* __restrict__ partial, double scale) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0;
  for (int q = 0; q < n_strips; ++q) s += partial[(size_t)i * n_strips + q];
  G[(size_t)i * n + i] = (float)(s * scale);
}

int densify_strips(size_t n_pos) { return (int)((n_pos + kDenseChunk - 1) / kDenseChunk); }

void set_gram_diagonal(float* G, int n, int n_strips, const double* partial, double scale, cudaStream_t st) {
  count_launches(1);
  k_set_diagonal<<<(n + 255) / 256, 256, 0, st>>>(G, n, n_strips, partial, scale);
  SCL_CUDA(cudaGetLastError());
}

void densify(const SpMat& A, const NormStats& S, int layout, size_t ld, __half* out_hi, __half* out_lo,
             cudaStream_t st, double* sumsq_partial, long long pos0, long long pos1) {
  count_launches(1);
  const bool cell_major = layout == 1;
  const int n_lines = cell_major ? A.N : A.M;
  const int line_len = cell_major ? A.M : A.N;
  SCL_REQUIRE(ld % 8 == 0 && ld >= (size_t)line_len, "leading dimension must be a multiple of 8 and >= line length");
  if (pos1 < 0) { pos0 = 0; pos1 = (long long)ld; }
  SCL_REQUIRE(pos0 % 8 == 0 && pos0 >= 0 && pos1 <= (long long)ld && pos1 > pos0 && (pos1 % 8 == 0 || pos1 == (long long)ld), "bad densify range");
  const long long total = (long long)n_lines * ((pos1 - pos0 + kDenseChunk - 1) / kDenseChunk);
  const int grid = (int)(total < 148LL * 6 ? total : 148LL * 6);
  const uint32_t* ptr = cell_major ? A.rowptr.p : A.colptr.p;
  const uint32_t* idx = cell_major ? A.colidx.p : A.rowval.p;
  const float* val = cell_major ? A.rval.p : A.val.p;
#define SCL_LAUNCH_DENSIFY(CM, LO)                                                                              \
  k_densify<CM, LO><<<grid, kDenseThreads, 0, st>>>(ptr, idx, val, S.tgc.p, S.sigma.p, S.mu.p, S.inv_s.p,       \
                                                    S.cent.p, S.inv_s_f.p, S.mu_f.p, S.cent_f.p, n_lines,       \
                                                    line_len, ld, out_hi, out_lo, sumsq_partial, (int)pos0,   \
                                                    (int)pos1)
  if (cell_major) {
    if (out_lo) SCL_LAUNCH_DENSIFY(true, true); else SCL_LAUNCH_DENSIFY(true, false);
  } else {
    if (out_lo) SCL_LAUNCH_DENSIFY(false, true); else SCL_LAUNCH_DENSIFY(false, false);
  }
#undef SCL_LAUNCH_DENSIFY
  SCL_CUDA(cudaGetLastError());
}

}  // namespace scl
