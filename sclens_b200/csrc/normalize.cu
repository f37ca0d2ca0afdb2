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
#include <algorithm>
#include "tmp.cuh"

namespace scl {

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
// y = log1p(x / r_i) is evaluated once per non-zero and kept (Float64, CSC order) for the second variance
// pass and for k_gene_center.
__global__ void __launch_bounds__(128) k_gene_stats(const uint32_t* __restrict__ colptr,
                                                    const uint32_t* __restrict__ rowval,
                                                    const float* __restrict__ val, const double* __restrict__ tgc,
                                                    int N, int M, double* __restrict__ y_csc, double* __restrict__ ybar,
                                                    double* __restrict__ sigma, double* __restrict__ mu,
                                                    float* __restrict__ mu_f, float* __restrict__ inv_sigma_f) {
  __shared__ double red[32];
  for (int j = blockIdx.x; j < M; j += gridDim.x) {
    const uint32_t b = colptr[j], e = colptr[j + 1];
    double s1 = 0;
    for (uint32_t t = b + threadIdx.x; t < e; t += blockDim.x) {
      const double y = log1p((double)val[t] * (1.0 / tgc[rowval[t]]));
      y_csc[t] = y;
      s1 += y;
    }
    s1 = block_sum(s1, red);
    const double m = s1 / (double)N;
    double s2 = 0;
    for (uint32_t t = b + threadIdx.x; t < e; t += blockDim.x) {   // the thread re-reads its own stores
      double d = y_csc[t] - m;
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
// z = y / sigma_j is kept (Float64, CSR order): the cell-major writer turns it into its sparse patch.
__global__ void k_cell_l2(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ colidx,
                          const float* __restrict__ rval, const double* __restrict__ tgc,
                          const double* __restrict__ sigma, const double* __restrict__ mu,
                          const double* __restrict__ scalars, int N, double* __restrict__ z_csr,
                          double* __restrict__ l2) {
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
      z_csr[t] = z;
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
// Also emits the gene-major writer's sparse patch  z_ij / s_i  (Float32, CSC order).
__global__ void __launch_bounds__(128) k_gene_center(const uint32_t* __restrict__ colptr,
                                                     const uint32_t* __restrict__ rowval,
                                                     const double* __restrict__ y_csc,
                                                     const double* __restrict__ sigma, const double* __restrict__ mu,
                                                     const double* __restrict__ inv_s,
                                                     const double* __restrict__ scalars, int N, int M,
                                                     float* __restrict__ patch_csc, double* __restrict__ cent,
                                                     float* __restrict__ cent_f) {
  __shared__ double red[32];
  const double sum_inv_s = scalars[2];
  for (int j = blockIdx.x; j < M; j += gridDim.x) {
    const uint32_t b = colptr[j], e = colptr[j + 1];
    const double sd = sigma[j];
    double s = 0;
    for (uint32_t t = b + threadIdx.x; t < e; t += blockDim.x) {
      const double w = y_csc[t] / sd * inv_s[rowval[t]];
      patch_csc[t] = (float)w;
      s += w;
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
  S.y_csc.ensure(A.nnz); S.z_csr.ensure(A.nnz); S.patch_csc.ensure(A.nnz);
  count_launches(8);
  const int wgrid = min((N + 7) / 8, 148 * 8);
  k_row_sum<<<wgrid, 256, 0, st>>>(A.rowptr.p, A.rval.p, N, S.tgc.p);
  k_gene_stats<<<min(M, 148 * 16), 128, 0, st>>>(A.colptr.p, A.rowval.p, A.val.p, S.tgc.p, N, M, S.y_csc.p, S.ybar.p,
                                                  S.sigma.p, S.mu.p, S.mu_f.p, S.inv_sigma_f.p);
  k_reduce<<<1, 1024, 0, st>>>(S.mu.p, M, 0, S.scalars.p + 0);
  k_cell_l2<<<wgrid, 256, 0, st>>>(A.rowptr.p, A.colidx.p, A.rval.p, S.tgc.p, S.sigma.p, S.mu.p, S.scalars.p, N,
                                   S.z_csr.p, S.l2.p);
  k_reduce<<<1, 1024, 0, st>>>(S.l2.p, N, 1, S.scalars.p + 1);
  k_inv_s<<<(N + 255) / 256, 256, 0, st>>>(S.l2.p, S.scalars.p, N, S.inv_s.p, S.inv_s_f.p);
  k_reduce<<<1, 1024, 0, st>>>(S.inv_s.p, N, 2, S.scalars.p + 2);
  k_gene_center<<<min(M, 148 * 16), 128, 0, st>>>(A.colptr.p, A.rowval.p, S.y_csc.p, S.sigma.p, S.mu.p, S.inv_s.p,
                                                   S.scalars.p, N, M, S.patch_csc.p, S.cent.p, S.cent_f.p);
  SCL_CUDA(cudaGetLastError());
}

// ---- fused densify + normalise writer -----------------------------------------------------
// out_ij = z_ij/s_i - mu_j/s_i - c_j: a rank-structured background plus a sparse patch.  A CTA owns a strip of
// kStripW positions for a block of lines.  Each thread keeps the per-position factors of its eight positions
// in registers for the whole line block, so the background costs one FMA per element and no memory traffic;
// the patches of a line reach their owner threads through an 8 KB shared-memory overlay (double buffered,
// one barrier per line; the next line's patch loads are in flight while the current line is written).  Every
// thread emits one 16-byte store per line and matrix (hi, optional lo), 512 contiguous bytes per warp.
// CELL_MAJOR=false: line = gene j (CSC), positions = cells; CELL_MAJOR=true: line = cell i (CSR), positions = genes.
static constexpr int kElemsPerThread = 8;
static constexpr int kStripW = kDenseThreads * kElemsPerThread;   // 2048 positions
static constexpr int kLinesPerCta = 32;
// overlay slot of strip-relative position r: thread r/8 reads its positions as two conflict-free float4 planes
__device__ __forceinline__ uint32_t slot_of(uint32_t r) { return ((r & 4u) ? (uint32_t)(kStripW / 2) : 0u) + ((r >> 3) << 2) + (r & 3u); }

// off[line * (n_strips + 1) + s] = first entry of the line at position >= pos0 + s * kStripW
__global__ void k_strip_offsets(const uint32_t* __restrict__ ptr, const uint32_t* __restrict__ idx, int n_lines,
                                int n_strips, int pos0, uint32_t* __restrict__ off) {
  const long long total = (long long)n_lines * (n_strips + 1);
  for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < total; w += (long long)gridDim.x * blockDim.x) {
    const int line = (int)(w / (n_strips + 1)), s = (int)(w % (n_strips + 1));
    const uint32_t target = (uint32_t)pos0 + (uint32_t)s * (uint32_t)kStripW;
    uint32_t lo = ptr[line], hi = ptr[line + 1];
    while (lo < hi) {
      uint32_t mid = (lo + hi) >> 1;
      if (idx[mid] < target) lo = mid + 1; else hi = mid;
    }
    off[w] = lo;
  }
}

template <bool CELL_MAJOR, bool WITH_LO, bool WITH_SQ>
__global__ void __launch_bounds__(kDenseThreads)
k_densify(const uint32_t* __restrict__ off, const uint32_t* __restrict__ idx, const double* __restrict__ z_csr,
          const float* __restrict__ patch_csc, const double* __restrict__ inv_s, const float* __restrict__ inv_s_f,
          const float* __restrict__ mu_f, const float* __restrict__ cent_f, int n_lines, int line_len, size_t ld,
          int n_strips, int pos0, int pos1, __half* __restrict__ out_hi, __half* __restrict__ out_lo,
          double* __restrict__ sumsq_partial) {
  __shared__ __align__(16) float overlay[2][kStripW];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int strip = blockIdx.x % n_strips;
  const int line0 = (blockIdx.x / n_strips) * kLinesPerCta;
  const int line1 = min(n_lines, line0 + kLinesPerCta);
  const int base = pos0 + strip * kStripW;
  const int p0 = base + tid * kElemsPerThread;          // first position of this thread
  const bool active = p0 < pos1;                         // pos1 is a multiple of 8 or == ld (also a multiple of 8)
  // per-position factors: P multiplies the line scalar, Q is added (cell-major only); 0 beyond the line (zero pad)
  float P[kElemsPerThread], Q[kElemsPerThread];
  bool tail = false;
#pragma unroll
  for (int q = 0; q < kElemsPerThread; ++q) {
    const int pos = p0 + q;
    const bool ok = pos < line_len;
    tail |= !ok;
    if (CELL_MAJOR) {
      P[q] = ok ? mu_f[pos] : 0.f;
      Q[q] = ok ? -cent_f[pos] : 0.f;
    } else {
      P[q] = ok ? inv_s_f[pos] : 0.f;
      Q[q] = 0.f;
    }
  }
  for (int i = tid; i < kStripW; i += kDenseThreads) { overlay[0][i] = 0.f; overlay[1][i] = 0.f; }

  // strip boundaries of this CTA's lines inside the sparse arrays
  __shared__ uint32_t seg[kLinesPerCta][2];
  if (tid < 2 * (line1 - line0)) {
    const int l = tid >> 1;
    seg[l][tid & 1] = off[(size_t)(line0 + l) * (n_strips + 1) + strip + (tid & 1)];
  }
  __syncthreads();
  // patch of `line` that this thread carries into the overlay (first one in registers, the rest in the slow loop)
  auto load_patch = [&](int line, uint32_t& t, uint32_t& t_end, uint32_t& pos, float& v) {
    t = seg[line - line0][0] + tid;
    t_end = seg[line - line0][1];
    pos = 0; v = 0.f;
    if (t < t_end) {
      pos = idx[t];
      v = CELL_MAJOR ? (float)(z_csr[t] * inv_s[line]) : patch_csc[t];
    }
  };
  uint32_t nt = 0, nt_end = 0, npos = 0;
  float nv = 0.f;
  if (line0 < line1) load_patch(line0, nt, nt_end, npos, nv);
  int buf = 0;
  for (int line = line0; line < line1; ++line, buf ^= 1) {
    // scatter this line's patches (loaded during the previous iteration)
    if (nt < nt_end) {
      overlay[buf][slot_of(npos - base)] = nv;
      for (uint32_t t = nt + kDenseThreads; t < nt_end; t += kDenseThreads)
        overlay[buf][slot_of(idx[t] - base)] = CELL_MAJOR ? (float)(z_csr[t] * inv_s[line]) : patch_csc[t];
    }
    const float a = CELL_MAJOR ? -inv_s_f[line] : -mu_f[line];
    const float c = CELL_MAJOR ? 0.f : -cent_f[line];
    if (line + 1 < line1) load_patch(line + 1, nt, nt_end, npos, nv); else nt_end = 0;
    __syncthreads();
    float sq = 0.f;
    if (active) {
      float4* ov0 = reinterpret_cast<float4*>(&overlay[buf][tid * 4]);
      float4* ov1 = reinterpret_cast<float4*>(&overlay[buf][kStripW / 2 + tid * 4]);
      const float4 d0 = *ov0, d1 = *ov1;
      const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
      if ((d0.x != 0.f) | (d0.y != 0.f) | (d0.z != 0.f) | (d0.w != 0.f)) *ov0 = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((d1.x != 0.f) | (d1.y != 0.f) | (d1.z != 0.f) | (d1.w != 0.f)) *ov1 = make_float4(0.f, 0.f, 0.f, 0.f);
      float f[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) f[q] = fmaf(a, P[q], CELL_MAJOR ? Q[q] : c) + d[q];
      if (!CELL_MAJOR && tail) {
#pragma unroll
        for (int q = 0; q < 8; ++q) if (p0 + q >= line_len) f[q] = 0.f;
      }
      __align__(16) __half2 h2[4];
      __align__(16) __half2 l2[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        h2[q] = __floats2half2_rn(f[2 * q], f[2 * q + 1]);
        const float2 hb = __half22float2(h2[q]);
        if (WITH_LO) {
          l2[q] = __floats2half2_rn(f[2 * q] - hb.x, f[2 * q + 1] - hb.y);
          if (WITH_SQ) {
            const float2 lb = __half22float2(l2[q]);
            const float e0 = hb.x + lb.x, e1 = hb.y + lb.y;
            sq = fmaf(e0, e0, fmaf(e1, e1, sq));
          }
        } else if (WITH_SQ) {
          sq = fmaf(hb.x, hb.x, fmaf(hb.y, hb.y, sq));
        }
      }
      const size_t o = (size_t)line * ld + (size_t)p0;
      *reinterpret_cast<uint4*>(out_hi + o) = *reinterpret_cast<const uint4*>(h2);
      if (WITH_LO) *reinterpret_cast<uint4*>(out_lo + o) = *reinterpret_cast<const uint4*>(l2);
    }
    if (WITH_SQ) {
      // squares of binary16 values are exact in Float32; the 8-term thread sums and the 32-lane tree round to
      // nearest (unbiased, ~1e-7 relative); the per-warp partials are summed in Float64 in a fixed order
#pragma unroll
      for (int o2 = 16; o2; o2 >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o2);
      if (lane == 0) sumsq_partial[((size_t)line * n_strips + strip) * (kDenseThreads / 32) + warp] = (double)sq;
    }
  }
}

// G[i][i] = scale * sum over the line's partial sums of squares (fixed order: deterministic)
__global__ void k_set_diagonal(float* __restrict__ G, int n, int n_parts, const double* __restrict__ partial, double scale) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0;
  for (int q = 0; q < n_parts; ++q) s += partial[(size_t)i * n_parts + q];
  G[(size_t)i * n + i] = (float)(s * scale);
}

// number of sum-of-squares partials per line for a range of n_pos positions
int densify_strips(size_t n_pos) { return (int)((n_pos + kStripW - 1) / kStripW) * (kDenseThreads / 32); }

void set_gram_diagonal(float* G, int n, int n_parts, const double* partial, double scale, cudaStream_t st) {
  count_launches(1);
  k_set_diagonal<<<(n + 255) / 256, 256, 0, st>>>(G, n, n_parts, partial, scale);
  SCL_CUDA(cudaGetLastError());
}

void densify(const SpMat& A, const NormStats& S, int layout, size_t ld, __half* out_hi, __half* out_lo,
             cudaStream_t st, double* sumsq_partial, long long pos0, long long pos1) {
  count_launches(2);
  const bool cell_major = layout == 1;
  const int n_lines = cell_major ? A.N : A.M;
  const int line_len = cell_major ? A.M : A.N;
  SCL_REQUIRE(ld % 8 == 0 && ld >= (size_t)line_len, "leading dimension must be a multiple of 8 and >= line length");
  if (pos1 < 0) { pos0 = 0; pos1 = (long long)ld; }
  SCL_REQUIRE(pos0 % 8 == 0 && pos0 >= 0 && pos1 <= (long long)ld && pos1 > pos0 && (pos1 % 8 == 0 || pos1 == (long long)ld), "bad densify range");
  const int n_strips = (int)((pos1 - pos0 + kStripW - 1) / kStripW);
  const uint32_t* ptr = cell_major ? A.rowptr.p : A.colptr.p;
  const uint32_t* idx = cell_major ? A.colidx.p : A.rowval.p;
  Tmp<uint32_t> off((size_t)n_lines * (n_strips + 1), st);
  {
    const long long total = (long long)n_lines * (n_strips + 1);
    const int grid = (int)std::min<long long>((total + 255) / 256, 148LL * 16);
    k_strip_offsets<<<grid, 256, 0, st>>>(ptr, idx, n_lines, n_strips, (int)pos0, off.p);
  }
  const long long ctas = (long long)n_strips * ((n_lines + kLinesPerCta - 1) / kLinesPerCta);
  SCL_REQUIRE(ctas < (1LL << 31), "densify grid too large");
#define SCL_LAUNCH_DENSIFY(CM, LO, SQ)                                                                            \
  k_densify<CM, LO, SQ><<<(unsigned)ctas, kDenseThreads, 0, st>>>(off.p, idx, S.z_csr.p, S.patch_csc.p, S.inv_s.p,  \
                                                                 S.inv_s_f.p, S.mu_f.p, S.cent_f.p, n_lines,     \
                                                                 line_len, ld, n_strips, (int)pos0, (int)pos1,   \
                                                                 out_hi, out_lo, sumsq_partial)
  const bool lo = out_lo != nullptr, sq = sumsq_partial != nullptr;
  if (cell_major) {
    if (lo) { if (sq) SCL_LAUNCH_DENSIFY(true, true, true); else SCL_LAUNCH_DENSIFY(true, true, false); }
    else    { if (sq) SCL_LAUNCH_DENSIFY(true, false, true); else SCL_LAUNCH_DENSIFY(true, false, false); }
  } else {
    if (lo) { if (sq) SCL_LAUNCH_DENSIFY(false, true, true); else SCL_LAUNCH_DENSIFY(false, true, false); }
    else    { if (sq) SCL_LAUNCH_DENSIFY(false, false, true); else SCL_LAUNCH_DENSIFY(false, false, false); }
  }
#undef SCL_LAUNCH_DENSIFY
  SCL_CUDA(cudaGetLastError());
}

}  // namespace scl
