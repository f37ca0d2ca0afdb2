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

// ---- per-gene mean / corrected std of y (CSC, one warp per gene) --------------------------
// y = log1p(x / r_i) is evaluated once per non-zero and kept (Float64, CSC order) for k_gene_center.  Four
// independent entries per lane and iteration keep enough loads in flight to hide the gather latency; the
// lane-strided order and the shuffle tree are fixed, so the sums are deterministic.
static constexpr int kStatUnroll = 4;

__global__ void __launch_bounds__(256) k_gene_stats(const uint32_t* __restrict__ colptr,
                                                    const uint32_t* __restrict__ rowval,
                                                    const float* __restrict__ val, const double* __restrict__ tgc,
                                                    int N, int M, double* __restrict__ y_csc, double* __restrict__ ybar,
                                                    double* __restrict__ sigma, double* __restrict__ mu,
                                                    float* __restrict__ mu_f, float* __restrict__ inv_sigma_f) {
  const int lane = threadIdx.x & 31;
  int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int per_grid = (gridDim.x * blockDim.x) >> 5;
  for (int jj = j; jj < M; jj += per_grid) {
    const int j = M - 1 - jj;   // genes arrive sorted by mean expression (:224): densest columns first, short tail
    const uint32_t b = colptr[j], e = colptr[j + 1];
    double s1 = 0, s2 = 0;
    for (uint32_t t = b + lane; t < e; t += 32 * kStatUnroll) {
      float v[kStatUnroll];
      double r[kStatUnroll];
#pragma unroll
      for (int u = 0; u < kStatUnroll; ++u) {
        const uint32_t tt = t + 32 * u;
        const bool ok = tt < e;
        v[u] = ok ? val[tt] : 0.f;
        r[u] = tgc[ok ? rowval[tt] : 0];
      }
#pragma unroll
      for (int u = 0; u < kStatUnroll; ++u) {
        const uint32_t tt = t + 32 * u;
        if (tt < e) {
          const double y = log1p((double)v[u] * (1.0 / r[u]));
          y_csc[tt] = y;
          s1 += y;
          s2 = fma(y, y, s2);
        }
      }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      // sum over all N cells of (y - m)^2 with y = 0 at the implicit zeros: sum y^2 - N m^2
      const double m = s1 / (double)N;
      double var = (s2 - (double)N * m * m) / (double)(N - 1);
      if (var < 0) var = 0;
      const double sd = sqrt(var);
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
__global__ void __launch_bounds__(256) k_cell_l2(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ colidx,
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
    const uint32_t b = rowptr[row], e = rowptr[row + 1];
    double a = 0, bb = 0;
    for (uint32_t t = b + lane; t < e; t += 32 * kStatUnroll) {
      float v[kStatUnroll];
      double sg[kStatUnroll], m[kStatUnroll];
#pragma unroll
      for (int u = 0; u < kStatUnroll; ++u) {
        const uint32_t tt = t + 32 * u;
        const bool ok = tt < e;
        const uint32_t c = ok ? colidx[tt] : 0;
        v[u] = ok ? rval[tt] : 0.f;
        sg[u] = sigma[c];
        m[u] = mu[c];
      }
#pragma unroll
      for (int u = 0; u < kStatUnroll; ++u) {
        const uint32_t tt = t + 32 * u;
        if (tt < e) {
          const double z = log1p((double)v[u] * inv_r) / sg[u];
          z_csr[tt] = z;
          a = fma(z, z, a);
          bb = fma(z, m[u], bb);
        }
      }
    }
    a = warp_sum(a);
    bb = warp_sum(bb);
    if (lane == 0) l2[row] = sqrt(a - 2.0 * bb + mu2);
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

// ---- per-gene centre after cell scaling (CSC, one warp per gene) ---------------------------
// Also emits the gene-major writer's sparse patch - the final value (z_ij - mu_j)/s_i - c_j of every stored entry,
// evaluated in Float64 and rounded once to Float32 (CSC order) - and the gene's exact sum of squares over all cells
// (the Gram diagonal when genes are the Gram side): background in closed form plus a correction per stored entry,
//   sum_i w^2 = sum_i bg_i^2 + sum_nz (u^2 - 2 u bg),   u = z/s_i,  bg = mu_j/s_i + c_j,  w = u - bg.
__global__ void __launch_bounds__(256) k_gene_center(const uint32_t* __restrict__ colptr,
                                                     const uint32_t* __restrict__ rowval,
                                                     const double* __restrict__ y_csc,
                                                     const double* __restrict__ sigma, const double* __restrict__ mu,
                                                     const double* __restrict__ inv_s,
                                                     const double* __restrict__ scalars, int N, int M,
                                                     float* __restrict__ patch_csc, double* __restrict__ cent,
                                                     float* __restrict__ cent_f, double* __restrict__ sumsq_gene) {
  const int lane = threadIdx.x & 31;
  int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int per_grid = (gridDim.x * blockDim.x) >> 5;
  const double sum_inv_s = scalars[2], sum_inv_s2 = scalars[3];
  for (int jj = j; jj < M; jj += per_grid) {
    const int j = M - 1 - jj;   // densest columns first
    const uint32_t b = colptr[j], e = colptr[j + 1];
    const double sd = sigma[j], m = mu[j];
    double su = 0, suu = 0, sui = 0;    // sum u, sum u^2, sum u/s_i
    for (uint32_t t = b + lane; t < e; t += 32 * kStatUnroll) {
      double y[kStatUnroll], is[kStatUnroll];
#pragma unroll
      for (int u = 0; u < kStatUnroll; ++u) {
        const uint32_t tt = t + 32 * u;
        const bool ok = tt < e;
        y[u] = ok ? y_csc[tt] : 0.0;
        is[u] = inv_s[ok ? rowval[tt] : 0];
      }
#pragma unroll
      for (int u = 0; u < kStatUnroll; ++u) {
        const double uu = y[u] / sd * is[u];      // 0 for the padding lanes (y = 0)
        su += uu;
        suu = fma(uu, uu, suu);
        sui = fma(uu, is[u], sui);
      }
    }
    su = warp_sum(su);
    suu = warp_sum(suu);
    sui = warp_sum(sui);
    const double c = (su - m * sum_inv_s) / (double)N;
    if (lane == 0) {
      cent[j] = c;
      cent_f[j] = (float)c;
      sumsq_gene[j] = m * m * sum_inv_s2 + 2.0 * m * c * sum_inv_s + (double)N * c * c + suu - 2.0 * m * sui - 2.0 * c * su;
    }
    for (uint32_t t = b + lane; t < e; t += 32 * kStatUnroll) {
      double y[kStatUnroll], is[kStatUnroll];
#pragma unroll
      for (int u = 0; u < kStatUnroll; ++u) {
        const uint32_t tt = t + 32 * u;
        const bool ok = tt < e;
        y[u] = ok ? y_csc[tt] : 0.0;
        is[u] = inv_s[ok ? rowval[tt] : 0];
      }
#pragma unroll
      for (int u = 0; u < kStatUnroll; ++u) {
        const uint32_t tt = t + 32 * u;
        if (tt < e) patch_csc[tt] = (float)((y[u] / sd - m) * is[u] - c);
      }
    }
  }
}

// out[0] = sum a_i b_i (one block, deterministic)
__global__ void __launch_bounds__(1024) k_dot(const double* __restrict__ a, const double* __restrict__ b, int n,
                                              double* __restrict__ out) {
  __shared__ double red[32];
  double s = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += a[i] * b[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) *out = s;
}

// ---- cell-side finishing pass (CSR, one warp per cell) --------------------------------------
// Emits the cell-major writer's sparse patch - the final value (z_ij - mu_j)/s_i - c_j of every stored entry, Float64
// rounded once to Float32, CSR order - and the exact sum of squares of every cell's normalised row (the Gram diagonal
// when cells are the Gram side): sum_j (mu_j/s_i + c_j)^2 in closed form from |mu|^2, mu.c, |c|^2, plus w^2 - bg^2 at
// the stored entries.
__global__ void __launch_bounds__(256) k_cell_finish(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ colidx,
                                                     const double* __restrict__ z_csr, const double* __restrict__ mu,
                                                     const double* __restrict__ cent, const double* __restrict__ inv_s,
                                                     const double* __restrict__ scalars, int N,
                                                     float* __restrict__ patch_csr, double* __restrict__ sumsq_cell) {
  const int lane = threadIdx.x & 31;
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nrows_per_grid = (gridDim.x * blockDim.x) >> 5;
  const double mu2 = scalars[0], muc = scalars[4], c2 = scalars[5];
  for (; row < N; row += nrows_per_grid) {
    const double is = inv_s[row];
    const uint32_t b = rowptr[row], e = rowptr[row + 1];
    double dq = 0;
    for (uint32_t t = b + lane; t < e; t += 32 * kStatUnroll) {
      double z[kStatUnroll], m[kStatUnroll], cc[kStatUnroll];
#pragma unroll
      for (int u = 0; u < kStatUnroll; ++u) {
        const uint32_t tt = t + 32 * u;
        const bool ok = tt < e;
        const uint32_t c = ok ? colidx[tt] : 0;
        z[u] = ok ? z_csr[tt] : 0.0;
        m[u] = mu[c];
        cc[u] = cent[c];
      }
#pragma unroll
      for (int u = 0; u < kStatUnroll; ++u) {
        const uint32_t tt = t + 32 * u;
        if (tt < e) {
          const double bg = is * m[u] + cc[u], w = (z[u] - m[u]) * is - cc[u];
          patch_csr[tt] = (float)w;
          dq += w * w - bg * bg;
        }
      }
    }
    dq = warp_sum(dq);
    if (lane == 0) sumsq_cell[row] = is * is * mu2 + 2.0 * is * muc + c2 + dq;
  }
}

void compute_norm_stats(const SpMat& A, NormStats& S, cudaStream_t st) {
  const int N = A.N, M = A.M;
  SCL_REQUIRE(N > 1 && M > 0 && A.nnz > 0, "matrix too small to normalise");
  S.tgc.ensure(N); S.l2.ensure(N); S.inv_s.ensure(N); S.inv_s_f.ensure(N);
  S.ybar.ensure(M); S.sigma.ensure(M); S.mu.ensure(M); S.cent.ensure(M);
  S.mu_f.ensure(M); S.cent_f.ensure(M); S.inv_sigma_f.ensure(M);
  S.scalars.ensure(8);
  S.y_csc.ensure(A.nnz); S.z_csr.ensure(A.nnz); S.patch_csc.ensure(A.nnz);
  S.sumsq_gene.ensure(M); S.sumsq_cell.ensure(N); S.patch_csr.ensure(A.nnz);
  count_launches(12);
  const int wgrid = min((N + 7) / 8, 148 * 8);
  k_row_sum<<<wgrid, 256, 0, st>>>(A.rowptr.p, A.rval.p, N, S.tgc.p);
  const int ggrid = min((M + 7) / 8, 148 * 8);
  k_gene_stats<<<ggrid, 256, 0, st>>>(A.colptr.p, A.rowval.p, A.val.p, S.tgc.p, N, M, S.y_csc.p, S.ybar.p,
                                                  S.sigma.p, S.mu.p, S.mu_f.p, S.inv_sigma_f.p);
  k_reduce<<<1, 1024, 0, st>>>(S.mu.p, M, 0, S.scalars.p + 0);
  k_cell_l2<<<wgrid, 256, 0, st>>>(A.rowptr.p, A.colidx.p, A.rval.p, S.tgc.p, S.sigma.p, S.mu.p, S.scalars.p, N,
                                   S.z_csr.p, S.l2.p);
  k_reduce<<<1, 1024, 0, st>>>(S.l2.p, N, 1, S.scalars.p + 1);
  k_inv_s<<<(N + 255) / 256, 256, 0, st>>>(S.l2.p, S.scalars.p, N, S.inv_s.p, S.inv_s_f.p);
  k_reduce<<<1, 1024, 0, st>>>(S.inv_s.p, N, 2, S.scalars.p + 2);
  k_reduce<<<1, 1024, 0, st>>>(S.inv_s.p, N, 0, S.scalars.p + 3);
  k_gene_center<<<ggrid, 256, 0, st>>>(A.colptr.p, A.rowval.p, S.y_csc.p, S.sigma.p, S.mu.p, S.inv_s.p,
                                                   S.scalars.p, N, M, S.patch_csc.p, S.cent.p, S.cent_f.p,
                                                   S.sumsq_gene.p);
  k_dot<<<1, 1024, 0, st>>>(S.mu.p, S.cent.p, M, S.scalars.p + 4);
  k_reduce<<<1, 1024, 0, st>>>(S.cent.p, M, 0, S.scalars.p + 5);
  k_cell_finish<<<wgrid, 256, 0, st>>>(A.rowptr.p, A.colidx.p, S.z_csr.p, S.mu.p, S.cent.p, S.inv_s.p, S.scalars.p, N,
                                       S.patch_csr.p, S.sumsq_cell.p);
  SCL_CUDA(cudaGetLastError());
}

// ---- fused densify + normalise writer -----------------------------------------------------
// out_ij = (z_ij - mu_j)/s_i - c_j: a rank-structured background (z = 0) overridden at the stored entries.  A CTA owns a strip of
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

template <bool CELL_MAJOR, bool WITH_LO>
__global__ void __launch_bounds__(kDenseThreads)
k_densify(const uint32_t* __restrict__ off, const uint32_t* __restrict__ idx, const float* __restrict__ patch,
          const float* __restrict__ inv_s_f, const float* __restrict__ mu_f,
          const float* __restrict__ cent_f, int n_lines, int line_len, size_t ld,
          int n_strips, int pos0, int pos1, __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
  // two lines per barrier, double buffered: overlay[set][line of the pair][slot]
  __shared__ __align__(16) float overlay[2][2][kStripW];
  __shared__ uint32_t seg[kLinesPerCta][2];     // strip boundaries of this CTA's lines inside the sparse arrays
  __shared__ float line_a[kLinesPerCta], line_c[kLinesPerCta];
  const int tid = threadIdx.x;
  const int strip = blockIdx.x % n_strips;
  const int line0 = (blockIdx.x / n_strips) * kLinesPerCta;
  const int line1 = min(n_lines, line0 + kLinesPerCta);
  const int n_my = line1 - line0;
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
  // overlay slots hold the final value of a stored entry, or NaN ("no stored entry here: background")
  const float kNone = __int_as_float(0x7fc00000);
  for (int i = tid; i < 4 * kStripW; i += kDenseThreads) (&overlay[0][0][0])[i] = kNone;
  if (tid < 2 * n_my) {
    const int l = tid >> 1;
    seg[l][tid & 1] = off[(size_t)(line0 + l) * (n_strips + 1) + strip + (tid & 1)];
  } else if (tid >= 64 && tid < 64 + n_my) {
    const int l = tid - 64;
    line_a[l] = CELL_MAJOR ? -inv_s_f[line0 + l] : -mu_f[line0 + l];
    line_c[l] = CELL_MAJOR ? 0.f : -cent_f[line0 + l];
  }
  __syncthreads();

  // first stored entry of line l (CTA-relative) this thread carries into the overlay; the rest go the slow way
  auto load_patch = [&](int l, uint32_t& t, uint32_t& t_end, uint32_t& pos, float& v) {
    t = 0; t_end = 0; pos = 0; v = 0.f;
    if (l < n_my) {
      t = seg[l][0] + tid;
      t_end = seg[l][1];
      if (t < t_end) {
        pos = idx[t];
        v = patch[t];
      }
    }
  };
  auto scatter = [&](float* ov, uint32_t t, uint32_t t_end, uint32_t pos, float v) {
    if (t < t_end) {
      ov[slot_of(pos - base)] = v;
      for (uint32_t u = t + kDenseThreads; u < t_end; u += kDenseThreads) ov[slot_of(idx[u] - base)] = patch[u];
    }
  };
  auto emit = [&](float* ov, int l, __half* dst_hi, __half* dst_lo) {
    float4* ov0 = reinterpret_cast<float4*>(ov + tid * 4);
    float4* ov1 = reinterpret_cast<float4*>(ov + kStripW / 2 + tid * 4);
    const float4 d0 = *ov0, d1 = *ov1;
    const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
    if ((d0.x == d0.x) | (d0.y == d0.y) | (d0.z == d0.z) | (d0.w == d0.w)) *ov0 = make_float4(kNone, kNone, kNone, kNone);
    if ((d1.x == d1.x) | (d1.y == d1.y) | (d1.z == d1.z) | (d1.w == d1.w)) *ov1 = make_float4(kNone, kNone, kNone, kNone);
    const float a = line_a[l], c = line_c[l];
    float f[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) f[q] = d[q] == d[q] ? d[q] : fmaf(a, P[q], CELL_MAJOR ? Q[q] : c);
    if (!CELL_MAJOR && tail) {
#pragma unroll
      for (int q = 0; q < 8; ++q) if (p0 + q >= line_len) f[q] = 0.f;
    }
    __align__(16) __half2 h2[4];
    __align__(16) __half2 l2[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      h2[q] = __floats2half2_rn(f[2 * q], f[2 * q + 1]);
      if (WITH_LO) {
        const float2 hb = __half22float2(h2[q]);
        l2[q] = __floats2half2_rn(f[2 * q] - hb.x, f[2 * q + 1] - hb.y);
      }
    }
    *reinterpret_cast<uint4*>(dst_hi) = *reinterpret_cast<const uint4*>(h2);
    if (WITH_LO) *reinterpret_cast<uint4*>(dst_lo) = *reinterpret_cast<const uint4*>(l2);
  };

  uint32_t t0, e0, q0, t1, e1, q1;
  float v0, v1;
  load_patch(0, t0, e0, q0, v0);
  load_patch(1, t1, e1, q1, v1);
  __half* row_hi = out_hi + (size_t)line0 * ld + (size_t)p0;     // this thread's 16 bytes of the current line
  __half* row_lo = WITH_LO ? out_lo + (size_t)line0 * ld + (size_t)p0 : nullptr;
  int set = 0;
  for (int l = 0; l < n_my; l += 2, set ^= 1) {
    // scatter this pair's patches (loaded during the previous iteration), start the next pair's loads
    scatter(overlay[set][0], t0, e0, q0, v0);
    scatter(overlay[set][1], t1, e1, q1, v1);
    load_patch(l + 2, t0, e0, q0, v0);
    load_patch(l + 3, t1, e1, q1, v1);
    __syncthreads();
    if (active) {
      emit(overlay[set][0], l, row_hi, row_lo);
      if (l + 1 < n_my) emit(overlay[set][1], l + 1, row_hi + ld, WITH_LO ? row_lo + ld : nullptr);
    }
    row_hi += 2 * ld;
    if (WITH_LO) row_lo += 2 * ld;
  }
}

// G[i][i] = scale * sumsq[i]
__global__ void k_set_diagonal(float* __restrict__ G, int n, const double* __restrict__ sumsq, double scale) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) G[(size_t)i * n + i] = (float)(sumsq[i] * scale);
}

// Exact (Float64) sums of squares of the normalised matrix's lines on the Gram side: the Gram diagonal.  The tensor
// core sees the binary16 roundings of these values; their squares differ from the exact ones by an unbiased ~4e-6
// relative per diagonal entry (K >= 1e4 terms), which moves no eigenvalue by more than ~1e-7.
const double* gram_diagonal(const SpMat& A, NormStats& S, bool gene_side, cudaStream_t st) {
  (void)A; (void)st;
  return gene_side ? S.sumsq_gene.p : S.sumsq_cell.p;
}

void set_gram_diagonal(float* G, int n, const double* sumsq, double scale, cudaStream_t st) {
  count_launches(1);
  k_set_diagonal<<<(n + 255) / 256, 256, 0, st>>>(G, n, sumsq, scale);
  SCL_CUDA(cudaGetLastError());
}

void densify(const SpMat& A, const NormStats& S, int layout, size_t ld, __half* out_hi, __half* out_lo,
             cudaStream_t st, long long pos0, long long pos1) {
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
#define SCL_LAUNCH_DENSIFY(CM, LO)                                                                               \
  k_densify<CM, LO><<<(unsigned)ctas, kDenseThreads, 0, st>>>(off.p, idx, cell_major ? S.patch_csr.p : S.patch_csc.p, \
                                                             S.inv_s_f.p, S.mu_f.p, S.cent_f.p, n_lines, line_len, \
                                                             ld, n_strips, (int)pos0, (int)pos1, out_hi, out_lo)
  if (cell_major) {
    if (out_lo) SCL_LAUNCH_DENSIFY(true, true); else SCL_LAUNCH_DENSIFY(true, false);
  } else {
    if (out_lo) SCL_LAUNCH_DENSIFY(false, true); else SCL_LAUNCH_DENSIFY(false, false);
  }
#undef SCL_LAUNCH_DENSIFY
  SCL_CUDA(cudaGetLastError());
}

}  // namespace scl
