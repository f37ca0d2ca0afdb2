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
#include <cstdlib>
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

// ---- log1p on [0, 1] -------------------------------------------------------------------------
// u = x_ij / r_i lies in (0, 1] and is almost always tiny (a count over a cell's total).  log1p(u) = 2 atanh(s),
// s = u / (2 + u): for u <= 1/16, s^2 <= 9.2e-4 and the odd series truncated after s^13 is exact to < 1e-22
// relative, so the result carries the rounding of the division and of the last FMA only (~1 ulp).  A third of the
// Float64 instructions of the library log1p; the statistics passes re-evaluate it instead of storing y per entry.
__device__ __forceinline__ double log1p_unit(double u) {
  if (!(fabs(u) <= 0.0625)) return log1p(u);
  const double s = u / (2.0 + u), w = s * s;
  double p = 1.0 / 13.0;
  p = fma(p, w, 1.0 / 11.0);
  p = fma(p, w, 1.0 / 9.0);
  p = fma(p, w, 1.0 / 7.0);
  p = fma(p, w, 1.0 / 5.0);
  p = fma(p, w, 1.0 / 3.0);
  const double s2 = s + s;
  return fma(s2 * w, p, s2);
}

// ---- per-cell total counts (CSR, one warp per cell) -------------------------------------
__global__ void k_row_sum(const uint32_t* __restrict__ rowptr, const float* __restrict__ rval, int N,
                          double* __restrict__ tgc, double2* __restrict__ cell_par) {
  const int lane = threadIdx.x & 31;
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nrows_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (; row < N; row += nrows_per_grid) {
    double s = 0;
    for (uint32_t t = rowptr[row] + lane; t < rowptr[row + 1]; t += 32) s += (double)rval[t];
    s = warp_sum(s);
    if (lane == 0) {
      tgc[row] = s;
      cell_par[row] = make_double2(1.0 / s, 0.0);
    }
  }
}

// ---- line passes ----------------------------------------------------------------------------
// Every statistics pass walks the lines of one orientation (genes: CSC, cells: CSR), gathers the other side's
// parameters per stored entry (packed so that one 32-byte sector holds everything a pass needs: cell_par[i] =
// {1/r_i, 1/s_i}, gene_par[j] = {1/sigma_j, mu_j, c_j, -}) and reduces a few Float64 sums per line.  Line lengths
// are heavy-tailed (a gene expressed everywhere holds N entries, the median gene a few hundred), so a line of up
// to kHeavyLine entries is reduced by one warp and a longer one by the whole CTA; lines are dealt to CTAs round-robin
// (densest genes first when the caller asks for reversed order).  Lane/thread-strided partial sums and fixed
// shuffle / shared-memory trees: the sums depend on the line's length only, never on scheduling (deterministic).
static constexpr int kStatThreads = 256;

// One strided walk over the entries [b, e) of a line.  The (index, value) pairs of the next step are requested before
// the current step's gathers and Float64 arithmetic, so a warp always has a step's worth of streaming loads in flight
// (the gathers hit L1/L2; the stream is what has to cover the HBM latency).
template <class Op, int kStatUnroll>
__device__ __forceinline__ void line_span(const Op& op, const typename Op::Ctx& cx, uint32_t b, uint32_t e, int first,
                                          int stride, double (&acc)[Op::NACC]) {
  uint32_t t = b + first;
  uint32_t idx[kStatUnroll], nidx[kStatUnroll];
  float v[kStatUnroll], nv[kStatUnroll];
#pragma unroll
  for (int u = 0; u < kStatUnroll; ++u) {
    const uint32_t tt = t + stride * u;
    const bool ok = tt < e;
    idx[u] = ok ? op.idx[tt] : 0u;
    v[u] = ok ? op.val[tt] : 0.f;
  }
  while (t < e) {
    const uint32_t tn = t + stride * kStatUnroll;
#pragma unroll
    for (int u = 0; u < kStatUnroll; ++u) {
      const uint32_t tt = tn + stride * u;
      const bool ok = tt < e;
      nidx[u] = ok ? op.idx[tt] : 0u;
      nv[u] = ok ? op.val[tt] : 0.f;
    }
    typename Op::Ld ld[kStatUnroll];
#pragma unroll
    for (int u = 0; u < kStatUnroll; ++u) op.gather(idx[u], ld[u]);
#pragma unroll
    for (int u = 0; u < kStatUnroll; ++u) {
      const uint32_t tt = t + stride * u;
      if (tt < e) op.accum(tt, cx, v[u], ld[u], acc);
    }
#pragma unroll
    for (int u = 0; u < kStatUnroll; ++u) { idx[u] = nidx[u]; v[u] = nv[u]; }
    t = tn;
  }
}

template <class Op, int U, int MINB>
__global__ void __launch_bounds__(kStatThreads, MINB) k_lines(const Op op, const uint32_t* __restrict__ ptr, int n_lines,
                                                              int reversed, uint32_t kHeavyLine) {
  __shared__ double red[Op::NACC][kStatThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // lines of at most kHeavyLine entries: one warp each
  for (long long li = blockIdx.x + (long long)warp * gridDim.x; li < n_lines; li += (long long)gridDim.x * (kStatThreads / 32)) {
    const int line = reversed ? n_lines - 1 - (int)li : (int)li;
    const uint32_t b = ptr[line], e = ptr[line + 1];
    if (e - b > kHeavyLine) continue;
    const typename Op::Ctx cx = op.begin(line);
    double acc[Op::NACC];
#pragma unroll
    for (int a = 0; a < Op::NACC; ++a) acc[a] = 0;
    line_span<Op, U>(op, cx, b, e, lane, 32, acc);
#pragma unroll
    for (int a = 0; a < Op::NACC; ++a) acc[a] = warp_sum(acc[a]);
    if (lane == 0) op.finish(line, cx, acc);
  }
  // longer lines: the whole CTA (the loop and its branch are uniform over the CTA)
  for (long long li = blockIdx.x; li < n_lines; li += gridDim.x) {
    const int line = reversed ? n_lines - 1 - (int)li : (int)li;
    const uint32_t b = ptr[line], e = ptr[line + 1];
    if (e - b <= kHeavyLine) continue;
    const typename Op::Ctx cx = op.begin(line);
    double acc[Op::NACC];
#pragma unroll
    for (int a = 0; a < Op::NACC; ++a) acc[a] = 0;
    line_span<Op, U>(op, cx, b, e, tid, kStatThreads, acc);
#pragma unroll
    for (int a = 0; a < Op::NACC; ++a) acc[a] = warp_sum(acc[a]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
      for (int a = 0; a < Op::NACC; ++a) red[a][warp] = acc[a];
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int a = 0; a < Op::NACC; ++a) {
        double t = 0;
        for (int w = 0; w < kStatThreads / 32; ++w) t += red[a][w];
        acc[a] = t;
      }
      op.finish(line, cx, acc);
    }
  }
}

// per-gene mean / corrected std of y = log1p(x / r_i)  (:681-686)
struct GeneStatsOp {
  static constexpr int NACC = 2;
  const uint32_t* __restrict__ idx;   // rowval
  const float* __restrict__ val;
  const double2* __restrict__ cell_par;
  int N;
  double *ybar, *sigma, *mu;
  float *mu_f, *inv_sigma_f;
  double4* gene_par;
  struct Ctx {};
  struct Ld { double inv_r; };
  __device__ __forceinline__ Ctx begin(int) const { return Ctx(); }
  __device__ __forceinline__ void gather(uint32_t i, Ld& l) const { l.inv_r = cell_par[i].x; }
  __device__ __forceinline__ void accum(uint32_t, const Ctx&, float v, const Ld& l, double (&acc)[NACC]) const {
    const double y = log1p_unit((double)v * l.inv_r);
    acc[0] += y;
    acc[1] = fma(y, y, acc[1]);
  }
  __device__ __forceinline__ void finish(int j, const Ctx&, const double (&acc)[NACC]) const {
    // sum over all N cells of (y - m)^2 with y = 0 at the implicit zeros: sum y^2 - N m^2
    const double m = acc[0] / (double)N;
    double var = (acc[1] - (double)N * m * m) / (double)(N - 1);
    if (var < 0) var = 0;
    const double sd = sqrt(var);
    ybar[j] = m;
    sigma[j] = sd;
    mu[j] = m / sd;
    mu_f[j] = (float)(m / sd);
    inv_sigma_f[j] = (float)(1.0 / sd);
    gene_par[j] = make_double4(1.0 / sd, m / sd, 0.0, 0.0);
  }
};

// per-cell l2 norm after the mean shift, from the sparse entries only  (:688-689, :603)
struct CellL2Op {
  static constexpr int NACC = 2;
  const uint32_t* __restrict__ idx;   // colidx
  const float* __restrict__ val;      // rval
  const double2* __restrict__ cell_par;
  const double4* __restrict__ gene_par;
  const double* __restrict__ scalars;
  double* l2;
  struct Ctx { double inv_r; };
  struct Ld { double inv_sd, mu; };
  __device__ __forceinline__ Ctx begin(int row) const { return Ctx{cell_par[row].x}; }
  __device__ __forceinline__ void gather(uint32_t j, Ld& l) const {
    const double2 g = *reinterpret_cast<const double2*>(&gene_par[j]);
    l.inv_sd = g.x;
    l.mu = g.y;
  }
  __device__ __forceinline__ void accum(uint32_t, const Ctx& cx, float v, const Ld& l, double (&acc)[NACC]) const {
    const double z = log1p_unit((double)v * cx.inv_r) * l.inv_sd;
    acc[0] = fma(z, z, acc[0]);
    acc[1] = fma(z, l.mu, acc[1]);
  }
  __device__ __forceinline__ void finish(int row, const Ctx&, const double (&acc)[NACC]) const {
    l2[row] = sqrt(acc[0] - 2.0 * acc[1] + scalars[0]);
  }
};

// per-gene centre after cell scaling (:695) and the gene's exact sum of squares over all cells (the Gram diagonal
// when genes are the Gram side): background in closed form plus a correction per stored entry,
//   sum_i w^2 = sum_i bg_i^2 + sum_nz (u^2 - 2 u bg),   u = z/s_i,  bg = mu_j/s_i + c_j,  w = u - bg.
struct GeneCenterOp {
  static constexpr int NACC = 3;
  const uint32_t* __restrict__ idx;   // rowval
  const float* __restrict__ val;
  const double2* __restrict__ cell_par;
  const double* __restrict__ scalars;
  int N;
  double4* gene_par;
  double *cent, *sumsq_gene;
  float* cent_f;
  struct Ctx { double inv_sd, mu; };
  struct Ld { double2 c; };
  __device__ __forceinline__ Ctx begin(int j) const {
    const double2 g = *reinterpret_cast<const double2*>(&gene_par[j]);
    return Ctx{g.x, g.y};
  }
  __device__ __forceinline__ void gather(uint32_t i, Ld& l) const { l.c = cell_par[i]; }
  __device__ __forceinline__ void accum(uint32_t, const Ctx& cx, float v, const Ld& l, double (&acc)[NACC]) const {
    const double uu = log1p_unit((double)v * l.c.x) * cx.inv_sd * l.c.y;
    acc[0] += uu;
    acc[1] = fma(uu, uu, acc[1]);
    acc[2] = fma(uu, l.c.y, acc[2]);
  }
  __device__ __forceinline__ void finish(int j, const Ctx& cx, const double (&acc)[NACC]) const {
    const double sum_inv_s = scalars[2], sum_inv_s2 = scalars[3], m = cx.mu;
    const double su = acc[0], suu = acc[1], sui = acc[2];
    const double c = (su - m * sum_inv_s) / (double)N;
    cent[j] = c;
    cent_f[j] = (float)c;
    gene_par[j].z = c;
    sumsq_gene[j] = m * m * sum_inv_s2 + 2.0 * m * c * sum_inv_s + (double)N * c * c + suu - 2.0 * m * sui - 2.0 * c * su;
  }
};

// gene-major writer's sparse patch: the final value (z_ij - mu_j)/s_i - c_j of every stored entry, evaluated in
// Float64 and rounded once to Float32 (CSC order)
struct GenePatchOp {
  static constexpr int NACC = 1;
  const uint32_t* __restrict__ idx;   // rowval
  const float* __restrict__ val;
  const double2* __restrict__ cell_par;
  const double4* __restrict__ gene_par;
  float* patch_csc;
  struct Ctx { double inv_sd, mu, c; };
  struct Ld { double2 c; };
  __device__ __forceinline__ Ctx begin(int j) const {
    const double4 g = gene_par[j];
    return Ctx{g.x, g.y, g.z};
  }
  __device__ __forceinline__ void gather(uint32_t i, Ld& l) const { l.c = cell_par[i]; }
  __device__ __forceinline__ void accum(uint32_t t, const Ctx& cx, float v, const Ld& l, double (&)[NACC]) const {
    patch_csc[t] = (float)((log1p_unit((double)v * l.c.x) * cx.inv_sd - cx.mu) * l.c.y - cx.c);
  }
  __device__ __forceinline__ void finish(int, const Ctx&, const double (&)[NACC]) const {}
};

// cell-major writer's sparse patch (CSR order) and the exact sum of squares of every cell's normalised row (the Gram
// diagonal when cells are the Gram side): sum_j (mu_j/s_i + c_j)^2 in closed form from |mu|^2, mu.c, |c|^2, plus
// w^2 - bg^2 at the stored entries.
struct CellFinishOp {
  static constexpr int NACC = 1;
  const uint32_t* __restrict__ idx;   // colidx
  const float* __restrict__ val;      // rval
  const double2* __restrict__ cell_par;
  const double4* __restrict__ gene_par;
  const double* __restrict__ scalars;
  float* patch_csr;
  double* sumsq_cell;
  struct Ctx { double inv_r, is; };
  struct Ld { double4 g; };
  __device__ __forceinline__ Ctx begin(int row) const {
    const double2 c = cell_par[row];
    return Ctx{c.x, c.y};
  }
  __device__ __forceinline__ void gather(uint32_t j, Ld& l) const { l.g = gene_par[j]; }
  __device__ __forceinline__ void accum(uint32_t t, const Ctx& cx, float v, const Ld& l, double (&acc)[NACC]) const {
    const double z = log1p_unit((double)v * cx.inv_r) * l.g.x;
    const double bg = cx.is * l.g.y + l.g.z, w = (z - l.g.y) * cx.is - l.g.z;
    patch_csr[t] = (float)w;
    acc[0] += w * w - bg * bg;
  }
  __device__ __forceinline__ void finish(int row, const Ctx& cx, const double (&acc)[NACC]) const {
    const double mu2 = scalars[0], muc = scalars[4], c2 = scalars[5];
    sumsq_cell[row] = cx.is * cx.is * mu2 + 2.0 * cx.is * muc + c2 + acc[0];
  }
};

// ---- small deterministic reductions over the per-line vectors ---------------------------------
// Each block reduces a strided slice with a fixed tree and writes its partial sums; the block that finishes last
// adds the partials in block order and derives the scalars, so the result does not depend on scheduling.
//   MODE 0: a = mu            -> scalars[0] = |mu|^2
//   MODE 1: a = l             -> scalars[1] = mean(l), [2] = sum(1/s), [3] = sum(1/s^2),  s_i = l_i / mean(l)
//   MODE 2: a = mu, b = cent  -> scalars[4] = mu.c, [5] = |c|^2
static constexpr int kRedBlocks = 64;
template <int MODE>
__global__ void __launch_bounds__(256) k_reduce_vec(const double* __restrict__ a, const double* __restrict__ b, int n,
                                                    double* __restrict__ partial /* [3][kRedBlocks] */,
                                                    unsigned int* __restrict__ counter, double* __restrict__ scalars) {
  __shared__ double red[3][8];
  double s0 = 0, s1 = 0, s2 = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double x = a[i];
    if (MODE == 0) {
      s0 = fma(x, x, s0);
    } else if (MODE == 1) {
      const double r = 1.0 / x;
      s0 += x;
      s1 += r;
      s2 = fma(r, r, s2);
    } else {
      const double y = b[i];
      s0 = fma(x, y, s0);
      s1 = fma(y, y, s1);
    }
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = s0; red[1][warp] = s1; red[2][warp] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0, t1 = 0, t2 = 0;
    for (int w = 0; w < 8; ++w) { t0 += red[0][w]; t1 += red[1][w]; t2 += red[2][w]; }
    partial[0 * kRedBlocks + blockIdx.x] = t0;
    partial[1 * kRedBlocks + blockIdx.x] = t1;
    partial[2 * kRedBlocks + blockIdx.x] = t2;
    __threadfence();
    if (atomicAdd(counter, 1u) == gridDim.x - 1) {
      __threadfence();
      t0 = t1 = t2 = 0;
      const volatile double* vp = partial;
      for (unsigned g = 0; g < gridDim.x; ++g) {
        t0 += vp[0 * kRedBlocks + g];
        t1 += vp[1 * kRedBlocks + g];
        t2 += vp[2 * kRedBlocks + g];
      }
      if (MODE == 0) {
        scalars[0] = t0;
      } else if (MODE == 1) {
        const double mean = t0 / (double)n;
        scalars[1] = mean;
        scalars[2] = mean * t1;
        scalars[3] = mean * mean * t2;
      } else {
        scalars[4] = t0;
        scalars[5] = t1;
      }
      *counter = 0;   // ready for the next launch on this stream
    }
  }
}

__global__ void k_inv_s(const double* __restrict__ l2, const double* __restrict__ scalars, int N,
                        double* __restrict__ inv_s, float* __restrict__ inv_s_f, double2* __restrict__ cell_par) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) {
    double v = scalars[1] / l2[i];
    inv_s[i] = v;
    inv_s_f[i] = (float)v;
    cell_par[i].y = v;
  }
}

// Launch configuration of the line passes: (entries per lane and step, CTAs per SM) and the line length above which
// the whole CTA takes a line.  SCL_STAT_VARIANT / SCL_STAT_HEAVY override the defaults (tuning studies only).
struct StatTune { int variant; uint32_t heavy; int writer; };
static StatTune& stat_tune() {
  static StatTune t = [] {
    StatTune x{0, 4096u, 0};
    if (const char* v = getenv("SCL_STAT_VARIANT")) x.variant = atoi(v);
    if (const char* v = getenv("SCL_STAT_HEAVY")) x.heavy = (uint32_t)atoi(v);
    if (const char* v = getenv("SCL_DENSIFY")) x.writer = atoi(v);
    return x;
  }();
  return t;
}
// tuning studies (scl_debug_set_tuning): negative values keep the current setting
void set_norm_tuning(int stat_variant, int stat_heavy, int writer) {
  StatTune& t = stat_tune();
  if (stat_variant >= 0) t.variant = stat_variant;
  if (stat_heavy > 0) t.heavy = (uint32_t)stat_heavy;
  if (writer >= 0) t.writer = writer;
}
template <class Op>
static void launch_lines(const Op& op, const uint32_t* ptr, int n_lines, int reversed, cudaStream_t st) {
  const StatTune t = stat_tune();
  auto grid = [&](int per_sm) { return std::max(1, std::min(n_lines, sm_count() * per_sm)); };
  switch (t.variant) {
    case 1: k_lines<Op, 4, 3><<<grid(3), kStatThreads, 0, st>>>(op, ptr, n_lines, reversed, t.heavy); break;
    case 2: k_lines<Op, 8, 2><<<grid(2), kStatThreads, 0, st>>>(op, ptr, n_lines, reversed, t.heavy); break;
    case 3: k_lines<Op, 2, 4><<<grid(4), kStatThreads, 0, st>>>(op, ptr, n_lines, reversed, t.heavy); break;
    default: k_lines<Op, 4, 4><<<grid(4), kStatThreads, 0, st>>>(op, ptr, n_lines, reversed, t.heavy); break;
  }
}

void compute_norm_stats(const SpMat& A, NormStats& S, cudaStream_t st) {
  const int N = A.N, M = A.M;
  SCL_REQUIRE(N > 1 && M > 0 && A.nnz > 0, "matrix too small to normalise");
  S.tgc.ensure(N); S.l2.ensure(N); S.inv_s.ensure(N); S.inv_s_f.ensure(N);
  S.ybar.ensure(M); S.sigma.ensure(M); S.mu.ensure(M); S.cent.ensure(M);
  S.mu_f.ensure(M); S.cent_f.ensure(M); S.inv_sigma_f.ensure(M);
  S.cell_par.ensure(N); S.gene_par.ensure(M);
  S.sumsq_gene.ensure(M); S.sumsq_cell.ensure(N);
  if (!S.scalars.p) {
    S.scalars.ensure(8);
    S.red_partial.ensure(3 * kRedBlocks);
    S.red_counter.ensure(1);
    SCL_CUDA(cudaMemsetAsync(S.red_counter.p, 0, sizeof(unsigned int), st));
  }
  S.have_patch[0] = S.have_patch[1] = false;
  count_launches(7);
  const int wgrid = min((N + 7) / 8, 148 * 8);
  k_row_sum<<<wgrid, 256, 0, st>>>(A.rowptr.p, A.rval.p, N, S.tgc.p, S.cell_par.p);
  // genes arrive sorted by mean expression (:224): reversed order starts the densest columns first
  launch_lines(GeneStatsOp{A.rowval.p, A.val.p, S.cell_par.p, N, S.ybar.p, S.sigma.p, S.mu.p, S.mu_f.p, S.inv_sigma_f.p,
                           S.gene_par.p},
               A.colptr.p, M, 1, st);
  k_reduce_vec<0><<<kRedBlocks, 256, 0, st>>>(S.mu.p, nullptr, M, S.red_partial.p, S.red_counter.p, S.scalars.p);
  launch_lines(CellL2Op{A.colidx.p, A.rval.p, S.cell_par.p, S.gene_par.p, S.scalars.p, S.l2.p}, A.rowptr.p, N, 0, st);
  k_reduce_vec<1><<<kRedBlocks, 256, 0, st>>>(S.l2.p, nullptr, N, S.red_partial.p, S.red_counter.p, S.scalars.p);
  k_inv_s<<<(N + 255) / 256, 256, 0, st>>>(S.l2.p, S.scalars.p, N, S.inv_s.p, S.inv_s_f.p, S.cell_par.p);
  launch_lines(GeneCenterOp{A.rowval.p, A.val.p, S.cell_par.p, S.scalars.p, N, S.gene_par.p, S.cent.p, S.sumsq_gene.p,
                            S.cent_f.p},
               A.colptr.p, M, 1, st);
  SCL_CUDA(cudaGetLastError());
}

// The writer's sparse patch for one layout (0: gene-major / CSC order, 1: cell-major / CSR order), computed on first
// use after compute_norm_stats: a Gram on the gene side never needs the cell-side patch and vice versa.
void ensure_patch(const SpMat& A, NormStats& S, int layout, cudaStream_t st) {
  SCL_REQUIRE(layout == 0 || layout == 1, "layout must be 0 or 1");
  if (S.have_patch[layout]) return;
  if (layout == 0) {
    S.patch_csc.ensure(A.nnz);
    count_launches(1);
    launch_lines(GenePatchOp{A.rowval.p, A.val.p, S.cell_par.p, S.gene_par.p, S.patch_csc.p}, A.colptr.p, A.M, 1, st);
  } else {
    S.patch_csr.ensure(A.nnz);
    count_launches(2);
    k_reduce_vec<2><<<kRedBlocks, 256, 0, st>>>(S.mu.p, S.cent.p, A.M, S.red_partial.p, S.red_counter.p, S.scalars.p);
    launch_lines(CellFinishOp{A.colidx.p, A.rval.p, S.cell_par.p, S.gene_par.p, S.scalars.p, S.patch_csr.p, S.sumsq_cell.p},
                 A.rowptr.p, A.N, 0, st);
  }
  SCL_CUDA(cudaGetLastError());
  S.have_patch[layout] = true;
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

// ---- TMA-store writer ---------------------------------------------------------------------------
// Same decomposition (a CTA owns a strip of kStripW positions for kLinesPerCta lines, eight positions per thread with
// their factors in registers), but a line's strip is composed in shared memory and leaves the SM as ONE bulk
// asynchronous store (cp.async.bulk.global.shared, 4 KB): the background is written by its owner threads as one
// 16-byte shared-memory store each, the stored entries are dropped on top as 2-byte stores, and no thread reads the
// tile back, tests for patches or issues a global store.  Two lines per step, a ring of kRing steps: the stores of
// step s stream out while steps s+1.. are composed; thread 0 waits (wait_group.read) until the stores that used a
// ring slot have finished reading it before the slot is written again.
static constexpr int kRing = 3;

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst),
               "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}

template <bool CELL_MAJOR, bool WITH_LO>
__global__ void __launch_bounds__(kDenseThreads)
k_densify_tma(const uint32_t* __restrict__ off, const uint32_t* __restrict__ idx, const float* __restrict__ patch,
              const float* __restrict__ inv_s_f, const float* __restrict__ mu_f, const float* __restrict__ cent_f,
              int n_lines, int line_len, size_t ld, int n_strips, int pos0, int pos1, __half* __restrict__ out_hi,
              __half* __restrict__ out_lo) {
  extern __shared__ __align__(128) unsigned char dens_smem[];
  __half* tile_hi = reinterpret_cast<__half*>(dens_smem);     // [kRing][2][kStripW]
  __half* tile_lo = tile_hi + kRing * 2 * kStripW;            // [kRing][2][kStripW] (WITH_LO only)
  __shared__ uint32_t seg[kLinesPerCta][2];
  __shared__ float line_a[kLinesPerCta], line_c[kLinesPerCta];
  const int tid = threadIdx.x;
  const int strip = blockIdx.x % n_strips;
  const int line0 = (blockIdx.x / n_strips) * kLinesPerCta;
  const int line1 = min(n_lines, line0 + kLinesPerCta);
  const int n_my = line1 - line0;
  const int base = pos0 + strip * kStripW;
  const int p0 = base + tid * kElemsPerThread;
  const bool active = p0 < pos1;
  const uint32_t bytes = (uint32_t)(min(kStripW, pos1 - base)) * 2u;   // multiple of 16: pos1 % 8 == 0 or pos1 == ld
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
  if (tid < 2 * n_my) {
    const int l = tid >> 1;
    seg[l][tid & 1] = off[(size_t)(line0 + l) * (n_strips + 1) + strip + (tid & 1)];
  } else if (tid >= 64 && tid < 64 + n_my) {
    const int l = tid - 64;
    line_a[l] = CELL_MAJOR ? -inv_s_f[line0 + l] : -mu_f[line0 + l];
    line_c[l] = CELL_MAJOR ? 0.f : -cent_f[line0 + l];
  }
  __syncthreads();

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
  auto put = [&](__half* th, __half* tl, uint32_t r, float v) {
    const __half h = __float2half_rn(v);
    th[r] = h;
    if (WITH_LO) tl[r] = __float2half_rn(v - __half2float(h));
  };
  auto scatter = [&](__half* th, __half* tl, uint32_t t, uint32_t t_end, uint32_t pos, float v) {
    if (t < t_end) {
      put(th, tl, pos - base, v);
      for (uint32_t u = t + kDenseThreads; u < t_end; u += kDenseThreads) put(th, tl, idx[u] - base, patch[u]);
    }
  };
  auto background = [&](__half* th, __half* tl, int l) {
    const float a = line_a[l], c = line_c[l];
    float f[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) f[q] = fmaf(a, P[q], CELL_MAJOR ? Q[q] : c);
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
    *reinterpret_cast<uint4*>(th + tid * kElemsPerThread) = *reinterpret_cast<const uint4*>(h2);
    if (WITH_LO) *reinterpret_cast<uint4*>(tl + tid * kElemsPerThread) = *reinterpret_cast<const uint4*>(l2);
  };

  uint32_t t0, e0, q0, t1, e1, q1;
  float v0, v1;
  load_patch(0, t0, e0, q0, v0);
  load_patch(1, t1, e1, q1, v1);
  int slot = 0;
  for (int l = 0; l < n_my; l += 2) {
    __half* th0 = tile_hi + (size_t)(slot * 2) * kStripW;
    __half* th1 = th0 + kStripW;
    __half* tl0 = tile_lo + (size_t)(slot * 2) * kStripW;
    __half* tl1 = tl0 + kStripW;
    const bool two = l + 1 < n_my;
    if (active) {
      background(th0, tl0, l);
      if (two) background(th1, tl1, l + 1);
    }
    __syncthreads();
    scatter(th0, tl0, t0, e0, q0, v0);
    scatter(th1, tl1, t1, e1, q1, v1);
    load_patch(l + 2, t0, e0, q0, v0);
    load_patch(l + 3, t1, e1, q1, v1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy tile writes -> visible to the bulk copy
    // the slot the NEXT step composes in was read by the stores committed kRing - 1 steps ago
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kRing - 2) : "memory");
    __syncthreads();
    if (tid == 0) {
      __half* g0 = out_hi + (size_t)(line0 + l) * ld + (size_t)base;
      bulk_store(g0, th0, bytes);
      if (two) bulk_store(g0 + ld, th1, bytes);
      if (WITH_LO) {
        __half* gl0 = out_lo + (size_t)(line0 + l) * ld + (size_t)base;
        bulk_store(gl0, tl0, bytes);
        if (two) bulk_store(gl0 + ld, tl1, bytes);
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    slot = slot + 1 == kRing ? 0 : slot + 1;
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory stays valid until read
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
  if (!gene_side) ensure_patch(A, S, 1, st);   // the cell-side sums of squares come out of the cell-major finishing pass
  return gene_side ? S.sumsq_gene.p : S.sumsq_cell.p;
}

void set_gram_diagonal(float* G, int n, const double* sumsq, double scale, cudaStream_t st) {
  count_launches(1);
  k_set_diagonal<<<(n + 255) / 256, 256, 0, st>>>(G, n, sumsq, scale);
  SCL_CUDA(cudaGetLastError());
}

void densify(const SpMat& A, NormStats& S, int layout, size_t ld, __half* out_hi, __half* out_lo,
             cudaStream_t st, long long pos0, long long pos1) {
  ensure_patch(A, S, layout, st);
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
  if (stat_tune().writer == 1) {   // TMA-store writer
    const size_t smem = (size_t)kRing * 2 * kStripW * sizeof(__half) * (out_lo ? 2 : 1);
#define SCL_LAUNCH_TMA(CM, LO)                                                                                         \
  do {                                                                                                                 \
    SCL_CUDA(cudaFuncSetAttribute(k_densify_tma<CM, LO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    k_densify_tma<CM, LO><<<(unsigned)ctas, kDenseThreads, smem, st>>>(off.p, idx, cell_major ? S.patch_csr.p : S.patch_csc.p, \
                                                                    S.inv_s_f.p, S.mu_f.p, S.cent_f.p, n_lines, line_len, \
                                                                    ld, n_strips, (int)pos0, (int)pos1, out_hi, out_lo); \
  } while (0)
    if (cell_major) {
      if (out_lo) SCL_LAUNCH_TMA(true, true); else SCL_LAUNCH_TMA(true, false);
    } else {
      if (out_lo) SCL_LAUNCH_TMA(false, true); else SCL_LAUNCH_TMA(false, false);
    }
#undef SCL_LAUNCH_TMA
    SCL_CUDA(cudaGetLastError());
    return;
  }
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
