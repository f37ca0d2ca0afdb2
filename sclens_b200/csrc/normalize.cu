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

// ---- strips ------------------------------------------------------------------------------------
// The dense writer and the strip passes tile a matrix into strips of kStripW positions (cells for gene lines, genes
// for cell lines); k_strip_offsets locates every line's entries inside every strip.
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


// ---- log1p on [0, 1] -------------------------------------------------------------------------
// u = x_ij / r_i lies in (0, 1] and is almost always tiny (a count over a cell's total).  log1p(u) = 2 atanh(s),
// s = u / (2 + u): for u <= 1/16, s^2 <= 9.2e-4 and the odd series truncated after s^13 is exact to < 1e-22
// relative, so the result carries the rounding of the division and of the last FMA only (~1 ulp).  A third of the
// Float64 instructions of the library log1p; the statistics passes re-evaluate it instead of storing y per entry.
__device__ __forceinline__ double log1p_unit(double u) {
  if (!(fabs(u) <= 0.0625)) return log1p(u);
  // s = u / d, d = 2 + u in [1.9375, 2.0625]: hardware reciprocal seed (2^-23), two Newton steps (-> below 2^-53),
  // one residual correction of the quotient; no special cases can occur, so none of the library division's guards
  const double d = 2.0 + u;
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  double s = u * r;
  s = fma(fma(-d, s, u), r, s);
  const double w = s * s;
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

// Three-stage variant: the (index, value) stream runs two steps ahead and the gathers one step ahead of the Float64
// arithmetic, so neither the HBM latency of the stream nor the L2 latency of the gathers is waited for in a step.
template <class Op, int U>
__device__ __forceinline__ void line_span_deep(const Op& op, const typename Op::Ctx& cx, uint32_t b, uint32_t e, int first,
                                               int stride, double (&acc)[Op::NACC]) {
  uint32_t t = b + first;
  const uint32_t step = (uint32_t)stride * U;
  float va[U], vb[U], vc[U];
  uint32_t ib[U], ic[U];
  typename Op::Ld la[U], lb[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const uint32_t t0 = t + stride * u, t1 = t0 + step;
    const bool ok0 = t0 < e, ok1 = t1 < e;
    const uint32_t i0 = ok0 ? op.idx[t0] : 0u;
    va[u] = ok0 ? op.val[t0] : 0.f;
    ib[u] = ok1 ? op.idx[t1] : 0u;
    vb[u] = ok1 ? op.val[t1] : 0.f;
    op.gather(i0, la[u]);
  }
  while (t < e) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t t2 = t + 2 * step + stride * u;
      const bool ok = t2 < e;
      ic[u] = ok ? op.idx[t2] : 0u;
      vc[u] = ok ? op.val[t2] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) op.gather(ib[u], lb[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t tt = t + stride * u;
      if (tt < e) op.accum(tt, cx, va[u], la[u], acc);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) { va[u] = vb[u]; la[u] = lb[u]; vb[u] = vc[u]; ib[u] = ic[u]; }
    t += step;
  }
}

template <class Op, int U, int MINB, bool DEEP = false>
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
    if (DEEP) line_span_deep<Op, U>(op, cx, b, e, lane, 32, acc); else line_span<Op, U>(op, cx, b, e, lane, 32, acc);
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
    if (DEEP) line_span_deep<Op, U>(op, cx, b, e, tid, kStatThreads, acc); else line_span<Op, U>(op, cx, b, e, tid, kStatThreads, acc);
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
  const double* centre = nullptr;    // per-gene median of y (centering="median"); nullptr: the mean is the centre
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
    const double ctr = centre ? centre[j] : m;
    ybar[j] = m;
    sigma[j] = sd;
    mu[j] = ctr / sd;
    mu_f[j] = (float)(ctr / sd);
    inv_sigma_f[j] = (float)(1.0 / sd);
    gene_par[j] = make_double4(1.0 / sd, ctr / sd, 0.0, 0.0);
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
  int no_center = 0;                 // centering="median": norm_l is the last step, no re-centring (:653-654)
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
    const double c = no_center ? 0.0 : (su - m * sum_inv_s) / (double)N;
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

// ---- strip passes -----------------------------------------------------------------------------
// The line passes above gather one 32-byte L2 sector per stored entry (the other side's parameters): 3.5 GB of L2
// traffic per pass at 68k x 20k beside 0.9 GB of streamed entries, which is what bounds them (measured: 9 TB/s of L2
// sectors at 1.8 TB/s of HBM).  A strip pass tiles the same work the way the dense writer does: a CTA owns a strip
// of kStripW positions for a block of lines, copies the strip's parameters into shared memory once (16-24 bytes per
// position) and walks the lines' segments inside the strip (found with the writer's strip-offset table), one warp
// per segment, gathering from shared memory.  Per-(line, strip) partial sums go to a small table that a finishing
// kernel adds in strip order - deterministic, and balanced because a unit of work is at most kStripW entries.
static constexpr int kStripLines = 64;    // lines per CTA

template <class Op>
__global__ void __launch_bounds__(kStatThreads, 4)
k_strips(const Op op, const uint32_t* __restrict__ off, int n_lines, int line_len, int n_strips, int reversed,
         double* __restrict__ partial) {
  extern __shared__ __align__(16) double strip_par[];   // [Op::NPAR][kStripW]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int strip = blockIdx.x % n_strips;
  const int lb = blockIdx.x / n_strips;
  const int base = strip * kStripW;
  for (int i = tid; i < kStripW; i += kStatThreads) op.stage(strip_par, i, min(base + i, line_len - 1));
  __syncthreads();
  constexpr int U = 4;
  for (int k = warp; k < kStripLines; k += kStatThreads / 32) {
    const int li = lb * kStripLines + k;
    if (li >= n_lines) break;
    const int line = reversed ? n_lines - 1 - li : li;
    const uint32_t b = off[(size_t)line * (n_strips + 1) + strip], e = off[(size_t)line * (n_strips + 1) + strip + 1];
    double acc[Op::NACC];
#pragma unroll
    for (int a = 0; a < Op::NACC; ++a) acc[a] = 0;
    if (b < e) {
      const typename Op::Ctx cx = op.begin(line);
      uint32_t t = b + lane;
      uint32_t idx[U], nidx[U];
      float v[U], nv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t tt = t + 32 * u;
        const bool ok = tt < e;
        idx[u] = ok ? op.idx[tt] : (uint32_t)base;
        v[u] = ok ? op.val[tt] : 0.f;
      }
      while (t < e) {
        const uint32_t tn = t + 32 * U;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const uint32_t tt = tn + 32 * u;
          const bool ok = tt < e;
          nidx[u] = ok ? op.idx[tt] : (uint32_t)base;
          nv[u] = ok ? op.val[tt] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const uint32_t tt = t + 32 * u;
          if (tt < e) op.accum(tt, cx, v[u], strip_par, idx[u] - (uint32_t)base, acc);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) { idx[u] = nidx[u]; v[u] = nv[u]; }
        t = tn;
      }
    }
    if (Op::NACC_OUT > 0) {
#pragma unroll
      for (int a = 0; a < Op::NACC; ++a) acc[a] = warp_sum(acc[a]);
      if (lane == 0) {
#pragma unroll
        for (int a = 0; a < Op::NACC_OUT; ++a) partial[((size_t)a * n_strips + strip) * n_lines + line] = acc[a];
      }
    }
  }
}

// per-line sums = partial sums added in strip order, then the pass's closing formula
template <class Op>
__global__ void k_strips_finish(const Op op, const double* __restrict__ partial, int n_lines, int n_strips) {
  const int line = blockIdx.x * blockDim.x + threadIdx.x;
  if (line >= n_lines) return;
  double acc[Op::NACC];
#pragma unroll
  for (int a = 0; a < Op::NACC; ++a) {
    double t = 0;
    for (int s = 0; s < n_strips; ++s) t += partial[((size_t)a * n_strips + s) * n_lines + line];
    acc[a] = t;
  }
  op.finish(line, op.begin(line), acc);
}

// The strip versions of the five passes: same arithmetic and closing formulas as the line-pass operators (which they
// wrap), parameters read from the CTA's shared-memory copy.
struct SGeneStatsOp : GeneStatsOp {
  static constexpr int NPAR = 1, NACC_OUT = NACC;
  __device__ __forceinline__ void stage(double* sp, int i, int pos) const { sp[i] = cell_par[pos].x; }
  __device__ __forceinline__ void accum(uint32_t t, const Ctx& cx, float v, const double* sp, uint32_t r, double (&acc)[NACC]) const {
    Ld l;
    l.inv_r = sp[r];
    GeneStatsOp::accum(t, cx, v, l, acc);
  }
};
struct SCellL2Op : CellL2Op {
  static constexpr int NPAR = 2, NACC_OUT = NACC;
  __device__ __forceinline__ void stage(double* sp, int i, int pos) const {
    const double2 g = *reinterpret_cast<const double2*>(&gene_par[pos]);
    sp[i] = g.x;
    sp[kStripW + i] = g.y;
  }
  __device__ __forceinline__ void accum(uint32_t t, const Ctx& cx, float v, const double* sp, uint32_t r, double (&acc)[NACC]) const {
    Ld l;
    l.inv_sd = sp[r];
    l.mu = sp[kStripW + r];
    CellL2Op::accum(t, cx, v, l, acc);
  }
};
struct SGeneCenterOp : GeneCenterOp {
  static constexpr int NPAR = 2, NACC_OUT = NACC;
  __device__ __forceinline__ void stage(double* sp, int i, int pos) const {
    const double2 c = cell_par[pos];
    sp[i] = c.x;
    sp[kStripW + i] = c.y;
  }
  __device__ __forceinline__ void accum(uint32_t t, const Ctx& cx, float v, const double* sp, uint32_t r, double (&acc)[NACC]) const {
    Ld l;
    l.c = make_double2(sp[r], sp[kStripW + r]);
    GeneCenterOp::accum(t, cx, v, l, acc);
  }
};
struct SGenePatchOp : GenePatchOp {
  static constexpr int NPAR = 2, NACC_OUT = 0;
  __device__ __forceinline__ void stage(double* sp, int i, int pos) const {
    const double2 c = cell_par[pos];
    sp[i] = c.x;
    sp[kStripW + i] = c.y;
  }
  __device__ __forceinline__ void accum(uint32_t t, const Ctx& cx, float v, const double* sp, uint32_t r, double (&acc)[NACC]) const {
    Ld l;
    l.c = make_double2(sp[r], sp[kStripW + r]);
    GenePatchOp::accum(t, cx, v, l, acc);
  }
};
struct SCellFinishOp : CellFinishOp {
  static constexpr int NPAR = 3, NACC_OUT = NACC;
  __device__ __forceinline__ void stage(double* sp, int i, int pos) const {
    const double4 g = gene_par[pos];
    sp[i] = g.x;
    sp[kStripW + i] = g.y;
    sp[2 * kStripW + i] = g.z;
  }
  __device__ __forceinline__ void accum(uint32_t t, const Ctx& cx, float v, const double* sp, uint32_t r, double (&acc)[NACC]) const {
    Ld l;
    l.g = make_double4(sp[r], sp[kStripW + r], sp[2 * kStripW + r], 0.0);
    CellFinishOp::accum(t, cx, v, l, acc);
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
    StatTune x{4, 4096u, -1};   // writer -1: automatic (staged TMA-store writer for hi-only output, overlay writer with lo)
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
  if (writer >= -1) t.writer = writer;
}
template <class Op>
static void launch_lines(const Op& op, const uint32_t* ptr, int n_lines, int reversed, cudaStream_t st) {
  const StatTune t = stat_tune();
  auto grid = [&](int per_sm) { return std::max(1, std::min(n_lines, sm_count() * per_sm)); };
  switch (t.variant) {
    case 1: k_lines<Op, 4, 3><<<grid(3), kStatThreads, 0, st>>>(op, ptr, n_lines, reversed, t.heavy); break;
    case 2: k_lines<Op, 8, 2><<<grid(2), kStatThreads, 0, st>>>(op, ptr, n_lines, reversed, t.heavy); break;
    case 3: k_lines<Op, 2, 4><<<grid(4), kStatThreads, 0, st>>>(op, ptr, n_lines, reversed, t.heavy); break;
    case 4: k_lines<Op, 4, 3, true><<<grid(3), kStatThreads, 0, st>>>(op, ptr, n_lines, reversed, t.heavy); break;
    case 5: k_lines<Op, 2, 4, true><<<grid(4), kStatThreads, 0, st>>>(op, ptr, n_lines, reversed, t.heavy); break;
    case 6: k_lines<Op, 4, 4, true><<<grid(4), kStatThreads, 0, st>>>(op, ptr, n_lines, reversed, t.heavy); break;
    case 7: k_lines<Op, 4, 2, true><<<grid(2), kStatThreads, 0, st>>>(op, ptr, n_lines, reversed, t.heavy); break;
    default: k_lines<Op, 4, 4><<<grid(4), kStatThreads, 0, st>>>(op, ptr, n_lines, reversed, t.heavy); break;
  }
}

template <class Op>
static void launch_strips(const Op& op, const uint32_t* off, int n_lines, int line_len, int n_strips, int reversed,
                          double* partial, cudaStream_t st) {
  const int blocks = (n_lines + kStripLines - 1) / kStripLines;
  const size_t smem = (size_t)Op::NPAR * kStripW * sizeof(double);
  count_launches(Op::NACC_OUT > 0 ? 2 : 1);
  k_strips<Op><<<(unsigned)(blocks * n_strips), kStatThreads, smem, st>>>(op, off, n_lines, line_len, n_strips, reversed, partial);
  if (Op::NACC_OUT > 0) k_strips_finish<Op><<<(n_lines + 255) / 256, 256, 0, st>>>(op, partial, n_lines, n_strips);
}

// strip offsets of both orientations (kept for the dense writer) and the partial-sum table
static void prepare_strips(const SpMat& A, NormStats& S, cudaStream_t st) {
  const int N = A.N, M = A.M;
  const int ns_c = (N + kStripW - 1) / kStripW, ns_r = (M + kStripW - 1) / kStripW;   // strips along cells / along genes
  S.off[0].ensure((size_t)M * (ns_c + 1));
  S.off[1].ensure((size_t)N * (ns_r + 1));
  S.strip_partial.ensure((size_t)3 * std::max((size_t)M * ns_c, (size_t)N * ns_r));
  count_launches(2);
  {
    const long long total = (long long)M * (ns_c + 1);
    k_strip_offsets<<<(int)std::min<long long>((total + 255) / 256, 148LL * 16), 256, 0, st>>>(A.colptr.p, A.rowval.p, M, ns_c, 0, S.off[0].p);
  }
  {
    const long long total = (long long)N * (ns_r + 1);
    k_strip_offsets<<<(int)std::min<long long>((total + 255) / 256, 148LL * 16), 256, 0, st>>>(A.rowptr.p, A.colidx.p, N, ns_r, 0, S.off[1].p);
  }
  S.off_strips[0] = ns_c;
  S.off_strips[1] = ns_r;
  S.off_valid[0] = S.off_valid[1] = true;
}

// ---- per-gene median of y = log1p(x / r_i) over all N cells (centering="median", :294-299) --------------------------
// `mapslices(median, X, dims=1)` on the dense Float32 matrix: the N - nnz_j implicit zeros sort first, so the value at sorted
// position k is 0 for k < zeros and the (k - zeros)-th smallest stored value otherwise; the median is the mean of positions
// floor((N-1)/2) and ceil((N-1)/2).  Stored values are selected by an 8-bit radix select over the bit pattern of
// u = x / r_i (positive doubles order like their bit patterns, and y is increasing in u): one CTA per gene, eight passes.
__device__ double radix_select_u(const uint32_t* __restrict__ rowval, const float* __restrict__ val, uint32_t b, uint32_t e,
                                 const double2* __restrict__ cell_par, uint32_t rank, uint32_t* hist) {
  unsigned long long prefix = 0, mask = 0;
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (uint32_t t = b + threadIdx.x; t < e; t += blockDim.x) {
      const unsigned long long key = (unsigned long long)__double_as_longlong((double)val[t] * cell_par[rowval[t]].x);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xff], 1u);
    }
    __syncthreads();
    // bucket holding the wanted rank (every thread walks the 256 counters: uniform result, no broadcast needed)
    uint32_t acc = 0;
    int bucket = 0;
    for (; bucket < 256; ++bucket) {
      const uint32_t c = hist[bucket];
      if (rank < acc + c) break;
      acc += c;
    }
    rank -= acc;
    prefix |= (unsigned long long)bucket << shift;
    mask |= 0xffull << shift;
    __syncthreads();
  }
  return __longlong_as_double((long long)prefix);
}

__global__ void __launch_bounds__(256) k_gene_median(const uint32_t* __restrict__ colptr, const uint32_t* __restrict__ rowval,
                                                     const float* __restrict__ val, const double2* __restrict__ cell_par, int N,
                                                     int M, double* __restrict__ med) {
  __shared__ uint32_t hist[256];
  for (int j = blockIdx.x; j < M; j += gridDim.x) {
    const uint32_t b = colptr[j], e = colptr[j + 1], nz = e - b, zeros = (uint32_t)N - nz;
    const uint32_t k1 = (uint32_t)(N - 1) / 2, k2 = (uint32_t)N / 2;     // the two middle positions (equal when N is odd)
    double m = 0.0;
    if (k2 >= zeros) {                                                   // otherwise both middle values are zeros
      const double u2 = radix_select_u(rowval, val, b, e, cell_par, k2 - zeros, hist);
      const double y2 = log1p(u2);
      double y1 = 0.0;
      if (k1 == k2) y1 = y2;
      else if (k1 >= zeros) y1 = log1p(radix_select_u(rowval, val, b, e, cell_par, k1 - zeros, hist));
      m = 0.5 * (y1 + y2);
    }
    if (threadIdx.x == 0) med[j] = m;
  }
}

static void compute_norm_stats_strips(const SpMat& A, NormStats& S, cudaStream_t st) {
  const int N = A.N, M = A.M;
  prepare_strips(A, S, st);
  const int ns_c = S.off_strips[0], ns_r = S.off_strips[1];
  count_launches(4);
  const int wgrid = min((N + 7) / 8, 148 * 8);
  k_row_sum<<<wgrid, 256, 0, st>>>(A.rowptr.p, A.rval.p, N, S.tgc.p, S.cell_par.p);
  SGeneStatsOp gs;
  static_cast<GeneStatsOp&>(gs) = GeneStatsOp{A.rowval.p, A.val.p, S.cell_par.p, N, S.ybar.p, S.sigma.p, S.mu.p, S.mu_f.p,
                                              S.inv_sigma_f.p, S.gene_par.p};
  launch_strips(gs, S.off[0].p, M, N, ns_c, 1, S.strip_partial.p, st);
  k_reduce_vec<0><<<kRedBlocks, 256, 0, st>>>(S.mu.p, nullptr, M, S.red_partial.p, S.red_counter.p, S.scalars.p);
  SCellL2Op cl;
  static_cast<CellL2Op&>(cl) = CellL2Op{A.colidx.p, A.rval.p, S.cell_par.p, S.gene_par.p, S.scalars.p, S.l2.p};
  launch_strips(cl, S.off[1].p, N, M, ns_r, 0, S.strip_partial.p, st);
  k_reduce_vec<1><<<kRedBlocks, 256, 0, st>>>(S.l2.p, nullptr, N, S.red_partial.p, S.red_counter.p, S.scalars.p);
  k_inv_s<<<(N + 255) / 256, 256, 0, st>>>(S.l2.p, S.scalars.p, N, S.inv_s.p, S.inv_s_f.p, S.cell_par.p);
  SGeneCenterOp gc;
  static_cast<GeneCenterOp&>(gc) = GeneCenterOp{A.rowval.p, A.val.p, S.cell_par.p, S.scalars.p, N, S.gene_par.p, S.cent.p,
                                                S.sumsq_gene.p, S.cent_f.p};
  launch_strips(gc, S.off[0].p, M, N, ns_c, 1, S.strip_partial.p, st);
  SCL_CUDA(cudaGetLastError());
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
  S.off_valid[0] = S.off_valid[1] = false;
  const bool median = S.centering == 1;
  if (stat_tune().variant == 8 && !median) {
    compute_norm_stats_strips(A, S, st);
    return;
  }
  count_launches(7);
  const int wgrid = min((N + 7) / 8, 148 * 8);
  k_row_sum<<<wgrid, 256, 0, st>>>(A.rowptr.p, A.rval.p, N, S.tgc.p, S.cell_par.p);
  if (median) {
    S.median.ensure(M);
    count_launches(1);
    k_gene_median<<<min(M, 148 * 8), 256, 0, st>>>(A.colptr.p, A.rowval.p, A.val.p, S.cell_par.p, N, M, S.median.p);
  }
  // genes arrive sorted by mean expression (:224): reversed order starts the densest columns first
  launch_lines(GeneStatsOp{A.rowval.p, A.val.p, S.cell_par.p, N, S.ybar.p, S.sigma.p, S.mu.p, S.mu_f.p, S.inv_sigma_f.p,
                           S.gene_par.p, median ? S.median.p : nullptr},
               A.colptr.p, M, 1, st);
  k_reduce_vec<0><<<kRedBlocks, 256, 0, st>>>(S.mu.p, nullptr, M, S.red_partial.p, S.red_counter.p, S.scalars.p);
  launch_lines(CellL2Op{A.colidx.p, A.rval.p, S.cell_par.p, S.gene_par.p, S.scalars.p, S.l2.p}, A.rowptr.p, N, 0, st);
  k_reduce_vec<1><<<kRedBlocks, 256, 0, st>>>(S.l2.p, nullptr, N, S.red_partial.p, S.red_counter.p, S.scalars.p);
  k_inv_s<<<(N + 255) / 256, 256, 0, st>>>(S.l2.p, S.scalars.p, N, S.inv_s.p, S.inv_s_f.p, S.cell_par.p);
  launch_lines(GeneCenterOp{A.rowval.p, A.val.p, S.cell_par.p, S.scalars.p, N, S.gene_par.p, S.cent.p, S.sumsq_gene.p,
                            S.cent_f.p, median ? 1 : 0},
               A.colptr.p, M, 1, st);
  SCL_CUDA(cudaGetLastError());
}

// The writer's sparse patch for one layout (0: gene-major / CSC order, 1: cell-major / CSR order), computed on first
// use after compute_norm_stats: a Gram on the gene side never needs the cell-side patch and vice versa.
void ensure_patch(const SpMat& A, NormStats& S, int layout, cudaStream_t st) {
  SCL_REQUIRE(layout == 0 || layout == 1, "layout must be 0 or 1");
  if (S.have_patch[layout]) return;
  if (stat_tune().variant == 8 && S.off_valid[0] && S.off_valid[1]) {
    if (layout == 0) {
      S.patch_csc.ensure(A.nnz);
      SGenePatchOp gp;
      static_cast<GenePatchOp&>(gp) = GenePatchOp{A.rowval.p, A.val.p, S.cell_par.p, S.gene_par.p, S.patch_csc.p};
      launch_strips(gp, S.off[0].p, A.M, A.N, S.off_strips[0], 1, S.strip_partial.p, st);
    } else {
      S.patch_csr.ensure(A.nnz);
      count_launches(1);
      k_reduce_vec<2><<<kRedBlocks, 256, 0, st>>>(S.mu.p, S.cent.p, A.M, S.red_partial.p, S.red_counter.p, S.scalars.p);
      SCellFinishOp cf;
      static_cast<CellFinishOp&>(cf) = CellFinishOp{A.colidx.p, A.rval.p, S.cell_par.p, S.gene_par.p, S.scalars.p, S.patch_csr.p,
                                                    S.sumsq_cell.p};
      launch_strips(cf, S.off[1].p, A.N, A.M, S.off_strips[1], 0, S.strip_partial.p, st);
    }
    SCL_CUDA(cudaGetLastError());
    S.have_patch[layout] = true;
    return;
  }
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

// ---- TMA-store writer with staged patches ---------------------------------------------------------
// As k_densify_tma, but the (position, value) pairs of a line's strip reach the scatter through shared memory:
// every thread requests its own entries of the NEXT step's lines with 4-byte cp.async copies at the top of a step
// (up to KP entries per thread and line, i.e. KP * 256 entries per strip-line; lines denser than that take the
// direct path for the remainder) and reads them back a full step later, after cp.async.wait_group - no register
// scoreboard is held across the step and no thread waits for another's loads.
template <bool CELL_MAJOR, bool WITH_LO, int KP>
__global__ void __launch_bounds__(kDenseThreads)
k_densify_tma2(const uint32_t* __restrict__ off, const uint32_t* __restrict__ idx, const float* __restrict__ patch,
               const float* __restrict__ inv_s_f, const float* __restrict__ mu_f, const float* __restrict__ cent_f,
               int n_lines, int line_len, size_t ld, int n_strips, int pos0, int pos1, __half* __restrict__ out_hi,
               __half* __restrict__ out_lo) {
  extern __shared__ __align__(128) unsigned char dens_smem[];
  __half* tile_hi = reinterpret_cast<__half*>(dens_smem);                              // [kRing][2][kStripW]
  __half* tile_lo = tile_hi + (WITH_LO ? kRing * 2 * kStripW : 0);                     // [kRing][2][kStripW]
  uint32_t* st_idx = reinterpret_cast<uint32_t*>(tile_lo + kRing * 2 * kStripW);       // [2][2][KP][256]
  float* st_val = reinterpret_cast<float*>(st_idx + 2 * 2 * KP * kDenseThreads);       // [2][2][KP][256]
  __shared__ uint32_t seg[kLinesPerCta + 2][2];
  __shared__ float line_a[kLinesPerCta], line_c[kLinesPerCta];
  const int tid = threadIdx.x;
  const int strip = blockIdx.x % n_strips;
  const int line0 = (blockIdx.x / n_strips) * kLinesPerCta;
  const int line1 = min(n_lines, line0 + kLinesPerCta);
  const int n_my = line1 - line0;
  const int base = pos0 + strip * kStripW;
  const int p0 = base + tid * kElemsPerThread;
  const bool active = p0 < pos1;
  const uint32_t bytes = (uint32_t)(min(kStripW, pos1 - base)) * 2u;
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
  if (tid < 2 * (kLinesPerCta + 2)) {
    const int l = tid >> 1;
    seg[l][tid & 1] = l < n_my ? off[(size_t)(line0 + l) * (n_strips + 1) + strip + (tid & 1)] : 0u;   // empty beyond n_my
  } else if (tid >= 96 && tid < 96 + n_my) {
    const int l = tid - 96;
    line_a[l] = CELL_MAJOR ? -inv_s_f[line0 + l] : -mu_f[line0 + l];
    line_c[l] = CELL_MAJOR ? 0.f : -cent_f[line0 + l];
  }
  __syncthreads();

  // this thread's entries of lines l, l+1 -> staging buffer sb (one cp.async group per call, possibly empty)
  auto stage = [&](int l, int sb) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint32_t tb = seg[l + j][0], te = seg[l + j][1];
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        const uint32_t t = tb + tid + k * kDenseThreads;
        if (t < te) {
          const int slot = ((sb * 2 + j) * KP + k) * kDenseThreads + tid;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(st_idx + slot)),
                       "l"(idx + t) : "memory");
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(st_val + slot)),
                       "l"(patch + t) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto put = [&](__half* th, __half* tl, uint32_t r, float v) {
    const __half h = __float2half_rn(v);
    th[r] = h;
    if (WITH_LO) tl[r] = __float2half_rn(v - __half2float(h));
  };
  auto scatter = [&](__half* th, __half* tl, int l, int j, int sb) {
    const uint32_t tb = seg[l + j][0], te = seg[l + j][1];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      const uint32_t t = tb + tid + k * kDenseThreads;
      if (t < te) {
        const int slot = ((sb * 2 + j) * KP + k) * kDenseThreads + tid;
        put(th, tl, st_idx[slot] - base, st_val[slot]);
      }
    }
    for (uint32_t t = tb + tid + KP * kDenseThreads; t < te; t += kDenseThreads) put(th, tl, idx[t] - base, patch[t]);
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

  stage(0, 0);
  int slot = 0, sb = 0;
  for (int l = 0; l < n_my; l += 2, sb ^= 1) {
    stage(l + 2, sb ^ 1);     // lines beyond n_my have empty segments: an empty group
    __half* th0 = tile_hi + (size_t)(slot * 2) * kStripW;
    __half* th1 = th0 + kStripW;
    __half* tl0 = tile_lo + (size_t)(slot * 2) * kStripW;
    __half* tl1 = tl0 + kStripW;
    const bool two = l + 1 < n_my;
    if (active) {
      background(th0, tl0, l);
      if (two) background(th1, tl1, l + 1);
    }
    asm volatile("cp.async.wait_group 1;" ::: "memory");   // this step's staged entries (all but the newest group) have landed
    __syncthreads();
    scatter(th0, tl0, l, 0, sb);
    scatter(th1, tl1, l, 1, sb);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
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
             cudaStream_t st, long long pos0, long long pos1, size_t slice_stride) {
  ensure_patch(A, S, layout, st);
  count_launches(2);
  const bool cell_major = layout == 1;
  const int n_lines = cell_major ? A.N : A.M;
  const int line_len = cell_major ? A.M : A.N;
  SCL_REQUIRE(ld % 8 == 0 && ld >= (size_t)line_len, "leading dimension must be a multiple of 8 and >= line length");
  if (pos1 < 0) { pos0 = 0; pos1 = (long long)ld; }
  SCL_REQUIRE(pos0 % 8 == 0 && pos0 >= 0 && pos1 <= (long long)ld && pos1 > pos0 && (pos1 % 8 == 0 || pos1 == (long long)ld), "bad densify range");
  if (slice_stride) {
    // sliced output (a rank's block of the contraction axis): the buffer holds positions [pos0, pos1) only, lines
    // slice_stride apart - the kernels address out + line * stride + position, so the base moves back by pos0
    SCL_REQUIRE(slice_stride % 8 == 0 && (long long)slice_stride >= pos1 - pos0, "bad slice stride");
    out_hi -= pos0;
    if (out_lo) out_lo -= pos0;
    ld = slice_stride;
  }
  const int n_strips = (int)((pos1 - pos0 + kStripW - 1) / kStripW);
  const uint32_t* ptr = cell_major ? A.rowptr.p : A.colptr.p;
  const uint32_t* idx = cell_major ? A.colidx.p : A.rowval.p;
  // the strip passes leave the full-range offset table of both orientations behind; a rank's cell block needs its own
  const bool reuse = !slice_stride && pos0 == 0 && pos1 == (long long)ld && S.off_valid[layout] && S.off_strips[layout] == n_strips;
  Tmp<uint32_t> off_tmp(reuse ? 1 : (size_t)n_lines * (n_strips + 1), st);
  struct { const uint32_t* p; } off{reuse ? S.off[layout].p : off_tmp.p};
  if (!reuse) {
    const long long total = (long long)n_lines * (n_strips + 1);
    const int grid = (int)std::min<long long>((total + 255) / 256, 148LL * 16);
    k_strip_offsets<<<grid, 256, 0, st>>>(ptr, idx, n_lines, n_strips, (int)pos0, off_tmp.p);
  }
  const long long ctas = (long long)n_strips * ((n_lines + kLinesPerCta - 1) / kLinesPerCta);
  SCL_REQUIRE(ctas < (1LL << 31), "densify grid too large");
  // measured (profiles/r1_tune_norm_*.json): the staged TMA-store writer wins for hi-only output (4.5-4.9 TB/s vs
  // 3.8-4.45), the overlay writer for hi+lo output at the 68k x 20k shape (5.57 vs 4.9-5.2 TB/s)
  const int writer = stat_tune().writer == -1 ? (out_lo ? 0 : 2) : stat_tune().writer;
  if (writer >= 2) {   // TMA-store writer with cp.async-staged patches (2: one staged entry per thread and line, 3: two)
    const int kp = writer == 2 ? 1 : 2;
    const size_t smem = (size_t)kRing * 2 * kStripW * sizeof(__half) * (out_lo ? 2 : 1) + (size_t)2 * 2 * kp * kDenseThreads * 8;
#define SCL_LAUNCH_TMA2(CM, LO, KP)                                                                                     \
  do {                                                                                                                 \
    SCL_CUDA(cudaFuncSetAttribute(k_densify_tma2<CM, LO, KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_densify_tma2<CM, LO, KP><<<(unsigned)ctas, kDenseThreads, smem, st>>>(off.p, idx, cell_major ? S.patch_csr.p : S.patch_csc.p, \
                                                                         S.inv_s_f.p, S.mu_f.p, S.cent_f.p, n_lines, line_len, \
                                                                         ld, n_strips, (int)pos0, (int)pos1, out_hi, out_lo); \
  } while (0)
#define SCL_LAUNCH_TMA2_K(CM, LO) do { if (kp == 1) SCL_LAUNCH_TMA2(CM, LO, 1); else SCL_LAUNCH_TMA2(CM, LO, 2); } while (0)
    if (cell_major) {
      if (out_lo) SCL_LAUNCH_TMA2_K(true, true); else SCL_LAUNCH_TMA2_K(true, false);
    } else {
      if (out_lo) SCL_LAUNCH_TMA2_K(false, true); else SCL_LAUNCH_TMA2_K(false, false);
    }
#undef SCL_LAUNCH_TMA2_K
#undef SCL_LAUNCH_TMA2
    SCL_CUDA(cudaGetLastError());
    return;
  }
  if (writer == 1) {   // TMA-store writer
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
