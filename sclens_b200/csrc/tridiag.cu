// Symmetric tridiagonal eigenproblem in Float64 on the device: the middle third of the full-spectrum solve
// (_get_eigen, src/scLENS.jl:375-382).  The library's one-stage solver spends its time in three parts - Householder
// tridiagonalisation (Ssytrd), the tridiagonal eigenproblem (divide and conquer: a chain of FP32 GEMMs whose length
// depends on deflation), back-transformation (Sormtr).  This file replaces the middle part by two embarrassingly
// parallel kernels:
//   * eigenvalues by Sturm-count multisection (one thread per eigenvalue, three independent pivot chains per thread),
//   * eigenvectors by twisted factorisation (Parlett & Dhillon): forward LDL^T and backward UDU^T pivots of T - lambda I,
//     twist index r = argmin |gamma_r|, one recurrence outwards from r; one thread per eigenvector, no
//     reorthogonalisation - in Float64 the loss of orthogonality is eps64 * |T| / gap, i.e. < 1e-7 for every gap above
//     1e-9 |T|, far tighter than the FP32 tridiagonalisation that produced T.  Eigenvalues closer than that (exact
//     multiplicities: null directions of a binarised or centred matrix) form clusters that are re-solved by inverse
//     iteration with partial pivoting and modified Gram-Schmidt inside the cluster (dstein's scheme), one CTA each.
// All arrays are device arrays; vectors come out as rows of Z (vector-major == column-major n x m for Julia).
#include <algorithm>
#include <cmath>
#include <cfloat>
#include "common.cuh"
#include "eigen.h"
#include "tmp.cuh"

namespace scl {
namespace {

constexpr int kTile = 512;         // elements of d / e^2 staged per shared-memory tile
constexpr int kBisectThreads = 128;
constexpr int kBisectPasses = 28;  // quadrisection: the bracket shrinks 4x per pass, 2^-56 of the Gershgorin range

__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double t = fma(-x, r, 1.0);
  r = fma(r, t, r);
  t = fma(-x, r, 1.0);
  r = fma(r, t, r);
  return r;
}

// info[0] = gl, [1] = gu (Gershgorin, widened as dstebz does), [2] = pivmin, [3] = |T| (max |bound|)
__global__ void __launch_bounds__(1024) k_tri_prepare(const float* __restrict__ d32, const float* __restrict__ e32, int n,
                                                      double* __restrict__ d, double* __restrict__ e,
                                                      double* __restrict__ e2s, double* __restrict__ info) {
  __shared__ double s_lo[32], s_hi[32], s_e2[32];
  double lo = DBL_MAX, hi = -DBL_MAX, me2 = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double di = (double)d32[i];
    const double el = i > 0 ? fabs((double)e32[i - 1]) : 0.0, er = i < n - 1 ? fabs((double)e32[i]) : 0.0;
    d[i] = di;
    if (i < n - 1) e[i] = (double)e32[i];
    e2s[i] = el * el;            // e2s[i] = e_{i-1}^2, e2s[0] = 0: the pivot recurrence reads it at step i
    lo = fmin(lo, di - el - er);
    hi = fmax(hi, di + el + er);
    me2 = fmax(me2, el * el);
  }
  for (int o = 16; o; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    me2 = fmax(me2, __shfl_xor_sync(0xffffffffu, me2, o));
  }
  if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; s_e2[threadIdx.x >> 5] = me2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { lo = fmin(lo, s_lo[w]); hi = fmax(hi, s_hi[w]); me2 = fmax(me2, s_e2[w]); }
    const double tnorm = fmax(fabs(lo), fabs(hi));
    const double pivmin = 1e-290 * fmax(1.0, me2);
    const double fudge = 2.1 * tnorm * DBL_EPSILON * (double)n + 4.2 * pivmin;
    info[0] = lo - fudge;
    info[1] = hi + fudge;
    info[2] = pivmin;
    info[3] = tnorm;
  }
}

// w[k] = k-th smallest eigenvalue (k = k0 + thread), by quadrisection on the Sturm count  #{eigenvalues < x}
__global__ void __launch_bounds__(kBisectThreads) k_bisect(const double* __restrict__ d, const double* __restrict__ e2s, int n,
                                                            const double* __restrict__ info, int k0, int m,
                                                            double* __restrict__ w) {
  __shared__ double sd[kTile], se[kTile];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = k0 + min(t, m - 1);   // surplus threads of the last block shadow the last eigenvalue
  const double pivmin = info[2];
  double lo = info[0], hi = info[1];
  for (int pass = 0; pass < kBisectPasses; ++pass) {
    const double h = 0.25 * (hi - lo);
    const double x1 = lo + h, x2 = lo + 2.0 * h, x3 = lo + 3.0 * h;
    double q1 = 1.0, q2 = 1.0, q3 = 1.0;
    int c1 = 0, c2 = 0, c3 = 0;
    for (int base = 0; base < n; base += kTile) {
      __syncthreads();
      for (int j = threadIdx.x; j < kTile && base + j < n; j += blockDim.x) { sd[j] = d[base + j]; se[j] = e2s[base + j]; }
      __syncthreads();
      const int cnt = min(kTile, n - base);
#pragma unroll 4
      for (int j = 0; j < cnt; ++j) {
        const double a = sd[j], b = se[j];
        q1 = fma(-b, fast_rcp(q1), a - x1);
        q2 = fma(-b, fast_rcp(q2), a - x2);
        q3 = fma(-b, fast_rcp(q3), a - x3);
        if (fabs(q1) < pivmin) q1 = -pivmin;
        if (fabs(q2) < pivmin) q2 = -pivmin;
        if (fabs(q3) < pivmin) q3 = -pivmin;
        c1 += q1 < 0.0;
        c2 += q2 < 0.0;
        c3 += q3 < 0.0;
      }
    }
    if (c1 > k) hi = x1;
    else if (c2 > k) { lo = x1; hi = x2; }
    else if (c3 > k) { lo = x2; hi = x3; }
    else lo = x3;
  }
  if (t < m) w[t] = 0.5 * (lo + hi);
}

// Forward pivots D+ of T - lambda I for B eigenvalues at once; Dp is [n][B] (step-major: coalesced over eigenvalues).
__global__ void __launch_bounds__(128) k_twist_forward(const double* __restrict__ d, const double* __restrict__ e2s, int n,
                                                        const double* __restrict__ info, const double* __restrict__ lam,
                                                        int B, double* __restrict__ Dp) {
  __shared__ double sd[kTile], se[kTile];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = t < B;
  const double x = lam[live ? t : B - 1];
  const double pivmin = info[2];
  double q = 1.0;
  for (int base = 0; base < n; base += kTile) {
    __syncthreads();
    for (int j = threadIdx.x; j < kTile && base + j < n; j += blockDim.x) { sd[j] = d[base + j]; se[j] = e2s[base + j]; }
    __syncthreads();
    const int cnt = min(kTile, n - base);
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      q = fma(-se[j], fast_rcp(q), sd[j] - x);
      if (fabs(q) < pivmin) q = -pivmin;
      if (live) Dp[(size_t)(base + j) * B + t] = q;
    }
  }
}

// Backward pivots D- (stored), gamma_i = D+_i + D-_i - (d_i - lambda), twist index r = argmin |gamma_i|.
__global__ void __launch_bounds__(128) k_twist_backward(const double* __restrict__ d, const double* __restrict__ e2s, int n,
                                                         const double* __restrict__ info, const double* __restrict__ lam,
                                                         int B, const double* __restrict__ Dp, double* __restrict__ Dm,
                                                         int* __restrict__ twist) {
  __shared__ double sd[kTile], se[kTile];   // se[j] = e_{i}^2 for i = base + j, i.e. e2s[i + 1]
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = t < B;
  const int tt = live ? t : B - 1;
  const double x = lam[tt];
  const double pivmin = info[2];
  double q = 1.0, best = DBL_MAX;
  int r = n - 1;
  const int n_tiles = (n + kTile - 1) / kTile;
  for (int tile = n_tiles - 1; tile >= 0; --tile) {
    const int base = tile * kTile;
    __syncthreads();
    for (int j = threadIdx.x; j < kTile && base + j < n; j += blockDim.x) {
      sd[j] = d[base + j];
      se[j] = base + j + 1 < n ? e2s[base + j + 1] : 0.0;
    }
    __syncthreads();
    const int cnt = min(kTile, n - base);
#pragma unroll 4
    for (int j = cnt - 1; j >= 0; --j) {
      const int i = base + j;
      const double a = sd[j] - x;
      q = fma(-se[j], fast_rcp(q), a);
      if (fabs(q) < pivmin) q = -pivmin;
      const double g = fabs(Dp[(size_t)i * B + tt] + q - a);
      if (live) Dm[(size_t)i * B + t] = q;
      if (g < best) { best = g; r = i; }
    }
  }
  if (live) twist[t] = r;
}

// z_r = 1, z_i = -(e_i / D+_i) z_{i+1} (i < r), z_i = -(e_{i-1} / D-_i) z_{i-1} (i > r); z overwrites Dp; nrm2[t] = |z|^2
__global__ void __launch_bounds__(128) k_twist_vector(const double* __restrict__ e, int n, int B, double* __restrict__ Dp,
                                                       const double* __restrict__ Dm, const int* __restrict__ twist,
                                                       double* __restrict__ nrm2) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B) return;
  const int r = twist[t];
  double z = 1.0, s = 1.0;
  for (int i = r - 1; i >= 0; --i) {
    z = -(__ldg(e + i) * fast_rcp(Dp[(size_t)i * B + t])) * z;
    Dp[(size_t)i * B + t] = z;
    s = fma(z, z, s);
  }
  Dp[(size_t)r * B + t] = 1.0;
  z = 1.0;
  for (int i = r + 1; i < n; ++i) {
    z = -(__ldg(e + i - 1) * fast_rcp(Dm[(size_t)i * B + t])) * z;
    Dp[(size_t)i * B + t] = z;
    s = fma(z, z, s);
  }
  nrm2[t] = s;
}

// Z[t][i] = float(Zt[i][t] / |z_t|): 32 x 32 tiles through shared memory; bad[0] counts non-finite norms
__global__ void __launch_bounds__(256) k_transpose_unit(const double* __restrict__ Zt, int n, int B, const double* __restrict__ nrm2,
                                                         float* __restrict__ Z, long long ldz, int* __restrict__ bad) {
  __shared__ double tile[32][33];
  const int i0 = blockIdx.x * 32, t0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int a = ty; a < 32; a += 8) {
    const int i = i0 + a, t = t0 + tx;
    tile[a][tx] = (i < n && t < B) ? Zt[(size_t)i * B + t] : 0.0;
  }
  __syncthreads();
  for (int a = ty; a < 32; a += 8) {
    const int t = t0 + a, i = i0 + tx;
    if (t < B && i < n) {
      const double s = nrm2[t];
      if (i == 0 && !(s > 0.0 && s < DBL_MAX)) atomicAdd(bad, 1);
      Z[(long long)t * ldz + i] = (float)(tile[tx][a] * rsqrt(s));
    }
  }
}

// ---- clusters ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double hash_unit(uint32_t a, uint32_t b) {
  uint64_t h = ((uint64_t)a << 32 | b) * 0x9E3779B97F4A7C15ull;
  h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
  return (double)(h & 0xffffffu) / (double)0x800000 - 1.0;   // [-1, 1)
}

__device__ double block_sum(double v, double* red) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
  return s;
}

// One CTA per cluster [first, first + count): inverse iteration (T - lambda' I) x = b with partial pivoting (sequential,
// thread 0), modified Gram-Schmidt against the cluster's earlier vectors, three rounds per vector.  work: per cluster
// 5 n doubles (sub/diag/super/super2 copies + x); vecs: per cluster count x n doubles.
__global__ void __launch_bounds__(256) k_cluster_fix(const double* __restrict__ d, const double* __restrict__ e, int n,
                                                      const double* __restrict__ info, const double* __restrict__ w,
                                                      const int* __restrict__ cl_first, const int* __restrict__ cl_count,
                                                      const size_t* __restrict__ cl_vec_off, double* __restrict__ work,
                                                      double* __restrict__ vecs, float* __restrict__ Z, long long ldz, int v0) {
  __shared__ double red[8];
  const int c = blockIdx.x, first = cl_first[c], count = cl_count[c];
  double* dl = work + (size_t)c * 5 * n;
  double* dd = dl + n;
  double* du = dd + n;
  double* du2 = du + n;
  double* x = du2 + n;
  double* V = vecs + cl_vec_off[c];
  const double tnorm = info[3], pertol = 10.0 * DBL_EPSILON * fmax(tnorm, DBL_MIN), tiny = DBL_EPSILON * fmax(tnorm, DBL_MIN);
  double prev = 0;
  for (int mbr = 0; mbr < count; ++mbr) {
    double lam = w[first + mbr];
    if (mbr > 0 && lam - prev < pertol) lam = prev + pertol;   // dstein: keep the shifts of a cluster apart
    prev = lam;
    for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] = hash_unit((uint32_t)(first + mbr), (uint32_t)i);
    for (int round = 0; round < 3; ++round) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        dd[i] = d[i] - lam;
        dl[i] = i < n - 1 ? e[i] : 0.0;
        du[i] = i < n - 1 ? e[i] : 0.0;
        du2[i] = 0.0;
      }
      __syncthreads();
      if (threadIdx.x == 0) {   // dgtsv: elimination with row interchanges, then back substitution
        for (int i = 0; i < n - 1; ++i) {
          if (fabs(dd[i]) >= fabs(dl[i])) {
            if (dd[i] == 0.0) dd[i] = tiny;
            const double f = dl[i] / dd[i];
            dd[i + 1] -= f * du[i];
            x[i + 1] -= f * x[i];
            dl[i] = 0.0;
          } else {
            const double f = dd[i] / dl[i];
            dd[i] = dl[i];
            const double tmp = dd[i + 1];
            dd[i + 1] = du[i] - f * tmp;
            if (i < n - 2) { du2[i] = du[i + 1]; du[i + 1] = -f * du[i + 1]; }
            du[i] = tmp;
            const double tb = x[i];
            x[i] = x[i + 1];
            x[i + 1] = tb - f * x[i + 1];
          }
        }
        if (dd[n - 1] == 0.0) dd[n - 1] = tiny;
        x[n - 1] /= dd[n - 1];
        if (n > 1) x[n - 2] = (x[n - 2] - du[n - 2] * x[n - 1]) / dd[n - 2];
        for (int i = n - 3; i >= 0; --i) x[i] = (x[i] - du[i] * x[i + 1] - du2[i] * x[i + 2]) / dd[i];
      }
      __syncthreads();
      // scale first (the solve amplifies by 1/|lambda_j - lam|), then orthogonalise and normalise
      double mx = 0;
      for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmax(mx, fabs(x[i]));
      for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      __syncthreads();
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
      __syncthreads();
      mx = 0;
      for (int q = 0; q < (int)(blockDim.x >> 5); ++q) mx = fmax(mx, red[q]);
      const double inv = mx > 0 && mx < DBL_MAX ? 1.0 / mx : 1.0;
      for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] *= inv;
      __syncthreads();
      for (int p = 0; p < mbr; ++p) {
        const double* vp = V + (size_t)p * n;
        double dot = 0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) dot += x[i] * vp[i];
        dot = block_sum(dot, red);
        for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] -= dot * vp[i];
        __syncthreads();
      }
      double s = 0;
      for (int i = threadIdx.x; i < n; i += blockDim.x) s += x[i] * x[i];
      s = block_sum(s, red);
      const double rn = s > 0 ? rsqrt(s) : 0.0;
      for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] *= rn;
      __syncthreads();
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      V[(size_t)mbr * n + i] = x[i];
      const int row = first + mbr - v0;
      Z[(long long)row * ldz + i] = (float)x[i];
    }
    __syncthreads();
  }
}

__global__ void k_f64_to_f32(const double* __restrict__ a, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)a[i];
}

}  // namespace

// Eigenvalues (all n, ascending, Float64 in w64 and Float32 in w32) of the symmetric tridiagonal matrix (d32, e32) and
// the eigenvectors with ascending indices [v0, v1) as rows of Z (row j - v0, ldz floats apart, unit norm).
// Returns false when the vectors came out non-finite (the caller falls back to the library's one-stage solver).
bool tridiag_eigen(const float* d32, const float* e32, int n, double* w64, float* w32, int v0, int v1, float* Z,
                   long long ldz, cudaStream_t st, TridiagStats* stats) {
  SCL_REQUIRE(n >= 2 && v0 >= 0 && v1 <= n && v0 <= v1, "bad tridiagonal eigenproblem");
  Tmp<double> d(n, st), e(n, st), e2s(n, st), info(4, st);
  count_launches(3);
  k_tri_prepare<<<1, 1024, 0, st>>>(d32, e32, n, d.p, e.p, e2s.p, info.p);
  k_bisect<<<(n + kBisectThreads - 1) / kBisectThreads, kBisectThreads, 0, st>>>(d.p, e2s.p, n, info.p, 0, n, w64);
  k_f64_to_f32<<<(n + 255) / 256, 256, 0, st>>>(w64, n, w32);
  SCL_CUDA(cudaGetLastError());
  const int m = v1 - v0;
  if (stats) *stats = TridiagStats{};
  if (m == 0) return true;
  // vectors in batches: two n x B Float64 work arrays (at most ~4 GB each)
  const int Bmax = (int)std::max<size_t>(128, std::min<size_t>((size_t)m, ((size_t)4 << 30) / ((size_t)n * 8)));
  Tmp<int> bad(1, st);
  SCL_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), st));
  {
    const int B0 = std::min(m, Bmax);
    Tmp<double> Dp((size_t)n * B0, st), Dm((size_t)n * B0, st), nrm2(B0, st);
    Tmp<int> twist(B0, st);
    for (int b0 = 0; b0 < m; b0 += Bmax) {
      const int B = std::min(Bmax, m - b0);
      const int grid = (B + 127) / 128;
      count_launches(4);
      k_twist_forward<<<grid, 128, 0, st>>>(d.p, e2s.p, n, info.p, w64 + v0 + b0, B, Dp.p);
      k_twist_backward<<<grid, 128, 0, st>>>(d.p, e2s.p, n, info.p, w64 + v0 + b0, B, Dp.p, Dm.p, twist.p);
      k_twist_vector<<<grid, 128, 0, st>>>(e.p, n, B, Dp.p, Dm.p, twist.p, nrm2.p);
      k_transpose_unit<<<dim3((n + 31) / 32, (B + 31) / 32), 256, 0, st>>>(Dp.p, n, B, nrm2.p, Z + (long long)b0 * ldz, ldz, bad.p);
      SCL_CUDA(cudaGetLastError());
    }
  }
  // clusters: eigenvalues closer than 1e-9 |T| (twisted factorisation alone leaves their vectors parallel)
  std::vector<double> hw(n);
  double hinfo[4];
  int hbad = 0;
  SCL_CUDA(cudaMemcpyAsync(hw.data(), w64, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  SCL_CUDA(cudaMemcpyAsync(hinfo, info.p, sizeof(hinfo), cudaMemcpyDeviceToHost, st));
  SCL_CUDA(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  SCL_CUDA(cudaStreamSynchronize(st));
  if (hbad) return false;
  const double ctol = 1e-9 * hinfo[3];
  std::vector<int> first, count;
  std::vector<size_t> voff;
  size_t vtot = 0;
  for (int j = v0; j < v1;) {
    int k = j;
    while (k + 1 < v1 && hw[k + 1] - hw[k] <= ctol) ++k;
    if (k > j) {
      first.push_back(j);
      count.push_back(k - j + 1);
      voff.push_back(vtot);
      vtot += (size_t)(k - j + 1) * n;
    }
    j = k + 1;
  }
  if (stats) { stats->clusters = (int)first.size(); stats->clustered = (int)(vtot / (size_t)n); }
  if (first.empty()) return true;
  if (vtot / (size_t)n > 2048) return false;   // a spectrum this degenerate is not what this path is for
  const int nc = (int)first.size();
  Tmp<int> d_first(nc, st), d_count(nc, st);
  Tmp<size_t> d_voff(nc, st);
  Tmp<double> work((size_t)nc * 5 * n, st), vecs(vtot, st);
  SCL_CUDA(cudaMemcpyAsync(d_first.p, first.data(), nc * sizeof(int), cudaMemcpyHostToDevice, st));
  SCL_CUDA(cudaMemcpyAsync(d_count.p, count.data(), nc * sizeof(int), cudaMemcpyHostToDevice, st));
  SCL_CUDA(cudaMemcpyAsync(d_voff.p, voff.data(), nc * sizeof(size_t), cudaMemcpyHostToDevice, st));
  count_launches(1);
  k_cluster_fix<<<nc, 256, 0, st>>>(d.p, e.p, n, info.p, w64, d_first.p, d_count.p, d_voff.p, work.p, vecs.p, Z, ldz, v0);
  SCL_CUDA(cudaGetLastError());
  SCL_CUDA(cudaStreamSynchronize(st));   // the host vectors above are read by the copies
  return true;
}

}  // namespace scl
