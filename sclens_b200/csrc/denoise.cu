// get_denoised_df (src/scLENS.jl:889-931) on the device: the robust-signal reconstruction
//   d = pca_n1 * gene_basis[sig_id, :] * sqrt(M)                 (:890-911, Float32 GEMM, inner dimension r = #robust signals)
// pushed back through the recorded normalisation (:913-927)
//   out_ij = mean(TGC) * max(exp(((d_ij + c_j) * l_i / mean(l)) * sigma_j + ybar_j) - 1, 0) / rowsum_i .
// r is a handful, so this is not a tensor-core GEMM: the N x M Float64 output (10.9 GB at 68k x 20k) is the traffic and
// Float64 exp the arithmetic.  A CTA owns 32 consecutive cells (contiguous inside every gene column of the column-major
// output); its eight warps stride over the genes, so a warp's store of one gene is 256 contiguous bytes.  Two sweeps
// over the genes: row sums (fixed-order combination of the eight partial sums), then the normalised values - the
// reconstruction is recomputed rather than held (r FMAs per element).
#include "common.cuh"
#include "tmp.cuh"

namespace scl {

template <typename OutT>
__global__ void __launch_bounds__(256)
k_denoise(const float* __restrict__ A /* [r][N] = pca_n1 column-major */, const float* __restrict__ G /* [M][r] = g_mat column-major */,
          int r, int N, int M, const double* __restrict__ cent, const double* __restrict__ sigma, const double* __restrict__ ybar,
          const double* __restrict__ l2, double inv_mean_l, double sqrtM, double mean_tgc, OutT* __restrict__ out) {
  extern __shared__ float smA[];   // [r][32]
  __shared__ double part[8][32];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  const bool ok = i < N;
  for (int q = grp; q < r; q += 8) smA[q * 32 + lane] = ok ? A[(size_t)q * N + i] : 0.f;
  __syncthreads();
  const double si = ok ? l2[i] * inv_mean_l : 0.0;   // norm_tgc / mean(norm_tgc)  (:922)
  auto value = [&](int j) {
    const float* g = G + (size_t)j * r;
    float d = 0.f;
    for (int q = 0; q < r; ++q) d = fmaf(smA[q * 32 + lane], g[q], d);
    const double r3 = (((double)d * sqrtM + cent[j]) * si) * sigma[j] + ybar[j];   // :921-923
    const double v = exp(r3) - 1.0;                                               // :924
    return v < 0.0 ? 0.0 : v;                                                      // :925
  };
  double sum = 0;
  for (int j = grp; j < M; j += 8) sum += value(j);
  part[grp][lane] = sum;
  __syncthreads();
  double total = 0;
#pragma unroll
  for (int g = 0; g < 8; ++g) total += part[g][lane];
  if (!ok) return;
  const double scale = mean_tgc / total;                                           // :926-927
  for (int j = grp; j < M; j += 8) out[(size_t)j * N + i] = (OutT)(value(j) * scale);
}

// all pointers are device pointers; out is N x M column-major (Float64 or Float32)
void denoise(const float* dA, const float* dG, int r, int N, int M, const double* cent, const double* sigma,
             const double* ybar, const double* l2, double mean_l, double mean_tgc, void* d_out, bool out_f32,
             cudaStream_t st) {
  SCL_REQUIRE(r >= 1 && r <= 1024, "number of robust signals out of range for the denoising kernel");
  count_launches(1);
  const size_t smem = (size_t)r * 32 * sizeof(float);
  const int grid = (N + 31) / 32;
  const double sqrtM = sqrt((double)M);
  if (out_f32) {
    SCL_CUDA(cudaFuncSetAttribute(k_denoise<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_denoise<float><<<grid, 256, smem, st>>>(dA, dG, r, N, M, cent, sigma, ybar, l2, 1.0 / mean_l, sqrtM, mean_tgc, (float*)d_out);
  } else {
    SCL_CUDA(cudaFuncSetAttribute(k_denoise<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_denoise<double><<<grid, 256, smem, st>>>(dA, dG, r, N, M, cent, sigma, ybar, l2, 1.0 / mean_l, sqrtM, mean_tgc, (double*)d_out);
  }
  SCL_CUDA(cudaGetLastError());
}

}  // namespace scl
