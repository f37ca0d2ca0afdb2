#pragma once
#include <cuda_runtime.h>
#include <vector>

namespace scl {

struct Solver {
  struct Impl;
  Impl* impl;
  explicit Solver(cudaStream_t st);
  ~Solver();
  Solver(const Solver&) = delete;
  Solver& operator=(const Solver&) = delete;
  // cusolverDnSsyevd('V'|'N','U'): eigenvalues ascending in dW, eigenvectors overwrite dA (col-major)
  void syevd(float* dA, int n, float* dW, bool vectors, cudaStream_t st);
  void dsyevd_small(double* dA, int n, double* dW, cudaStream_t st);
  int syevdx_smallest(float* dA, int n, float* dW, int iu, cudaStream_t st);
  // timing study of the library's symmetric eigensolvers on an n x n matrix (overwritten); returns milliseconds.
  // mode 0: Ssyevd vectors, 1: Ssyevd values only, 2: Ssyevdx vectors of the il..iu smallest, 3: Xsyevd (64-bit API) vectors
  double bench(float* dA, int n, float* dW, int mode, int il, int iu, cudaStream_t st);
};

int eig_api();   // SCL_EIG_API: bit 0 = Xsyevd for the full solves, bit 1 = Ssyevdx (index range) in the search steps

struct MpFit {
  std::vector<float> L_mp;
  double lambda_c, b_plus, b_minus, gamma, ks_static;
  int iters, n_signal;
  bool pass;
};
// _mp_calculation(L, Lr) + _tw + mp_check; caller passes Lr already without its largest value (:537)
MpFit mp_fit(const float* L, int nL, const float* Lr, int nLr);

}  // namespace scl
