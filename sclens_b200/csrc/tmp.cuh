// Stream-ordered temporaries (cudaMallocAsync pool: cached after the first use).
#pragma once
#include "common.cuh"

namespace scl {
template <typename T>
struct Tmp {
  T* p = nullptr;
  cudaStream_t st;
  Tmp(size_t n, cudaStream_t s) : st(s) { SCL_CUDA(cudaMallocAsync(&p, (n ? n : 1) * sizeof(T), s)); }
  Tmp(const Tmp&) = delete;
  Tmp& operator=(const Tmp&) = delete;
  ~Tmp() {
    if (p) cudaFreeAsync(p, st);
  }
};
}  // namespace scl
