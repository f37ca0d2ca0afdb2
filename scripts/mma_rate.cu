// Issue rate of the warp-level (legacy) tensor-core path on this GPU: cycles per mma.sync per SM sub-partition for
// m16n8k8 TF32, m16n8k16 FP16 and m16n8k16 BF16 (FP32 accumulate), with 1, 2 and 4 warps per sub-partition and ILP 4.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/mma_rate scripts/mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND>
__global__ void k(float* out, int iters, long long* cyc) {
  float d[4][4];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
  unsigned a0 = threadIdx.x * 2654435761u & 0x3f803f80u, a1 = a0 ^ 0x1000u, a2 = a0 ^ 0x2000u, a3 = a0 ^ 0x3000u, b0 = a0 ^ 0x100u, b1 = a0 ^ 0x200u;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else if (KIND == 1)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMallocManaged(&cyc, 8);
  const int iters = 20000;
  const char* names[3] = {"tf32 m16n8k8", "f16 m16n8k16", "bf16 m16n8k16"};
  for (int kind = 0; kind < 3; ++kind)
    for (int warps = 4; warps <= 16; warps *= 2) {
      for (int rep = 0; rep < 2; ++rep) {
        if (kind == 0) k<0><<<148, warps * 32>>>(out, iters, cyc);
        if (kind == 1) k<1><<<148, warps * 32>>>(out, iters, cyc);
        if (kind == 2) k<2><<<148, warps * 32>>>(out, iters, cyc);
        cudaDeviceSynchronize();
      }
      const double per = (double)*cyc / ((double)iters * 4 * (warps / 4));
      printf("%s  %2d warps/SM (%d per sub-partition): %.2f cycles per mma per sub-partition\n", names[kind], warps, warps / 4, per);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
