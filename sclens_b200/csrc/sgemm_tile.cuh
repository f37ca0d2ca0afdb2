// Register-tiled FP32 tile engine of the two-stage eigensolver (sy2sb.cu, backtrans.cu): a CTA of 256 threads accumulates
//   acc(i, j) += sum_k A(i, k) * B(j, k),   i < 128, j < TN (128 or 64)
// with 8 x (TN/16) outputs per thread, 16-deep k-tiles double-buffered through shared memory and the next tile's global
// loads held in registers while the current one is multiplied.  Operands are described, not copied: three source
// layouts cover every product of the reduction (panel times small matrix, symmetric matrix from its lower triangle times
// panel, rank-2k update, reflector blocks times eigenvector slabs), and anything outside an operand's extents reads as 0,
// so ragged edges need no special kernels.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace scl {
namespace tile {

constexpr int TM = 128, KT = 16, NT = 256, PAD = 4;

struct Opnd {
  const float* p;   // element (i = 0, k = 0)
  long long ld;
  int layout;       // 0: p[i + k*ld]   1: p[k + i*ld]   2: symmetric block kept in its lower triangle: i >= k ? p[i + k*ld] : p[k + i*ld]
  int mn, k;        // extents: i < mn, k < k (everything else is zero)
};

__host__ __device__ inline Opnd opnd(const float* p, long long ld, int layout, long long mn, long long k) {
  Opnd o;
  o.p = p; o.ld = ld; o.layout = layout;
  o.mn = (int)(mn < 0 ? 0 : (mn > 0x3fffffff ? 0x3fffffff : mn));
  o.k = (int)(k < 0 ? 0 : (k > 0x3fffffff ? 0x3fffffff : k));
  return o;
}

template <int W>   // operand tile width (128 or 64): W * KT / 4 float4 per tile, NT threads
struct Frag {
  static constexpr int N = W * KT / 4 / NT;
  float4 r[N];
};

template <int W>
__device__ __forceinline__ void g2r(const Opnd& o, int k0, Frag<W>& f) {
  const int tid = (int)threadIdx.x;
  const bool vec = (((uintptr_t)o.p & 15) == 0) && ((o.ld & 3) == 0);
#pragma unroll
  for (int q = 0; q < Frag<W>::N; ++q) {
    const int idx = tid + q * NT;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (o.layout == 1) {
      const int i = idx / (KT / 4), k = k0 + (idx % (KT / 4)) * 4;
      if (i < o.mn && k < o.k) {
        const float* s = o.p + (long long)i * o.ld + k;
        if (vec && k + 3 < o.k) {
          v = *reinterpret_cast<const float4*>(s);
        } else {
          v.x = s[0];
          if (k + 1 < o.k) v.y = s[1];
          if (k + 2 < o.k) v.z = s[2];
          if (k + 3 < o.k) v.w = s[3];
        }
      }
    } else {
      const int k = k0 + idx / (W / 4), i = (idx % (W / 4)) * 4;
      if (k < o.k && i < o.mn) {
        if (o.layout == 0) {
          const float* s = o.p + (long long)k * o.ld + i;
          if (vec && i + 3 < o.mn) {
            v = *reinterpret_cast<const float4*>(s);
          } else {
            v.x = s[0];
            if (i + 1 < o.mn) v.y = s[1];
            if (i + 2 < o.mn) v.z = s[2];
            if (i + 3 < o.mn) v.w = s[3];
          }
        } else {
          float e[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int ii = i + t;
            e[t] = ii < o.mn ? (ii >= k ? o.p[ii + (long long)k * o.ld] : o.p[k + (long long)ii * o.ld]) : 0.f;
          }
          v = make_float4(e[0], e[1], e[2], e[3]);
        }
      }
    }
    f.r[q] = v;
  }
}

template <int W, int P = PAD>
__device__ __forceinline__ void r2s(const Opnd& o, const Frag<W>& f, float* S /* [KT][W + P] */) {
  const int tid = (int)threadIdx.x;
#pragma unroll
  for (int q = 0; q < Frag<W>::N; ++q) {
    const int idx = tid + q * NT;
    if (o.layout == 1) {
      const int i = idx / (KT / 4), k = (idx % (KT / 4)) * 4;
      S[(k + 0) * (W + P) + i] = f.r[q].x;
      S[(k + 1) * (W + P) + i] = f.r[q].y;
      S[(k + 2) * (W + P) + i] = f.r[q].z;
      S[(k + 3) * (W + P) + i] = f.r[q].w;
    } else {
      const int k = idx / (W / 4), i = (idx % (W / 4)) * 4;
      *reinterpret_cast<float4*>(&S[k * (W + P) + i]) = f.r[q];
    }
  }
}

template <int TN>
struct Acc {
  static constexpr int NJ = TN / 16;   // 8 or 4 columns per thread
  float v[8][NJ];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b < NJ; ++b) v[a][b] = 0.f;
  }
  // tile coordinates of v[a][b] for this thread
  static __device__ __forceinline__ int row(int a) { return ((int)threadIdx.x & 15) * 4 + (a & 3) + (a >> 2) * 64; }
  static __device__ __forceinline__ int col(int b) { return ((int)threadIdx.x >> 4) * 4 + (b & 3) + (b >> 2) * 64; }
};

template <int TN>
struct Smem {
  static constexpr int floats = 2 * KT * (TM + PAD) + 2 * KT * (TN + PAD);
};

// acc += A(:, 0..klen) * B(:, 0..klen)'   (all 256 threads; smem = Smem<TN>::floats floats, 16-byte aligned)
template <int TN>
__device__ __forceinline__ void mac(Acc<TN>& acc, const Opnd& A, const Opnd& B, int klen, float* smem) {
  if (klen <= 0) return;
  float* As = smem;
  float* Bs = smem + 2 * KT * (TM + PAD);
  const int tx = (int)threadIdx.x & 15, ty = (int)threadIdx.x >> 4;
  const int nk = (klen + KT - 1) / KT;
  Frag<TM> fa;
  Frag<TN> fb;
  g2r<TM>(A, 0, fa);
  g2r<TN>(B, 0, fb);
  __syncthreads();   // the previous user of the buffers is done
  r2s<TM>(A, fa, As);
  r2s<TN>(B, fb, Bs);
  __syncthreads();
  for (int t = 0; t < nk; ++t) {
    const int cur = t & 1;
    if (t + 1 < nk) {
      g2r<TM>(A, (t + 1) * KT, fa);
      g2r<TN>(B, (t + 1) * KT, fb);
    }
    const float* as = As + cur * KT * (TM + PAD);
    const float* bs = Bs + cur * KT * (TN + PAD);
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&as[k * (TM + PAD) + tx * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&as[k * (TM + PAD) + 64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[Acc<TN>::NJ];
      {
        const float4 b0 = *reinterpret_cast<const float4*>(&bs[k * (TN + PAD) + ty * 4]);
        bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
        if constexpr (TN == 128) {
          const float4 b1 = *reinterpret_cast<const float4*>(&bs[k * (TN + PAD) + 64 + ty * 4]);
          bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
        }
      }
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < Acc<TN>::NJ; ++b) acc.v[a][b] = fmaf(av[a], bv[b], acc.v[a][b]);
    }
    if (t + 1 < nk) {
      r2s<TM>(A, fa, As + (cur ^ 1) * KT * (TM + PAD));
      r2s<TN>(B, fb, Bs + (cur ^ 1) * KT * (TN + PAD));
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Tensor-core variant of the same engine: mma.sync m16n8k8 in error-compensated TF32.  x = hi + lo with hi = tf32(x),
// lo = tf32(x - hi);  a b ~ a_lo b_hi + a_hi b_lo + a_hi b_hi, the three products of a k-step accumulated from a ZERO
// accumulator and added to the running FP32 sum with a rounded add (the tensor core itself adds with truncation: a long
// chain on one accumulator drifts by ~3e-8 per MMA, one-sided).  FP32-level accuracy (2^-21 per product) and FP32 range,
// so no operand needs scaling.  Same operand descriptions, same pipeline; 8 warps as 4 (rows) x 2 (columns), a warp owns
// 32 x TN/2 outputs as 2 x TN/16 accumulator fragments.  Shared-memory rows are padded to a stride of 8 (mod 32) floats, which
// makes every fragment load (4 k-rows x 8 consecutive columns) conflict-free.
__device__ __forceinline__ uint32_t rn_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = rn_tf32(x);
  lo = rn_tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32_zero(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "f"(0.f));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

constexpr int PADT = 8;

template <int TN>
struct AccT {
  static constexpr int NTL = TN / 16;   // 8-column fragments per warp
  float v[2][NTL][4];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < NTL; ++b) v[a][b][0] = v[a][b][1] = v[a][b][2] = v[a][b][3] = 0.f;
  }
  // tile coordinates of v[mt][nt][e] for this thread: rows g and g + 8 of the fragment (e >> 1), columns 2t and 2t + 1 (e & 1)
  static __device__ __forceinline__ int row(int mt, int e) {
    return (((int)threadIdx.x >> 5) & 3) * 32 + mt * 16 + (((int)threadIdx.x & 31) >> 2) + 8 * (e >> 1);
  }
  static __device__ __forceinline__ int col(int nt, int e) {
    return ((int)threadIdx.x >> 7) * (TN / 2) + nt * 8 + 2 * ((int)threadIdx.x & 3) + (e & 1);
  }
};

template <int TN>
struct SmemT {
  static constexpr int floats = 2 * KT * (TM + PADT) + 2 * KT * (TN + PADT);
};

template <int TN>
__device__ __forceinline__ void mac_tc(AccT<TN>& acc, const Opnd& A, const Opnd& Bo, int klen, float* smem) {
  if (klen <= 0) return;
  constexpr int LDA = TM + PADT, LDB = TN + PADT, NTL = TN / 16;
  float* As = smem;
  float* Bs = smem + 2 * KT * LDA;
  const int lane = (int)threadIdx.x & 31, warp = (int)threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int m0 = (warp & 3) * 32 + g, n0 = (warp >> 2) * (TN / 2) + g;
  const int nk = (klen + KT - 1) / KT;
  Frag<TM> fa;
  Frag<TN> fb;
  g2r<TM>(A, 0, fa);
  g2r<TN>(Bo, 0, fb);
  __syncthreads();   // the previous user of the buffers is done
  r2s<TM, PADT>(A, fa, As);
  r2s<TN, PADT>(Bo, fb, Bs);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      g2r<TM>(A, (kt + 1) * KT, fa);
      g2r<TN>(Bo, (kt + 1) * KT, fb);
    }
#pragma unroll
    for (int ks = 0; ks < KT / 8; ++ks) {
      const float* as = As + cur * KT * LDA + (8 * ks + t) * LDA + m0;
      const float* bs = Bs + cur * KT * LDB + (8 * ks + t) * LDB + n0;
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        split_tf32(as[mt * 16], ah[mt][0], al[mt][0]);
        split_tf32(as[mt * 16 + 8], ah[mt][1], al[mt][1]);
        split_tf32(as[4 * LDA + mt * 16], ah[mt][2], al[mt][2]);
        split_tf32(as[4 * LDA + mt * 16 + 8], ah[mt][3], al[mt][3]);
      }
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        uint32_t bh[2], bl[2];
        split_tf32(bs[nt * 8], bh[0], bl[0]);
        split_tf32(bs[4 * LDB + nt * 8], bh[1], bl[1]);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          float d[4];
          mma_tf32_zero(d, al[mt], bh);
          mma_tf32(d, ah[mt], bl);
          mma_tf32(d, ah[mt], bh);
          acc.v[mt][nt][0] += d[0]; acc.v[mt][nt][1] += d[1]; acc.v[mt][nt][2] += d[2]; acc.v[mt][nt][3] += d[3];
        }
      }
    }
    if (kt + 1 < nk) {
      r2s<TM, PADT>(A, fa, As + (cur ^ 1) * KT * LDA);
      r2s<TN, PADT>(Bo, fb, Bs + (cur ^ 1) * KT * LDB);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Split-binary16 variant (mma.sync m16n8k16): x = hi + lo / 2048 with hi = half(x), lo = half((x - hi) 2048) - the same 22
// significant bits as the TF32 pair at twice the contraction depth per instruction (scripts/mma_rate.cu: both issue once per
// 8 cycles and sub-partition on B200).  For operands that are O(1) or smaller in magnitude (eigenvector slabs, reflector
// panels and their T products; anything above 65504 would overflow).  The conversion happens once per element, on the way
// into shared memory (hi and lo planes); fragments are fetched with ldmatrix - plain from a tile stored [row][k]
// (kind 1: the operand is k-contiguous in global memory), transposed from a tile stored [k][row] (kind 0: row-contiguous,
// or the mirrored diagonal block).  Row pitches of 48 and 272 (144) bytes keep the eight rows of an ldmatrix on distinct banks.
// Same accumulator layout (AccT), same pipeline.
constexpr int KPH = 24;   // halves per row of a [row][k] tile (16 + 8 pad)

template <int TN>
struct SmemH {
  static constexpr int plane_a = TM * KPH, plane_b = TN * KPH;   // the larger of the two storage kinds
  static constexpr int halves = 2 * 2 * (plane_a + plane_b);     // two stages, hi + lo
};

__device__ __forceinline__ void cvt_hl4(const float4& v, uint2& h, uint2& l) {
  const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
  const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
  const __half2 l0 = __floats2half2_rn((v.x - f0.x) * 2048.f, (v.y - f0.y) * 2048.f);
  const __half2 l1 = __floats2half2_rn((v.z - f1.x) * 2048.f, (v.w - f1.y) * 2048.f);
  h.x = *reinterpret_cast<const uint32_t*>(&h0); h.y = *reinterpret_cast<const uint32_t*>(&h1);
  l.x = *reinterpret_cast<const uint32_t*>(&l0); l.y = *reinterpret_cast<const uint32_t*>(&l1);
}

// KIND 1: o.layout == 1, tile stored [W][KPH];  KIND 0: o.layout 0 or 2, tile stored [KT][W + 8]
template <int W, int KIND>
__device__ __forceinline__ void r2s_h(const Frag<W>& f, __half* Sh, __half* Sl) {
  const int tid = (int)threadIdx.x;
#pragma unroll
  for (int q = 0; q < Frag<W>::N; ++q) {
    const int idx = tid + q * NT;
    uint2 h, l;
    cvt_hl4(f.r[q], h, l);
    int off;
    if (KIND == 1) off = (idx / (KT / 4)) * KPH + (idx % (KT / 4)) * 4;
    else off = (idx / (W / 4)) * (W + 8) + (idx % (W / 4)) * 4;
    *reinterpret_cast<uint2*>(Sh + off) = h;
    *reinterpret_cast<uint2*>(Sl + off) = l;
  }
}

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_f16_zero(float (&d)[4], const uint32_t (&a)[4], const uint32_t* b) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "f"(0.f));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], const uint32_t* b) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// acc += A(:, 0..klen) * B(:, 0..klen)';  KA / KB: storage kind of the operand (1 for layout 1, 0 for layouts 0 and 2);
// smem: SmemH<TN>::halves halves, 16-byte aligned
template <int TN, int KA, int KB>
__device__ __forceinline__ void mac_h(AccT<TN>& acc, const Opnd& A, const Opnd& Bo, int klen, __half* smem) {
  if (klen <= 0) return;
  constexpr int NTL = TN / 16;
  constexpr int PA = SmemH<TN>::plane_a, PB = SmemH<TN>::plane_b, STG = 2 * (PA + PB);   // a stage: A hi, A lo, B hi, B lo
  const int lane = (int)threadIdx.x & 31, warp = (int)threadIdx.x >> 5;
  const int mi = lane >> 3, lr = lane & 7;
  const int m0 = (warp & 3) * 32, n0 = (warp >> 2) * (TN / 2);
  // lane offsets (halves) of the ldmatrix rows.  A: registers a0..a3 = (rows 0-7, k 0-7), (rows 8-15, k 0-7), (rows 0-7, k 8-15),
  // (rows 8-15, k 8-15).  B: (hi: k 0-7, k 8-15), (lo: k 0-7, k 8-15) of one 8-column fragment.
  const int offa = KA == 1 ? (m0 + (mi & 1) * 8 + lr) * KPH + (mi >> 1) * 8 : ((mi >> 1) * 8 + lr) * (TM + 8) + m0 + (mi & 1) * 8;
  const int offb = (mi >> 1) * PB + (KB == 1 ? (n0 + lr) * KPH + (mi & 1) * 8 : ((mi & 1) * 8 + lr) * (TN + 8) + n0);
  const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem);
  const int nk = (klen + KT - 1) / KT;
  Frag<TM> fa;
  Frag<TN> fb;
  g2r<TM>(A, 0, fa);
  g2r<TN>(Bo, 0, fb);
  __syncthreads();   // the previous user of the buffers is done
  r2s_h<TM, KA>(fa, smem, smem + PA);
  r2s_h<TN, KB>(fb, smem + 2 * PA, smem + 2 * PA + PB);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      g2r<TM>(A, (kt + 1) * KT, fa);
      g2r<TN>(Bo, (kt + 1) * KT, fb);
    }
    const uint32_t sa = s0 + 2u * (uint32_t)(cur * STG + offa), sb = s0 + 2u * (uint32_t)(cur * STG + 2 * PA + offb);
    uint32_t ah[2][4], al[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const uint32_t d = 2u * (uint32_t)(KA == 1 ? mt * 16 * KPH : mt * 16);
      if (KA == 1) {
        ldsm_x4(ah[mt], sa + d);
        ldsm_x4(al[mt], sa + d + 2u * PA);
      } else {
        ldsm_x4_t(ah[mt], sa + d);
        ldsm_x4_t(al[mt], sa + d + 2u * PA);
      }
    }
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) {
      uint32_t b[4];   // hi b0, hi b1, lo b0, lo b1
      if (KB == 1) ldsm_x4(b, sb + 2u * (uint32_t)(nt * 8 * KPH));
      else ldsm_x4_t(b, sb + 2u * (uint32_t)(nt * 8));
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        float dm[4], dc[4];
        mma_f16_zero(dm, ah[mt], b);
        mma_f16_zero(dc, al[mt], b);
        mma_f16(dc, ah[mt], b + 2);
        acc.v[mt][nt][0] += fmaf(dc[0], 1.f / 2048.f, dm[0]); acc.v[mt][nt][1] += fmaf(dc[1], 1.f / 2048.f, dm[1]);
        acc.v[mt][nt][2] += fmaf(dc[2], 1.f / 2048.f, dm[2]); acc.v[mt][nt][3] += fmaf(dc[3], 1.f / 2048.f, dm[3]);
      }
    }
    if (kt + 1 < nk) {
      __half* nx = smem + (cur ^ 1) * STG;
      r2s_h<TM, KA>(fa, nx, nx + PA);
      r2s_h<TN, KB>(fb, nx + 2 * PA, nx + 2 * PA + PB);
    }
    __syncthreads();
  }
}

// engine selector for kernels that are compiled for both tensor-core engines: ENG 1 = three-term TF32, 2 = split binary16
template <int ENG, int TN>
struct SmemE {
  static constexpr int bytes = ENG == 2 ? SmemH<TN>::halves * 2 : SmemT<TN>::floats * 4;
};
template <int ENG, int TN, int KA, int KB>
__device__ __forceinline__ void mac_e(AccT<TN>& acc, const Opnd& A, const Opnd& Bo, int klen, void* smem) {
  if constexpr (ENG == 2) mac_h<TN, KA, KB>(acc, A, Bo, klen, reinterpret_cast<__half*>(smem));
  else mac_tc<TN>(acc, A, Bo, klen, reinterpret_cast<float*>(smem));
}

}  // namespace tile
}  // namespace scl
