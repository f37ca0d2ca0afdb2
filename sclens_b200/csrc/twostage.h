// Two-stage symmetric eigensolver (replaces the one-stage tridiagonalisation inside _get_eigen, src/scLENS.jl:375-382):
//   stage 1  sy2sb.cu     dense -> band (half bandwidth kBand): CholeskyQR panels with Householder reconstruction, two-sided
//                         rank-2b updates - all matrix-matrix work, no matrix-vector products over the trailing matrix
//   stage 2  sb2st.cu     band -> tridiagonal by bulge chasing: one persistent kernel, sweeps pipelined through progress flags
//   (tridiag.cu)          eigenvalues / eigenvectors of T in Float64, as in the one-stage path
//   back     backtrans.cu Z = Q1 Q2 E: stage-2 reflectors in blocks of kBand sweeps x one chase level, stage-1 panels as I - V T V'
// The algebra is restated on the CPU in oracle/two_stage_ref.py (tests/test_two_stage_cpu.py).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdlib>
#include "common.cuh"

namespace scl {

constexpr int kBand = 64;          // half bandwidth of the intermediate band matrix = panel width of stage 1
constexpr int kLdab = 2 * kBand;   // band storage: AB[j * kLdab + (i - j)], 0 <= i - j < 2 kBand (room for the bulge)

struct TwoStageTimes {
  double ms[6] = {0, 0, 0, 0, 0, 0};   // sy2sb, sb2st, tridiagonal eigenproblem, Q2, Q1, copies
};

// stage 1.  A: n x n column-major, leading dimension lda (multiple of 4, 16-byte aligned base), lower triangle read.  On
// return: the band in AB (zero outside the half bandwidth), the panels' reflectors V in A below the band (panel k: columns
// k kBand .. of A, rows (k+1) kBand .. n, unit lower trapezoidal, stored explicitly), their T factors in T1 [panel][kBand^2]
// (row-major).  *d_fail (device flag) is raised when a panel could not be factored (rank-deficient panel).
// Returns the number of panels.
// aux (optional): a second stream and two events for the look-ahead - the next panel is factored there while the main stream
// finishes the trailing update.
struct Sy2sbAux {
  cudaStream_t stream = nullptr;
  cudaEvent_t ready = nullptr, done = nullptr;
};
// half_ok: the matrix passed half_range_ok, so the products of the trailing update may run in split binary16 (engine 2)
int sy2sb_lower(float* A, int n, long long lda, float* AB, float* T1, int* d_fail, cudaStream_t st, const Sy2sbAux* aux = nullptr,
                bool half_ok = false);
bool half_range_ok(const float* A, int n, long long lda, cudaStream_t st);

// stage 2.  AB as above (destroyed).  d (n), e (n - 1).  With keep = true the reflector of sweep s, chase level k is stored at
// V2 + s * ldv2 + k * kBand (kBand floats, v[0] = 1) and its tau at tau2 + s * ldt2 + k.
void sb2st(float* AB, int n, float* d, float* e, bool keep, float* V2, long long ldv2, float* tau2, long long ldt2,
           cudaStream_t st);
inline int sb2st_levels(int n) { return (n + kBand - 1) / kBand + 1; }

// Z: mvec vectors of length n, vector-major (vector v at Z + v * ldz, ldz a multiple of 4).  Z <- Q2 Z.
void apply_q2(const float* V2, long long ldv2, const float* tau2, long long ldt2, int n, float* Z, long long ldz, int mvec,
              cudaStream_t st);
// Z <- Q1 Z with the panels left in A by sy2sb_lower
void apply_q1(const float* A, int n, long long lda, const float* T1, int npanels, float* Z, long long ldz, int mvec,
              cudaStream_t st);

// which engine runs the matrix-matrix products.  Tile engines of sgemm_tile.cuh: 0 = FP32 FMA, 1 = tensor cores in three-term
// TF32 (mma.sync), 2 = tensor cores in split binary16 (mma.sync; operands must be O(1)).  3 (Q1 only): eight panels at a time on
// the tcgen05 GEMM of gemm_umma.cu in split binary16 (backtrans.cu, apply_q1_umma).  SCL_TILE_ENGINE sets stage 1 (default 2, which
// applies only to a matrix whose Frobenius norm passes half_range_ok - otherwise 1 runs; 304 vs 369 ms at n = 20 000), SCL_TILE_ENGINE_Q1 the Q1 back-transformation (default 3; measured at n = 20 000,
// smallest half / all vectors: 83 / 130 ms against 192 / 366 ms for 2, 251 / 474 ms for 1, 252 / 487 ms for 0).
// run-time overrides for the parity tests (scl_debug_set_two_stage): [0] Q2 variant, [1] stage-1 engine, [2] Q1 engine; < 0 = unset
extern std::atomic<int> g_two_stage_override[3];
inline int two_stage_choice(int which, const char* env, int dflt) {
  const int o = g_two_stage_override[which].load(std::memory_order_relaxed);
  if (o >= 0) return o;
  const char* e = getenv(env);
  return e ? atoi(e) : dflt;
}
inline int q2_variant() { return two_stage_choice(0, "SCL_Q2_VARIANT", 2); }
inline int tile_engine_s1() { return two_stage_choice(1, "SCL_TILE_ENGINE", 2); }
inline int tile_engine_q1() { return two_stage_choice(2, "SCL_TILE_ENGINE_Q1", 3); }

// every vector scaled to unit length (after the back-transformation: the tensor-core products of apply_q2 lose ~1e-7 n / 64)
void unit_vectors(float* Z, long long ldz, int n, int mvec, cudaStream_t st);

}  // namespace scl
