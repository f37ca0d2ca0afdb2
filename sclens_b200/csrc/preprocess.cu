// QC on the device: scLENS.preprocess (src/scLENS.jl:160-236) - SURVEY.md 8(f) rank 3.  The reference filters cells and
// genes with sparse column / row reductions on the host, drops genes that became empty, and orders the survivors by a
// stable sortperm of their Float32 mean (:224).  Indices are part of the parity contract (bit-exact), so every Float32 sum is
// accumulated sequentially in storage order exactly as SparseArrays' column / row reductions do; the kernels are one thread
// per line for that reason (QC runs once per data set; it is not a bandwidth kernel).
#include <algorithm>
#include "common.cuh"
#include "handle.h"
#include "tmp.cuh"

namespace scl {
namespace {

// per gene: stored entries (all non-zero in a canonical matrix) and their Float32 sum in row order (:184-189)
__global__ void k_qc_gene(const uint32_t* __restrict__ colptr, const float* __restrict__ val, int M, double min_tp_g,
                          double max_tp_g, int min_cells, uint8_t* __restrict__ fg) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  float s = 0.f;
  for (uint32_t t = colptr[j]; t < colptr[j + 1]; ++t) s += val[t];
  const int cnt = (int)(colptr[j + 1] - colptr[j]);
  fg[j] = ((double)s > min_tp_g && (double)s < max_tp_g && cnt >= min_cells) ? 1 : 0;
}

// per cell: genes, Float32 sums of all / mitochondrial / ribosomal counts in column order, the six conditions of :191-216
__global__ void k_qc_cell(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ colidx, const float* __restrict__ rval,
                          const uint8_t* __restrict__ gene_flags, int N, scl_qc_params p, uint8_t* __restrict__ fc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float s = 0.f, sm = 0.f, sr = 0.f;
  for (uint32_t t = rowptr[i]; t < rowptr[i + 1]; ++t) {
    const float v = rval[t];
    const uint8_t f = gene_flags[colidx[t]];
    s += v;
    if (f & 1) sm += v;
    if (f & 2) sr += v;
  }
  const int cnt = (int)(rowptr[i + 1] - rowptr[i]);
  bool keep = (double)s > p.min_tp_c && (double)s < p.max_tp_c && cnt >= p.min_genes_per_cell;
  if (p.mito_percent != 0) keep = keep && (double)(sm / s) < p.mito_percent / 100.0;     // Float32 ratio, strict (:201)
  if (p.ribo_percent != 0) keep = keep && (double)(sr / s) < p.ribo_percent / 100.0;
  if (p.max_genes_per_cell != 0) keep = keep && cnt < p.max_genes_per_cell;
  fc[i] = keep ? 1 : 0;
}

// exclusive scan of a 0/1 mask in one block (n up to a few million): pos[i] = number of ones before i, pos[n] = total
__global__ void __launch_bounds__(1024) k_mask_scan(const uint8_t* __restrict__ mask, int n, int* __restrict__ pos) {
  __shared__ int part[1024];
  const int tid = threadIdx.x, chunk = (n + 1023) / 1024, b = tid * chunk, e = min(n, b + chunk);
  int s = 0;
  for (int i = b; i < e; ++i) s += mask[i];
  part[tid] = s;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const int v = tid >= off ? part[tid - off] : 0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  int run = tid ? part[tid - 1] : 0;
  for (int i = b; i < e; ++i) {
    pos[i] = run;
    run += mask[i];
  }
  if (tid == 1023) pos[n] = part[1023];
}

// per kept gene over the kept cells: entries, Float32 sum (:220) and mean (:224); survivors = fg && sum != 0
__global__ void k_qc_gene_kept(const uint32_t* __restrict__ colptr, const uint32_t* __restrict__ rowval, const float* __restrict__ val,
                               const uint8_t* __restrict__ fc, int M, int n_cells_kept, uint8_t* __restrict__ fg,
                               float* __restrict__ mean, uint32_t* __restrict__ kept_cnt) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  float s = 0.f;
  uint32_t c = 0;
  if (fg[j]) {
    for (uint32_t t = colptr[j]; t < colptr[j + 1]; ++t)
      if (fc[rowval[t]]) { s += val[t]; ++c; }
  }
  kept_cnt[j] = c;
  mean[j] = s / (float)n_cells_kept;
  if (!(s != 0.f)) fg[j] = 0;
}

__global__ void k_compact_ids(const uint8_t* __restrict__ mask, const int* __restrict__ pos, int n, int32_t* __restrict__ ids) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && mask[i]) ids[pos[i]] = i;
}

// stable ascending order of the survivors' means (:224): rank = number of survivors that sort before this one
__global__ void __launch_bounds__(256) k_rank_genes(const int32_t* __restrict__ ids, const float* __restrict__ mean, int G,
                                                    int32_t* __restrict__ order, uint32_t* __restrict__ out_cnt,
                                                    const uint32_t* __restrict__ kept_cnt) {
  __shared__ float sm[256];
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  const float ma = a < G ? mean[ids[a]] : 0.f;
  int rank = 0;
  for (int base = 0; base < G; base += 256) {
    __syncthreads();
    if (base + (int)threadIdx.x < G) sm[threadIdx.x] = mean[ids[base + threadIdx.x]];
    __syncthreads();
    const int cnt = min(256, G - base);
    for (int q = 0; q < cnt; ++q) {
      const float mb = sm[q];
      rank += (mb < ma) || (mb == ma && base + q < a);
    }
  }
  if (a < G) {
    order[rank] = ids[a];
    out_cnt[rank] = kept_cnt[ids[a]];
  }
}

// output column p = gene order[p] restricted to the kept cells, rows renumbered (one warp per column)
__global__ void __launch_bounds__(256) k_emit_filtered(const uint32_t* __restrict__ colptr, const uint32_t* __restrict__ rowval,
                                                       const float* __restrict__ val, const uint8_t* __restrict__ fc,
                                                       const int* __restrict__ cell_pos, const int32_t* __restrict__ order, int G,
                                                       const uint32_t* __restrict__ out_colptr, uint32_t* __restrict__ out_row,
                                                       float* __restrict__ out_val) {
  const int lane = threadIdx.x & 31;
  for (int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < G; p += (gridDim.x * blockDim.x) >> 5) {
    const int j = order[p];
    uint32_t dst = out_colptr[p];
    for (uint32_t t0 = colptr[j]; t0 < colptr[j + 1]; t0 += 32) {
      const uint32_t t = t0 + lane;
      const bool ok = t < colptr[j + 1] && fc[rowval[t]];
      const uint32_t m = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const uint32_t o = dst + __popc(m & ((1u << lane) - 1u));
        out_row[o] = (uint32_t)cell_pos[rowval[t]];
        out_val[o] = val[t];
      }
      dst += __popc(m);
    }
  }
}

__global__ void __launch_bounds__(1024) k_scan_u32(const uint32_t* __restrict__ in, int n, uint32_t* __restrict__ out) {
  __shared__ uint32_t part[1024];
  const int tid = threadIdx.x, chunk = (n + 1023) / 1024, b = tid * chunk, e = min(n, b + chunk);
  uint32_t s = 0;
  for (int i = b; i < e; ++i) s += in[i];
  part[tid] = s;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const uint32_t v = tid >= off ? part[tid - off] : 0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  uint32_t run = tid ? part[tid - 1] : 0;
  for (int i = b; i < e; ++i) {
    const uint32_t v = in[i];
    out[i] = run;
    run += v;
  }
  if (tid == 1023) out[n] = part[1023];
}

}  // namespace

// X: raw counts (canonical CSC + CSR mirror).  gene_flags[j]: bit 0 = mitochondrial (r"^(?i)mt-." :196), bit 1 = ribosomal
// (r"^(?i)RP[SL]." :197) - the regular expressions run over the gene names on the host.  Returns false when no cell or no
// gene survives (:231-234).  fc_idx: surviving cells, ascending; gene_idx: surviving genes in output order.
bool preprocess_device(const SpMat& X, const uint8_t* d_gene_flags, const scl_qc_params& p, std::vector<int32_t>& fc_idx,
                       std::vector<int32_t>& gene_idx, SpMat& out, cudaStream_t st) {
  const int N = X.N, M = X.M;
  Tmp<uint8_t> fg(M, st), fc(N, st);
  Tmp<int> cell_pos(N + 1, st), gene_pos(M + 1, st);
  count_launches(9);
  k_qc_gene<<<(M + 127) / 128, 128, 0, st>>>(X.colptr.p, X.val.p, M, p.min_tp_g, p.max_tp_g, p.min_cells_per_gene, fg.p);
  k_qc_cell<<<(N + 127) / 128, 128, 0, st>>>(X.rowptr.p, X.colidx.p, X.rval.p, d_gene_flags, N, p, fc.p);
  k_mask_scan<<<1, 1024, 0, st>>>(fc.p, N, cell_pos.p);
  k_mask_scan<<<1, 1024, 0, st>>>(fg.p, M, gene_pos.p);
  int n_cells = 0, n_fg = 0;
  SCL_CUDA(cudaMemcpyAsync(&n_cells, cell_pos.p + N, sizeof(int), cudaMemcpyDeviceToHost, st));
  SCL_CUDA(cudaMemcpyAsync(&n_fg, gene_pos.p + M, sizeof(int), cudaMemcpyDeviceToHost, st));
  SCL_CUDA(cudaStreamSynchronize(st));
  fc_idx.clear();
  gene_idx.clear();
  if (n_cells == 0 || n_fg == 0) return false;           // !(any(fc_idx) && any(fg_idx))  (:218)
  Tmp<float> mean(M, st);
  Tmp<uint32_t> kept_cnt(M, st);
  k_qc_gene_kept<<<(M + 127) / 128, 128, 0, st>>>(X.colptr.p, X.rowval.p, X.val.p, fc.p, M, n_cells, fg.p, mean.p, kept_cnt.p);
  k_mask_scan<<<1, 1024, 0, st>>>(fg.p, M, gene_pos.p);   // survivors after the empty-gene drop (:220-221)
  int G = 0;
  SCL_CUDA(cudaMemcpyAsync(&G, gene_pos.p + M, sizeof(int), cudaMemcpyDeviceToHost, st));
  SCL_CUDA(cudaStreamSynchronize(st));
  if (G == 0) return false;
  Tmp<int32_t> ids(G, st), order(G, st), cells(n_cells, st);
  Tmp<uint32_t> out_cnt(G + 1, st);
  k_compact_ids<<<(M + 255) / 256, 256, 0, st>>>(fg.p, gene_pos.p, M, ids.p);
  k_compact_ids<<<(N + 255) / 256, 256, 0, st>>>(fc.p, cell_pos.p, N, cells.p);
  k_rank_genes<<<(G + 255) / 256, 256, 0, st>>>(ids.p, mean.p, G, order.p, out_cnt.p, kept_cnt.p);
  out.N = n_cells;
  out.M = G;
  out.colptr.ensure(G + 1);
  k_scan_u32<<<1, 1024, 0, st>>>(out_cnt.p, G, out.colptr.p);
  uint32_t nnz = 0;
  SCL_CUDA(cudaMemcpyAsync(&nnz, out.colptr.p + G, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  fc_idx.resize(n_cells);
  gene_idx.resize(G);
  SCL_CUDA(cudaMemcpyAsync(fc_idx.data(), cells.p, (size_t)n_cells * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  SCL_CUDA(cudaMemcpyAsync(gene_idx.data(), order.p, (size_t)G * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  SCL_CUDA(cudaStreamSynchronize(st));
  out.nnz = nnz;
  out.rowval.ensure(std::max<size_t>(1, nnz));
  out.val.ensure(std::max<size_t>(1, nnz));
  k_emit_filtered<<<std::min((G + 7) / 8, 148 * 16), 256, 0, st>>>(X.colptr.p, X.rowval.p, X.val.p, fc.p, cell_pos.p, order.p, G,
                                                                  out.colptr.p, out.rowval.p, out.val.p);
  SCL_CUDA(cudaGetLastError());
  build_csr_mirror(out, st);
  SCL_CUDA(cudaStreamSynchronize(st));
  return true;
}

}  // namespace scl
