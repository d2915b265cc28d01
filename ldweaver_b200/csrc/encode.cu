// Alignment encoding on the device (HBM-bound byte work).
//
// Replaces the counting / filtering / one-hot extraction loops of the reference's
// src/getACGTNsites.cpp (extractAlnParam :50-85 + :104-166, extractSNPs :222-267) and
// src/ACGTN2num_parallel.cpp:18-41.  gz inflate + FASTA tokenising stay on the host (fasta_host.cpp).
#include "encode.h"

namespace ldw {

// [Aa]->0 [Cc]->1 [Gg]->2 [Tt]->3, anything else -> 4  (src/getACGTNsites.cpp:59-69, quirk Q8)
__device__ __forceinline__ int classify_char(unsigned c) {
  unsigned u = c | 0x20u;  // only maps 'A'..'Z' onto 'a'..'z'; no other byte aliases onto a,c,g,t
  return u == 'a' ? 0 : u == 'c' ? 1 : u == 'g' ? 2 : u == 't' ? 3 : 4;
}

// Column histogram.  Each thread owns 4 adjacent columns and walks a slice of the rows, so a warp reads
// 128 contiguous bytes per row; per-thread counters are flushed with one atomic per (column, class).
// counts: int32 [L x 5] (== 5 x L column-major).  The "other" class is rows - sum(ACGT).
__global__ void __launch_bounds__(256) column_count_kernel(const uint8_t* __restrict__ aln, int64_t S, int64_t L,
                                                           int rows_per_slice, int32_t* counts) {
  int64_t j0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (j0 >= L) return;
  int64_t r0 = (int64_t)blockIdx.y * rows_per_slice;
  int64_t r1 = r0 + rows_per_slice < S ? r0 + rows_per_slice : S;
  int c[4][4];
#pragma unroll
  for (int q = 0; q < 4; q++)
#pragma unroll
    for (int a = 0; a < 4; a++) c[q][a] = 0;
  const bool vec = ((L & 3) == 0) && ((reinterpret_cast<uintptr_t>(aln) & 3) == 0) && (j0 + 3 < L);
  for (int64_t s = r0; s < r1; s++) {
    uint32_t w;
    if (vec) {
      w = *reinterpret_cast<const uint32_t*>(aln + s * L + j0);
    } else {
      w = 0;
#pragma unroll
      for (int q = 0; q < 4; q++)
        if (j0 + q < L) w |= (uint32_t)aln[s * L + j0 + q] << (8 * q);
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
      int k = classify_char((w >> (8 * q)) & 0xFF);
#pragma unroll
      for (int a = 0; a < 4; a++) c[q][a] += (k == a);
    }
  }
  int rows = (int)(r1 - r0);
#pragma unroll
  for (int q = 0; q < 4; q++) {
    if (j0 + q >= L) break;
    int sum = 0;
#pragma unroll
    for (int a = 0; a < 4; a++) {
      if (c[q][a]) atomicAdd(&counts[(j0 + q) * 5 + a], c[q][a]);
      sum += c[q][a];
    }
    if (rows - sum) atomicAdd(&counts[(j0 + q) * 5 + 4], rows - sum);
  }
}

// Site filter (src/getACGTNsites.cpp:104-166; quirk Q9).  flag[j] = 1 if column j is retained.
__global__ void site_filter_kernel(const int32_t* __restrict__ counts, int64_t L, int n, int filter, double gap_thresh,
                                   int min_maf, uint8_t* flag, int32_t* block_counts) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int keep = 0;
  if (j < L) {
    int c[5];
#pragma unroll
    for (int a = 0; a < 5; a++) c[a] = counts[j * 5 + a];
    int present = (c[0] > 0) + (c[1] > 0) + (c[2] > 0) + (c[3] > 0);
    if (present > 1 && ((double)c[4] / (double)n < gap_thresh)) {
      if (filter == 0) {
        // second largest of the four non-gap counts must exceed min_maf (:119-123)
        int hi = max(c[0], c[1]), lo = min(c[0], c[1]);
        int hi2 = max(c[2], c[3]), lo2 = min(c[2], c[3]);
        int second = max(min(hi, hi2), max(lo, lo2));
        keep = second > min_maf;
      } else {
        int mx = max(max(max(c[0], c[1]), max(c[2], c[3])), c[4]);
        keep = mx <= min_maf;  // :153
      }
    }
    flag[j] = (uint8_t)keep;
  }
  int total = __syncthreads_count(keep);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}

// Ordered compaction of the retained columns into 1-based POS.
__global__ void site_compact_kernel(const uint8_t* __restrict__ flag, int64_t L, const int32_t* __restrict__ block_off,
                                    int32_t* pos) {
  __shared__ int warp_tot[32];
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int keep = (j < L) ? flag[j] : 0;
  unsigned b = __ballot_sync(0xffffffffu, keep);
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_tot[warp] = __popc(b);
  __syncthreads();
  int base = block_off[blockIdx.x];
  for (int w = 0; w < warp; w++) base += warp_tot[w];
  if (keep) pos[base + __popc(b & ((1u << lane) - 1))] = (int32_t)(j + 1);
}

__global__ void counts_to_double_kernel(const int32_t* __restrict__ in, int64_t n, double* out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (double)in[i];
}

// Gather retained columns and classify: codes[k][s] = class(aln[s][pos[k]-1]).  32x32 tiles through shared
// memory so the writes are contiguous along sequences.
__global__ void __launch_bounds__(1024) extract_codes_kernel(const uint8_t* __restrict__ aln, int64_t S, int64_t L,
                                                             const int32_t* __restrict__ pos, int64_t n, uint8_t* codes) {
  __shared__ uint8_t tile[32][33];
  int64_t k0 = (int64_t)blockIdx.x * 32, s0 = (int64_t)blockIdx.y * 32;
  int tx = threadIdx.x, ty = threadIdx.y;
  int64_t k = k0 + tx, s = s0 + ty;
  if (k < n && s < S) tile[tx][ty] = (uint8_t)classify_char(aln[s * L + (pos[k] - 1)]);
  __syncthreads();
  k = k0 + ty;
  s = s0 + tx;
  if (k < n && s < S) codes[k * S + s] = tile[ty][tx];
}

// src/ACGTN2num_parallel.cpp:18-41: uppercase A/C/G/T, and 'N' or '-' -> row 4; anything else untouched.
__global__ void acgtn2num_kernel(double* nv, const char* __restrict__ ref, int64_t n) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  char cc = ref[c];
  int row = cc == 'A' ? 0 : cc == 'C' ? 1 : cc == 'G' ? 2 : cc == 'T' ? 3 : (cc == 'N' || cc == '-') ? 4 : -1;
  if (row >= 0) nv[c * 5 + row] = 0.0;
}

int column_counts_device(cudaStream_t st, const uint8_t* d_aln, int64_t S, int64_t L, int32_t* d_counts) {
  LDW_CUDA(cudaMemsetAsync(d_counts, 0, (size_t)L * 5 * 4, st));
  int slices = (int)((S + 63) / 64);
  if (slices > 64) slices = 64;
  if (slices < 1) slices = 1;
  int rps = (int)((S + slices - 1) / slices);
  dim3 grid((unsigned)((L + 4 * 256 - 1) / (4 * 256)), (unsigned)slices);
  column_count_kernel<<<grid, 256, 0, st>>>(d_aln, S, L, rps, d_counts);
  LDW_CUDA(cudaGetLastError());
  return 0;
}

int site_filter_device(cudaStream_t st, const int32_t* d_counts, int64_t L, int nseq, int filter, double gap_thresh,
                       double maf_thresh, int32_t* d_pos, int64_t* n_out) {
  // int truncation exactly as the reference's `int min_maf = n*maf_thresh` / `n*(1-maf_thresh)` (:105, :136)
  int min_maf = filter == 0 ? (int)(nseq * maf_thresh) : (int)(nseq * (1 - maf_thresh));
  int nb = (int)((L + 255) / 256);
  DevBuf flag, bc, bo, tot;
  LDW_TRY(flag.alloc((size_t)L));
  LDW_TRY(bc.alloc((size_t)nb * 4));
  LDW_TRY(bo.alloc((size_t)nb * 4));
  LDW_TRY(tot.alloc(4));
  site_filter_kernel<<<nb, 256, 0, st>>>(d_counts, L, nseq, filter, gap_thresh, min_maf, flag.as<uint8_t>(), bc.as<int32_t>());
  LDW_CUDA(cudaGetLastError());
  LDW_TRY(exclusive_scan_i32(st, bc.as<int32_t>(), nb, bo.as<int32_t>(), tot.as<int32_t>()));
  site_compact_kernel<<<nb, 256, 0, st>>>(flag.as<uint8_t>(), L, bo.as<int32_t>(), d_pos);
  LDW_CUDA(cudaGetLastError());
  int32_t total = 0;
  LDW_CUDA(cudaMemcpyAsync(&total, tot.p, 4, cudaMemcpyDeviceToHost, st));
  LDW_CUDA(cudaStreamSynchronize(st));
  *n_out = total;
  return 0;
}

int counts_to_double_device(cudaStream_t st, const int32_t* d_in, int64_t n, double* d_out) {
  if (n == 0) return 0;
  counts_to_double_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_in, n, d_out);
  LDW_CUDA(cudaGetLastError());
  return 0;
}

int extract_codes_device(cudaStream_t st, const uint8_t* d_aln, int64_t S, int64_t L, const int32_t* d_pos, int64_t n,
                         uint8_t* d_codes) {
  if (n == 0 || S == 0) return 0;
  dim3 grid((unsigned)((n + 31) / 32), (unsigned)((S + 31) / 32));
  extract_codes_kernel<<<grid, dim3(32, 32), 0, st>>>(d_aln, S, L, d_pos, n, d_codes);
  LDW_CUDA(cudaGetLastError());
  return 0;
}

int acgtn2num_device(cudaStream_t st, double* d_nv, const char* d_ref, int64_t n) {
  if (n == 0) return 0;
  acgtn2num_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_nv, d_ref, n);
  LDW_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ldw
