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

// Column histogram (src/getACGTNsites.cpp:50-85), HBM-bound: S*L bytes read once, 20 L bytes of counters written.
//
// A thread owns 16 adjacent columns -- one 128-bit load per row, a warp reads 512 contiguous bytes -- and walks a slice
// of at most 255 rows.  The sixteen bytes are classified four at a time inside their 32-bit words: for each nucleotide
// x, (w | 0x20202020) ^ x_x4 has a zero byte exactly where the column holds x or X, and the zero bytes are counted into
// byte-lane accumulators (4 columns per register, 16 registers: 4 words x 4 nucleotides; 255 rows cannot overflow a
// lane).  ~5 integer instructions per byte and class instead of a compare/select chain per byte.  At the end of the
// slice the lanes are unpacked and added to counts[L x 5] with one red.add per (column, class); "other" = rows - ACGT.
// Rows start at multiples of `pitch` (>= L, a multiple of 16: the upload pads them), so every load is aligned.
__device__ __forceinline__ uint32_t zero_bytes_to_ones(uint32_t t) {
  // 0x01 in every byte lane of t that is zero, 0x00 elsewhere (exact: no carries cross lanes)
  const uint32_t nz = ((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t;  // bit 7 of a lane set <=> lane != 0
  return (~nz & 0x80808080u) >> 7;
}

constexpr int CC_COLS = 16;          // columns per thread
constexpr int CC_MAX_ROWS = 255;     // rows per slice (byte-lane counters)

__global__ void __launch_bounds__(256) column_count_kernel(const uint8_t* __restrict__ aln, int64_t S, int64_t L, int64_t pitch,
                                                           int rows_per_slice, int32_t* counts) {
  const int64_t j0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * CC_COLS;
  if (j0 >= L) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_slice;
  const int64_t r1 = r0 + rows_per_slice < S ? r0 + rows_per_slice : S;
  if (r0 >= r1) return;
  uint32_t acc[4][4];  // [word][nucleotide], four byte lanes each
#pragma unroll
  for (int q = 0; q < 4; q++)
#pragma unroll
    for (int a = 0; a < 4; a++) acc[q][a] = 0;
  const uint4* p = reinterpret_cast<const uint4*>(aln + r0 * pitch + j0);
  const int64_t step = pitch / 16;
  int64_t s = r0;
  // four rows in flight per thread
  for (; s + 4 <= r1; s += 4) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) v[u] = __ldg(p + u * step);
    p += 4 * step;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const uint32_t lw = w[q] | 0x20202020u;  // folds 'A'..'Z' onto 'a'..'z'; no other byte aliases onto a, c, g, t
        acc[q][0] += zero_bytes_to_ones(lw ^ 0x61616161u);
        acc[q][1] += zero_bytes_to_ones(lw ^ 0x63636363u);
        acc[q][2] += zero_bytes_to_ones(lw ^ 0x67676767u);
        acc[q][3] += zero_bytes_to_ones(lw ^ 0x74747474u);
      }
    }
  }
  for (; s < r1; s++) {
    const uint4 v = __ldg(p);
    p += step;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const uint32_t lw = w[q] | 0x20202020u;
      acc[q][0] += zero_bytes_to_ones(lw ^ 0x61616161u);
      acc[q][1] += zero_bytes_to_ones(lw ^ 0x63636363u);
      acc[q][2] += zero_bytes_to_ones(lw ^ 0x67676767u);
      acc[q][3] += zero_bytes_to_ones(lw ^ 0x74747474u);
    }
  }
  const int rows = (int)(r1 - r0);
#pragma unroll
  for (int q = 0; q < 4; q++)
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int64_t j = j0 + q * 4 + b;
      if (j >= L) continue;  // padding columns of the last group
      int sum = 0;
#pragma unroll
      for (int a = 0; a < 4; a++) {
        const int c = (int)((acc[q][a] >> (8 * b)) & 0xFFu);
        if (c) atomicAdd(&counts[j * 5 + a], c);
        sum += c;
      }
      if (rows - sum) atomicAdd(&counts[j * 5 + 4], rows - sum);
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

// Gather retained columns and classify (src/getACGTNsites.cpp:222-267): codes[k][s] = class(aln[s][pos[k]-1]).
// HBM-bound gather: retained columns are sparse in the row (one in ~22 at 616 x 2.2 Mb), so the useful unit is the
// 32-byte sector; a warp takes 32 consecutive SNPs of ONE row, so neighbouring SNPs share sectors and every sector of the
// row that holds a SNP is fetched once.  A block covers 32 rows x 128 SNPs (four loads in flight per thread before the
// first use) and transposes through shared memory so that the writes are contiguous along sequences.
constexpr int EX_K = 128;  // SNPs per block
// `aln` holds rows [row0, row0 + S) of an alignment of S_total records (row chunks of a streamed upload).
__global__ void __launch_bounds__(1024) extract_codes_kernel(const uint8_t* __restrict__ aln, int64_t S, int64_t L, int64_t pitch,
                                                             const int32_t* __restrict__ pos, int64_t n, uint8_t* codes,
                                                             int64_t row0, int64_t S_total) {
  __shared__ uint8_t tile[EX_K][33];
  const int64_t k0 = (int64_t)blockIdx.x * EX_K, s0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t s = s0 + ty;
  uint8_t v[EX_K / 32];
#pragma unroll
  for (int u = 0; u < EX_K / 32; u++) {
    const int64_t k = k0 + u * 32 + tx;
    v[u] = (k < n && s < S) ? __ldg(aln + s * pitch + (__ldg(pos + k) - 1)) : (uint8_t)0;
  }
#pragma unroll
  for (int u = 0; u < EX_K / 32; u++) tile[u * 32 + tx][ty] = (uint8_t)classify_char(v[u]);
  __syncthreads();
  const int64_t so = s0 + tx;
#pragma unroll
  for (int u = 0; u < EX_K / 32; u++) {
    const int64_t k = k0 + u * 32 + ty;
    if (k < n && so < S) codes[k * S_total + row0 + so] = tile[u * 32 + ty][tx];
  }
}

// src/ACGTN2num_parallel.cpp:18-41: uppercase A/C/G/T, and 'N' or '-' -> row 4; anything else untouched.
__global__ void acgtn2num_kernel(double* nv, const char* __restrict__ ref, int64_t n) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  char cc = ref[c];
  int row = cc == 'A' ? 0 : cc == 'C' ? 1 : cc == 'G' ? 2 : cc == 'T' ? 3 : (cc == 'N' || cc == '-') ? 4 : -1;
  if (row >= 0) nv[c * 5 + row] = 0.0;
}

int column_counts_device(cudaStream_t st, const uint8_t* d_aln, int64_t S, int64_t L, int64_t pitch, int32_t* d_counts,
                         bool accumulate) {
  if (pitch < L || (pitch & 15) || (reinterpret_cast<uintptr_t>(d_aln) & 15))
    return set_error(LDW_ERR_INTERNAL, "column_counts_device: rows must be 16-byte aligned (pitch %lld)", (long long)pitch);
  if (!accumulate) LDW_CUDA(cudaMemsetAsync(d_counts, 0, (size_t)L * 5 * 4, st));
  // slices of at most 255 rows (byte-lane counters); more of them when the column groups alone cannot fill the SMs
  const int64_t groups = (L + CC_COLS - 1) / CC_COLS;
  const int64_t col_blocks = (groups + 255) / 256;
  int64_t slices = (S + CC_MAX_ROWS - 1) / CC_MAX_ROWS;
  while (slices * col_blocks < 4 * 148 && slices * 2 <= S && slices < 65535) slices *= 2;
  if (slices > 65535) return set_error(LDW_ERR_UNSUPPORTED, "alignment with more than 16.7 M sequences");
  const int rps = (int)((S + slices - 1) / slices);
  dim3 grid((unsigned)col_blocks, (unsigned)slices);
  column_count_kernel<<<grid, 256, 0, st>>>(d_aln, S, L, pitch, rps, d_counts);
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

int extract_codes_device(cudaStream_t st, const uint8_t* d_aln, int64_t S, int64_t L, int64_t pitch, const int32_t* d_pos,
                         int64_t n, uint8_t* d_codes, int64_t row0, int64_t S_total) {
  if (n == 0 || S == 0) return 0;
  if ((S + 31) / 32 > 65535) return set_error(LDW_ERR_UNSUPPORTED, "alignment with more than 2 M sequences");
  dim3 grid((unsigned)((n + EX_K - 1) / EX_K), (unsigned)((S + 31) / 32));
  extract_codes_kernel<<<grid, dim3(32, 32), 0, st>>>(d_aln, S, L, pitch, d_pos, n, d_codes, row0, S_total);
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
