// Internal (C++) interface of the encoding stage; see encode.cu.
#pragma once
#include "hdw.h"
#include "host_util.h"

namespace ldw {

// d_aln: S rows of `pitch` bytes (pitch >= L, a multiple of 16; base 16-byte aligned), the first L of each being the record
// accumulate: add to d_counts instead of clearing it first (row chunks of a streamed alignment)
int column_counts_device(cudaStream_t st, const uint8_t* d_aln, int64_t S, int64_t L, int64_t pitch, int32_t* d_counts /*[L x 5]*/,
                         bool accumulate = false);
int site_filter_device(cudaStream_t st, const int32_t* d_counts, int64_t L, int nseq, int filter, double gap_thresh,
                       double maf_thresh, int32_t* d_pos, int64_t* n_out);
int counts_to_double_device(cudaStream_t st, const int32_t* d_in, int64_t n, double* d_out);
// d_aln holds rows [row0, row0 + S) of an alignment of S_total records; codes is [n x S_total]
int extract_codes_device(cudaStream_t st, const uint8_t* d_aln, int64_t S, int64_t L, int64_t pitch, const int32_t* d_pos,
                         int64_t n, uint8_t* d_codes, int64_t row0, int64_t S_total);
int acgtn2num_device(cudaStream_t st, double* d_nv, const char* d_ref, int64_t n);

}  // namespace ldw
