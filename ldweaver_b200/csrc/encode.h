// Internal (C++) interface of the encoding stage; see encode.cu.
#pragma once
#include "hdw.h"
#include "host_util.h"

namespace ldw {

int column_counts_device(cudaStream_t st, const uint8_t* d_aln, int64_t S, int64_t L, int32_t* d_counts /*[L x 5]*/);
int site_filter_device(cudaStream_t st, const int32_t* d_counts, int64_t L, int nseq, int filter, double gap_thresh,
                       double maf_thresh, int32_t* d_pos, int64_t* n_out);
int counts_to_double_device(cudaStream_t st, const int32_t* d_in, int64_t n, double* d_out);
int extract_codes_device(cudaStream_t st, const uint8_t* d_aln, int64_t S, int64_t L, const int32_t* d_pos, int64_t n,
                         uint8_t* d_codes);
int acgtn2num_device(cudaStream_t st, double* d_nv, const char* d_ref, int64_t n);

}  // namespace ldw
