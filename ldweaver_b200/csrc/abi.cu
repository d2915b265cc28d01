// C-ABI entry points for context management, encoding and Hamming-distance weights.
// (The MI-scan entry points live in mi_scan.cu.)
#include "../../include/ldw.h"

#include "ctx.h"
#include "encode.h"
#include "hdw.h"

using namespace ldw;

namespace ldw {
int ctx_bind(ldw_ctx* ctx) {
  if (!ctx) return set_error(LDW_ERR_ARG, "null context");
  LDW_CUDA(cudaSetDevice(ctx->device));
  return 0;
}
}  // namespace ldw

extern "C" {

int ldw_abi_version(void) { return LDW_ABI_VERSION; }

const char* ldw_last_error(void) { return last_error_ref().c_str(); }

int ldw_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int ldw_create(int device, ldw_ctx** out) {
  if (!out) return set_error(LDW_ERR_ARG, "ldw_create: null out");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return set_error(LDW_ERR_CUDA, "ldw_create: no CUDA device available (%s); this library has no CPU fallback",
                     e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  }
  if (device < 0 || device >= n) return set_error(LDW_ERR_ARG, "ldw_create: device %d out of range [0,%d)", device, n);
  LDW_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  LDW_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return set_error(LDW_ERR_UNSUPPORTED, "ldw_create: device %d is sm_%d%d; this build targets sm_100a (B200) only", device,
                     prop.major, prop.minor);
  ldw_ctx* c = new ldw_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete c;
    return set_error(LDW_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
  }
  e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->select_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->upload_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    cudaStreamDestroy(c->stream);
    delete c;
    return set_error(LDW_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
  }
  *out = c;
  return 0;
}

void ldw_destroy(ldw_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->scan_ws && ctx->scan_ws_free) ctx->scan_ws_free(ctx->scan_ws);
  cudaDeviceSynchronize();
  ldw::dev_cache_trim();
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->select_stream) cudaStreamDestroy(ctx->select_stream);
  if (ctx->upload_stream) cudaStreamDestroy(ctx->upload_stream);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int ldw_aln_param(ldw_ctx* ctx, const uint8_t* aln, int64_t nseq, int64_t seq_len, int filter, double gap_thresh,
                  double maf_thresh, int32_t* pos_out, int64_t* n_snp_out, double* counts_out) {
  LDW_TRY(ctx_bind(ctx));
  if (!aln || !pos_out || !n_snp_out) return set_error(LDW_ERR_ARG, "ldw_aln_param: null argument");
  if (nseq <= 0 || seq_len <= 0) return set_error(LDW_ERR_ARG, "ldw_aln_param: empty alignment (%lld x %lld)", (long long)nseq, (long long)seq_len);
  if (nseq > 0x7fffffffLL || seq_len > 0x7fffffffLL) return set_error(LDW_ERR_UNSUPPORTED, "ldw_aln_param: dimensions exceed int32");
  if (filter != 0 && filter != 1) return set_error(LDW_ERR_ARG, "ldw_aln_param: filter must be 0 (default) or 1 (relaxed)");
  cudaStream_t st = ctx->stream;
  DevBuf d_aln, d_counts, d_pos, d_dbl;
  LDW_TRY(d_aln.alloc((size_t)nseq * seq_len));
  LDW_TRY(d_counts.alloc((size_t)seq_len * 5 * 4));
  LDW_TRY(d_pos.alloc((size_t)seq_len * 4));
  LDW_CUDA(cudaMemcpyAsync(d_aln.p, aln, (size_t)nseq * seq_len, cudaMemcpyHostToDevice, st));
  LDW_TRY(column_counts_device(st, d_aln.as<uint8_t>(), nseq, seq_len, d_counts.as<int32_t>()));
  int64_t n = 0;
  LDW_TRY(site_filter_device(st, d_counts.as<int32_t>(), seq_len, (int)nseq, filter, gap_thresh, maf_thresh, d_pos.as<int32_t>(), &n));
  *n_snp_out = n;
  if (n > 0) LDW_CUDA(cudaMemcpyAsync(pos_out, d_pos.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  if (counts_out) {
    LDW_TRY(d_dbl.alloc((size_t)seq_len * 5 * 8));
    LDW_TRY(counts_to_double_device(st, d_counts.as<int32_t>(), seq_len * 5, d_dbl.as<double>()));
    LDW_CUDA(cudaMemcpyAsync(counts_out, d_dbl.p, (size_t)seq_len * 5 * 8, cudaMemcpyDeviceToHost, st));
  }
  LDW_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int ldw_extract_snps(ldw_ctx* ctx, const uint8_t* aln, int64_t nseq, int64_t seq_len, const int32_t* pos, int64_t n_snp,
                     uint8_t* codes_out, double* table_out) {
  LDW_TRY(ctx_bind(ctx));
  if (!aln || !pos || !codes_out) return set_error(LDW_ERR_ARG, "ldw_extract_snps: null argument");
  if (nseq <= 0 || seq_len <= 0 || n_snp <= 0) return set_error(LDW_ERR_ARG, "ldw_extract_snps: empty input");
  for (int64_t k = 0; k < n_snp; k++)
    if (pos[k] < 1 || pos[k] > seq_len) return set_error(LDW_ERR_ARG, "ldw_extract_snps: POS[%lld]=%d outside 1..%lld", (long long)k, pos[k], (long long)seq_len);
  cudaStream_t st = ctx->stream;
  DevBuf d_aln, d_pos, d_codes, d_table, d_mask, d_r, d_dbl;
  LDW_TRY(d_aln.alloc((size_t)nseq * seq_len));
  LDW_TRY(d_pos.alloc((size_t)n_snp * 4));
  LDW_TRY(d_codes.alloc((size_t)n_snp * nseq));
  LDW_CUDA(cudaMemcpyAsync(d_aln.p, aln, (size_t)nseq * seq_len, cudaMemcpyHostToDevice, st));
  LDW_CUDA(cudaMemcpyAsync(d_pos.p, pos, (size_t)n_snp * 4, cudaMemcpyHostToDevice, st));
  LDW_TRY(extract_codes_device(st, d_aln.as<uint8_t>(), nseq, seq_len, d_pos.as<int32_t>(), n_snp, d_codes.as<uint8_t>()));
  LDW_CUDA(cudaMemcpyAsync(codes_out, d_codes.p, (size_t)n_snp * nseq, cudaMemcpyDeviceToHost, st));
  if (table_out) {
    LDW_TRY(d_table.alloc((size_t)n_snp * 5 * 4));
    LDW_TRY(d_mask.alloc((size_t)n_snp));
    LDW_TRY(d_r.alloc((size_t)n_snp));
    LDW_TRY(d_dbl.alloc((size_t)n_snp * 5 * 8));
    LDW_TRY(snp_allele_stats(st, d_codes.as<uint8_t>(), n_snp, nseq, d_table.as<int32_t>(), d_mask.as<uint8_t>(), d_r.as<uint8_t>()));
    LDW_TRY(counts_to_double_device(st, d_table.as<int32_t>(), n_snp * 5, d_dbl.as<double>()));
    LDW_CUDA(cudaMemcpyAsync(table_out, d_dbl.p, (size_t)n_snp * 5 * 8, cudaMemcpyDeviceToHost, st));
  }
  LDW_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int ldw_acgtn2num(ldw_ctx* ctx, double* nv, const char* ref, int64_t n) {
  LDW_TRY(ctx_bind(ctx));
  if (n == 0) return 0;
  if (!nv || !ref || n < 0) return set_error(LDW_ERR_ARG, "ldw_acgtn2num: bad argument");
  cudaStream_t st = ctx->stream;
  DevBuf d_nv, d_ref;
  LDW_TRY(d_nv.alloc((size_t)n * 5 * 8));
  LDW_TRY(d_ref.alloc((size_t)n));
  LDW_CUDA(cudaMemcpyAsync(d_nv.p, nv, (size_t)n * 5 * 8, cudaMemcpyHostToDevice, st));
  LDW_CUDA(cudaMemcpyAsync(d_ref.p, ref, (size_t)n, cudaMemcpyHostToDevice, st));
  LDW_TRY(acgtn2num_device(st, d_nv.as<double>(), d_ref.as<char>(), n));
  LDW_CUDA(cudaMemcpyAsync(nv, d_nv.p, (size_t)n * 5 * 8, cudaMemcpyDeviceToHost, st));
  LDW_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int ldw_hdw(ldw_ctx* ctx, const uint8_t* codes, int64_t n_snp, int64_t nseq, double threshold, int32_t* cnt_out,
            double* hdw_out, int32_t* dist_out) {
  LDW_TRY(ctx_bind(ctx));
  if (!codes || !hdw_out) return set_error(LDW_ERR_ARG, "ldw_hdw: null argument");
  if (n_snp <= 0 || nseq <= 0) return set_error(LDW_ERR_ARG, "ldw_hdw: empty input");
  if (nseq > 0x7fffffffLL / 4) return set_error(LDW_ERR_UNSUPPORTED, "ldw_hdw: too many sequences");
  cudaStream_t st = ctx->stream;
  int thresh = (int)((double)n_snp * threshold);  // as.integer(nsnp*threshold): truncation (:23)
  DevBuf d_codes, d_neigh, d_w, d_dist;
  LDW_TRY(d_codes.alloc((size_t)n_snp * nseq));
  LDW_TRY(d_neigh.alloc((size_t)nseq * 4));
  LDW_TRY(d_w.alloc((size_t)nseq * 8));
  if (dist_out) LDW_TRY(d_dist.alloc((size_t)nseq * nseq * 4));
  LDW_CUDA(cudaMemcpyAsync(d_codes.p, codes, (size_t)n_snp * nseq, cudaMemcpyHostToDevice, st));
  LDW_TRY(hdw_device(st, d_codes.as<uint8_t>(), n_snp, nseq, thresh, d_neigh.as<int32_t>(), d_w.as<double>(),
                     dist_out ? d_dist.as<int32_t>() : nullptr, ctx->num_sms));
  if (cnt_out) LDW_CUDA(cudaMemcpyAsync(cnt_out, d_neigh.p, (size_t)nseq * 4, cudaMemcpyDeviceToHost, st));
  LDW_CUDA(cudaMemcpyAsync(hdw_out, d_w.p, (size_t)nseq * 8, cudaMemcpyDeviceToHost, st));
  if (dist_out) LDW_CUDA(cudaMemcpyAsync(dist_out, d_dist.p, (size_t)nseq * nseq * 4, cudaMemcpyDeviceToHost, st));
  LDW_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"
