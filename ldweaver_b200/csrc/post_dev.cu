// mergeNsort_sr_links (R/computePairwiseMI.R:400-495) on the DEVICE-RESIDENT short-range table of the last scan.
//
// The reference walks all ~10^8 short-range links as R data.frames (group_by(len) + quantile, fitdist, pbeta).  The host
// implementation (post_host.cpp) needs the 2.9 GB of link columns in host memory first; here the table never leaves HBM:
// every pass over the N links is an HBM-bound kernel (or a CUB primitive used for ordering / compaction only), and only
// what is O(groups) or O(links above the fit) crosses PCIe:
//   1. sr_group_count_kernel   membership (0 < len < sr_dist, clusters clust1 / clust2, :372-376, :417-419) + histogram
//                              of the (cluster, length) groups                                       [20 B/link read]
//   2. cub::DeviceScan         group offsets
//   3. sr_group_scatter_kernel MI of every membership into its group's slice                         [+8 B/entry written]
//   4. cub::DeviceSegmentedSort per-group order (library call, ordering step only)
//   5. sr_group_q95_kernel     type-7 95th percentile per group from the sorted slice (:422) -- order statistics, so the
//                              atomics-based scatter order does not matter and the result is bit-identical to the host's
//      host: log-log decay fit per cluster (fastLm, :428-429; shared code, post_host.cpp)
//   6. sr_residual_kernel      d = MI - fit[len], subscripted by the VALUE of len as the reference does (:448); flags of
//                              the links above the fit per cluster, and fixed-order block partials of n, sum log d,
//                              sum log(1-d), sum d, sum d^2
//      host: beta fit from those sufficient statistics (moment start + Nelder-Mead, :452; shared code)
//   7. cub::DeviceSelect::Flagged per cluster: the links above the fit in scan order (the reference's row order)
//   8. sr_srp_kernel           srp_max = -pbeta(d, a, b, lower = F, log = T) in fp64 (continued fraction, :453), and the smallest
//                              MI among the same-cluster links with srp_max > cutoff
//   9. sr_keep_kernel + select + sr_gather_kernel: ONLY the rows sr_links_red (srp_max > cutoff, :494) and
//                              sr_links_ARACNE_check (MI >= min(sr_links_red$MI), :495) need -- a few 10^5 of the 4.5 x 10^6
//                              links above the fit at 616 x 100k -- and the (few) cross-cluster links, which the host keeps
//                              once each (:474-483; shared code), come to the host
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <chrono>
#include <cmath>
#include <memory>

#include "../../include/ldw.h"
#include "ctx.h"
#include "post_host.h"

using namespace ldw;

namespace {

constexpr int MAXC = 64;  // clusters handled on the device path (cds_var$nclust is 2-3 in practice)

struct SrCols {
  const int32_t *c1, *c2, *len;
  const double* mi;
  int64_t n;
};

// clusters a link is listed under (0 = none): clust1 and, if different, clust2; only 0 < len < sr_dist
__device__ __forceinline__ bool sr_member(const SrCols& T, int64_t i, double sr_dist, int nclust, int& ca, int& cb, int& l) {
  l = T.len[i];
  if (!(l > 0 && (double)l < sr_dist)) return false;
  ca = T.c1[i]; cb = T.c2[i];
  if (ca < 1 || ca > nclust) ca = 0;
  if (cb < 1 || cb > nclust || cb == ca) cb = 0;
  return (ca | cb) != 0;
}

__global__ void sr_group_count_kernel(SrCols T, double sr_dist, int nclust, int64_t nl, uint32_t* gcount) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < T.n; i += (int64_t)gridDim.x * blockDim.x) {
    int ca, cb, l;
    if (!sr_member(T, i, sr_dist, nclust, ca, cb, l)) continue;
    if (ca) atomicAdd(&gcount[(int64_t)(ca - 1) * nl + l], 1u);
    if (cb) atomicAdd(&gcount[(int64_t)(cb - 1) * nl + l], 1u);
  }
}

__global__ void u32_to_i64_kernel(const uint32_t* in, int64_t n, int64_t* out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

__global__ void sr_group_scatter_kernel(SrCols T, double sr_dist, int nclust, int64_t nl, const int64_t* goff, unsigned long long* cursor,
                                        double* keys) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < T.n; i += (int64_t)gridDim.x * blockDim.x) {
    int ca, cb, l;
    if (!sr_member(T, i, sr_dist, nclust, ca, cb, l)) continue;
    const double v = T.mi[i];
    if (ca) { const int64_t g = (int64_t)(ca - 1) * nl + l; keys[goff[g] + (int64_t)atomicAdd(&cursor[g], 1ull)] = v; }
    if (cb) { const int64_t g = (int64_t)(cb - 1) * nl + l; keys[goff[g] + (int64_t)atomicAdd(&cursor[g], 1ull)] = v; }
  }
}

// stats::quantile.default(type = 7), prob 0.95, from an ascending slice (post_host.cpp:quantile7 on unsorted data)
__global__ void sr_group_q95_kernel(const double* sorted, const int64_t* goff, int64_t G, double* gq) {
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const int64_t n = goff[g + 1] - goff[g];
  if (n <= 0) { gq[g] = 0.0; return; }
  const double* x = sorted + goff[g];
  // every product and sum rounded on its own (no FMA contraction): the host evaluates the same expression that way
  const double index = __dadd_rn(1.0, __dmul_rn((double)(n - 1), 0.95));
  const int64_t lo = (int64_t)floor(index), hi = (int64_t)ceil(index);
  double qs = x[lo - 1];
  if (hi > lo) {
    const double xhi = x[lo];  // smallest of the values above position lo
    if (index > (double)lo && xhi != qs) {
      const double h = index - (double)lo;
      qs = __dadd_rn(__dmul_rn(1.0 - h, qs), __dmul_rn(h, xhi));
    }
  }
  gq[g] = qs;
}

struct FitTab {           // fitted decay per cluster as the reference subscripts it: fit[c][len - 1], len <= ng[c]
  const double* val;      // concatenated, cluster c at off[c - 1]
  int64_t off[MAXC + 1];
  int nclust;
};

__device__ __forceinline__ bool sr_residual(const FitTab& F, int c, int l, double mi, double& d) {
  if ((int64_t)l > F.off[c] - F.off[c - 1]) return false;  // NA in R: drops out of which(diff > 0)
  d = mi - F.val[F.off[c - 1] + l - 1];
  return d > 0;
}

constexpr int RES_THREADS = 256;
// Each block owns one contiguous piece of the table; its partial sums are reduced in a fixed order (strided per-thread
// accumulation, then a shared-memory tree), so the totals do not depend on scheduling.
__global__ void __launch_bounds__(RES_THREADS) sr_residual_kernel(SrCols T, double sr_dist, FitTab F, int64_t piece, uint8_t* flags /*[nclust][N]*/,
                                                                  double* part /*[blocks][nclust][5]*/, uint32_t* bad) {
  __shared__ double sh[RES_THREADS];
  const int64_t lo = (int64_t)blockIdx.x * piece, hi = min(T.n, lo + piece);
  for (int c = 1; c <= F.nclust; c++) {
    double acc[5] = {0, 0, 0, 0, 0};
    for (int64_t i = lo + threadIdx.x; i < hi; i += RES_THREADS) {
      int ca, cb, l;
      uint8_t f = 0;
      if (sr_member(T, i, sr_dist, F.nclust, ca, cb, l) && (ca == c || cb == c)) {
        double d;
        if (sr_residual(F, c, l, T.mi[i], d)) {
          f = 1;
          if (d > 1) atomicMax(bad, (uint32_t)c);
          acc[0] += 1.0; acc[1] += log(d); acc[2] += log1p(-d); acc[3] += d; acc[4] += d * d;
        }
      }
      flags[(int64_t)(c - 1) * T.n + i] = f;
    }
    for (int k = 0; k < 5; k++) {
      sh[threadIdx.x] = acc[k];
      __syncthreads();
      for (int o = RES_THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
      }
      if (threadIdx.x == 0) part[((int64_t)blockIdx.x * F.nclust + (c - 1)) * 5 + k] = sh[0];
      __syncthreads();
    }
  }
}

// ---- log of the regularised incomplete beta function (continued fraction, modified Lentz): post_host.cpp:log_ibeta_cf ----
__device__ double d_log_ibeta_cf(double a, double b, double x, double logx, double log1mx, double lbeta) {
  const double tiny = 1e-300, eps = 1e-16;
  const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (fabs(d) < tiny) d = tiny;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m <= 10000; m++) {
    const double m2 = 2.0 * m;
    double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    h *= d * c;
    aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < eps) break;
  }
  return a * logx + b * log1mx - lbeta - log(a) + log(h);
}

__device__ double d_neg_log_upper_beta(double x, double a, double b, double lbeta) {
  const double logx = log(x), log1mx = log1p(-x);
  const double y = 1.0 - x;
  if (y < (b + 1.0) / (a + b + 2.0)) return -d_log_ibeta_cf(b, a, y, log1mx, logx, lbeta);
  const double lp = d_log_ibeta_cf(a, b, x, logx, log1mx, lbeta);
  return -log1p(-exp(lp));
}

struct FullCols {
  const int32_t *pos1, *pos2, *c1, *c2, *len, *blk;
  const double* mi;
};

// order-preserving 64-bit key of a double (for atomicMin on MI values, which may be negative)
__device__ __forceinline__ unsigned long long dbl_key(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// srp_max of the links above the fit of cluster c (rows[k]); dup[k] = the link joins two clusters (it is listed under both
// and de-duplicated on the host); *min_red_key = smallest MI among the same-cluster links that pass the cut-off.
__global__ void sr_srp_kernel(FullCols T, FitTab F, int c, double a, double b, double lbeta, double cutoff, const int64_t* rows, int64_t n,
                              double* srp, uint8_t* dup, unsigned long long* min_red_key) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t i = rows[k];
  const double mi = T.mi[i];
  const double d = mi - F.val[F.off[c - 1] + T.len[i] - 1];
  const double v = d < 1.0 ? d_neg_log_upper_beta(d, a, b, lbeta) : INFINITY;
  srp[k] = v;
  const bool is_dup = T.c1[i] != T.c2[i];
  dup[k] = is_dup ? 1 : 0;
  if (!is_dup && v > cutoff) atomicMin(min_red_key, dbl_key(mi));  // NaN compares false: dropped as in :458
}

// which rows of the cluster's list travel to the host: every cross-cluster link (few; de-duplicated there), and the
// same-cluster links that are in sr_links_red (srp_max > cutoff) or in sr_links_ARACNE_check (MI >= min_mi)
__global__ void sr_keep_kernel(FullCols T, const int64_t* rows, const double* srp, const uint8_t* dup, int64_t n, double cutoff, double min_mi,
                               int phase, uint8_t* keep) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const double v = srp[k];
  bool kp;
  if (phase == 0) kp = dup[k] != 0;
  else kp = !dup[k] && !isnan(v) && (v > cutoff || T.mi[rows[k]] >= min_mi);
  keep[k] = kp ? 1 : 0;
}

__global__ void sr_gather_kernel(FullCols T, const int64_t* rows, const double* srp, const int64_t* pick, int64_t n, int64_t* o_row, double* o_srp,
                                 int32_t* o_pos1, int32_t* o_pos2, int32_t* o_c1, int32_t* o_c2, int32_t* o_len, int32_t* o_blk, double* o_mi) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int64_t k = pick[t], i = rows[k];
  o_row[t] = i; o_srp[t] = srp[k];
  o_pos1[t] = T.pos1[i]; o_pos2[t] = T.pos2[i]; o_c1[t] = T.c1[i]; o_c2[t] = T.c2[i]; o_len[t] = T.len[i]; o_blk[t] = T.blk[i]; o_mi[t] = T.mi[i];
}

__global__ void max_len_kernel(SrCols T, double sr_dist, int nclust, int* out) {
  int m = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < T.n; i += (int64_t)gridDim.x * blockDim.x) {
    int ca, cb, l;
    if (sr_member(T, i, sr_dist, nclust, ca, cb, l)) m = max(m, l);
  }
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

}  // namespace

extern "C" int ldw_sr_postprocess_dev(ldw_ctx* ctx, int32_t nclust, double sr_dist, double srp_cutoff, ldw_sr_post* out, ldw_links* df_rows_out) {
  return ldw::guarded("ldw_sr_postprocess_dev", [&]() -> int {
    if (!ctx || !out) return set_error(LDW_ERR_ARG, "ldw_sr_postprocess_dev: null argument");
    memset(out, 0, sizeof(*out));
    if (df_rows_out) memset(df_rows_out, 0, sizeof(*df_rows_out));
    LDW_TRY(ctx_bind(ctx));
    if (nclust < 1 || nclust > MAXC) return set_error(LDW_ERR_UNSUPPORTED, "ldw_sr_postprocess_dev: nclust must be in 1..%d", MAXC);
    const ldw_ctx::DevSr& D = ctx->dev_sr;
    if (D.n < 0) return set_error(LDW_ERR_ARG, "ldw_sr_postprocess_dev: no short-range table on the device (run ldw_mi_scan over the whole job, "
                                                "n_parts = 1, without LDW_SCAN_NO_LINKS / LDW_SCAN_LR_ONLY, first)");
    const int64_t N = D.n;
    if (N > 0x7fffffffLL) return set_error(LDW_ERR_UNSUPPORTED, "ldw_sr_postprocess_dev: more than 2^31 short-range links");
    cudaStream_t st = ctx->stream;
    SrCols T{D.c1, D.c2, D.len, D.mi, N};
    std::unique_ptr<ldwpost::SrPostPriv> S(new ldwpost::SrPostPriv());
    S->fit_off.assign(1, 0);
    const int nb_stream = ctx->num_sms * 8;
    const bool dbg_t = getenv("LDW_DBG_TIMING") != nullptr;  // phase times on stderr (never stdout)
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
      if (!dbg_t) return;
      cudaStreamSynchronize(st);
      const auto t = std::chrono::steady_clock::now();
      fprintf(stderr, "[ldw_sr_postprocess_dev] %-34s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
      t_last = t;
    };

    // ---- groups: (cluster, length) ----
    int32_t maxlen = 0;
    if (sr_dist <= (double)(1 << 18)) {
      maxlen = (int32_t)std::max(0.0, ceil(sr_dist) - 1.0);
    } else {
      DevBuf d_m;
      LDW_TRY(d_m.alloc(4));
      LDW_CUDA(cudaMemsetAsync(d_m.p, 0, 4, st));
      if (N > 0) max_len_kernel<<<nb_stream, 256, 0, st>>>(T, sr_dist, nclust, d_m.as<int>());
      LDW_CUDA(cudaMemcpyAsync(&maxlen, d_m.p, 4, cudaMemcpyDeviceToHost, st));
      LDW_CUDA(cudaStreamSynchronize(st));
    }
    const int64_t nl = (int64_t)maxlen + 1, G = (int64_t)nclust * nl;
    if (G > ((int64_t)1 << 27)) return set_error(LDW_ERR_UNSUPPORTED, "ldw_sr_postprocess_dev: %lld (cluster, length) groups", (long long)G);
    DevBuf d_gcount, d_gcnt64, d_goff, d_cursor, d_gq;
    LDW_TRY(d_gcount.alloc((size_t)G * 4));
    LDW_TRY(d_gcnt64.alloc((size_t)(G + 1) * 8));
    LDW_TRY(d_goff.alloc((size_t)(G + 1) * 8));
    LDW_TRY(d_cursor.alloc((size_t)G * 8));
    LDW_TRY(d_gq.alloc((size_t)G * 8));
    LDW_CUDA(cudaMemsetAsync(d_gcount.p, 0, (size_t)G * 4, st));
    LDW_CUDA(cudaMemsetAsync(d_cursor.p, 0, (size_t)G * 8, st));
    LDW_CUDA(cudaMemsetAsync(d_gcnt64.p, 0, (size_t)(G + 1) * 8, st));
    if (N > 0) sr_group_count_kernel<<<nb_stream, 256, 0, st>>>(T, sr_dist, nclust, nl, d_gcount.as<uint32_t>());
    u32_to_i64_kernel<<<(unsigned)((G + 255) / 256), 256, 0, st>>>(d_gcount.as<uint32_t>(), G, d_gcnt64.as<int64_t>());
    LDW_CUDA(cudaGetLastError());
    {
      size_t tb = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, tb, d_gcnt64.as<int64_t>(), d_goff.as<int64_t>(), (int)(G + 1), st);
      DevBuf tmp;
      LDW_TRY(tmp.alloc(tb));
      LDW_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, d_gcnt64.as<int64_t>(), d_goff.as<int64_t>(), (int)(G + 1), st));
      LDW_CUDA(cudaStreamSynchronize(st));
    }
    std::vector<int64_t> goff((size_t)G + 1);
    LDW_CUDA(cudaMemcpyAsync(goff.data(), d_goff.p, (size_t)(G + 1) * 8, cudaMemcpyDeviceToHost, st));
    LDW_CUDA(cudaStreamSynchronize(st));
    const int64_t E = goff[G];
    lap("group histogram + offsets");
    for (int32_t c = 1; c <= nclust; c++)
      if (goff[(int64_t)c * nl] == goff[(int64_t)(c - 1) * nl])
        return set_error(LDW_ERR_ARG, "ldw_sr_postprocess: cluster %d holds no short-range link with 0 < len < sr_dist", (int)c);
    // ---- per-group 95th percentiles ----
    std::vector<double> gq_all((size_t)G);
    {
      DevBuf d_keys, d_sorted, tmp;
      LDW_TRY(d_keys.alloc((size_t)std::max<int64_t>(E, 1) * 8));
      LDW_TRY(d_sorted.alloc((size_t)std::max<int64_t>(E, 1) * 8));
      sr_group_scatter_kernel<<<nb_stream, 256, 0, st>>>(T, sr_dist, nclust, nl, d_goff.as<int64_t>(), d_cursor.as<unsigned long long>(), d_keys.as<double>());
      LDW_CUDA(cudaGetLastError());
      if (E > 0x7fffffffLL) return set_error(LDW_ERR_UNSUPPORTED, "ldw_sr_postprocess_dev: more than 2^31 (link, cluster) entries");
      size_t tb = 0;
      cub::DeviceSegmentedSort::SortKeys(nullptr, tb, d_keys.as<double>(), d_sorted.as<double>(), (int)E, (int)G, d_goff.as<int64_t>(),
                                         d_goff.as<int64_t>() + 1, st);
      LDW_TRY(tmp.alloc(tb));
      LDW_CUDA(cub::DeviceSegmentedSort::SortKeys(tmp.p, tb, d_keys.as<double>(), d_sorted.as<double>(), (int)E, (int)G,
                                                  d_goff.as<int64_t>(), d_goff.as<int64_t>() + 1, st));
      lap("scatter + segmented sort");
      sr_group_q95_kernel<<<(unsigned)((G + 255) / 256), 256, 0, st>>>(d_sorted.as<double>(), d_goff.as<int64_t>(), G, d_gq.as<double>());
      LDW_CUDA(cudaGetLastError());
      LDW_CUDA(cudaMemcpyAsync(gq_all.data(), d_gq.p, (size_t)G * 8, cudaMemcpyDeviceToHost, st));
      LDW_CUDA(cudaStreamSynchronize(st));
    }
    std::vector<int64_t> glist;
    std::vector<double> gq;
    for (int64_t g = 0; g < G; g++)
      if (goff[g + 1] > goff[g]) { glist.push_back(g); gq.push_back(gq_all[g]); }
    lap("percentiles to host");
    // ---- decay fits (host) ----
    LDW_TRY(ldwpost::decay_fits(*S, nclust, nl, glist, gq));
    lap("decay fits (host)");
    // ---- residuals: flags + sufficient statistics ----
    FitTab F;
    memset(&F, 0, sizeof(F));
    F.nclust = nclust;
    DevBuf d_fit;
    LDW_TRY(d_fit.alloc(std::max<size_t>(S->fit_val.size(), 1) * 8));
    LDW_CUDA(cudaMemcpyAsync(d_fit.p, S->fit_val.data(), S->fit_val.size() * 8, cudaMemcpyHostToDevice, st));
    F.val = d_fit.as<double>();
    for (int c = 0; c <= nclust; c++) F.off[c] = S->fit_off[c];
    const int64_t piece = 1 << 16;  // fixed piece size: the partial sums, hence the fit, do not depend on the device
    const int64_t nblocks = std::max<int64_t>(1, (N + piece - 1) / piece);
    DevBuf d_flags, d_part, d_bad;
    LDW_TRY(d_flags.alloc((size_t)nclust * (size_t)std::max<int64_t>(N, 1)));
    LDW_TRY(d_part.alloc((size_t)nblocks * nclust * 5 * 8));
    LDW_TRY(d_bad.alloc(4));
    LDW_CUDA(cudaMemsetAsync(d_bad.p, 0, 4, st));
    sr_residual_kernel<<<(unsigned)nblocks, RES_THREADS, 0, st>>>(T, sr_dist, F, piece, d_flags.as<uint8_t>(), d_part.as<double>(), d_bad.as<uint32_t>());
    LDW_CUDA(cudaGetLastError());
    std::vector<double> part((size_t)nblocks * nclust * 5);
    uint32_t bad = 0;
    LDW_CUDA(cudaMemcpyAsync(part.data(), d_part.p, part.size() * 8, cudaMemcpyDeviceToHost, st));
    LDW_CUDA(cudaMemcpyAsync(&bad, d_bad.p, 4, cudaMemcpyDeviceToHost, st));
    LDW_CUDA(cudaStreamSynchronize(st));
    if (bad) return set_error(LDW_ERR_ARG, "ldw_sr_postprocess: cluster %d: values must be in [0-1] to fit a beta distribution", (int)bad);
    lap("residual flags + sums");

    FullCols FT{D.pos1, D.pos2, D.c1, D.c2, D.len, D.blk, D.mi};
    // Per cluster: the links above the fit in scan order, their srp_max, the smallest MI of those that pass the cut-off.
    struct PerC { DevBuf rows, srp, dup; int64_t npos = 0; };
    std::vector<PerC> pc((size_t)nclust + 1);
    DevBuf d_tmp, d_nsel, d_minkey, d_keep, d_pick;
    LDW_TRY(d_nsel.alloc(8));
    LDW_TRY(d_minkey.alloc(8));
    LDW_CUDA(cudaMemsetAsync(d_minkey.p, 0xFF, 8, st));
    for (int32_t c = 1; c <= nclust; c++) {
      long double cnt = 0, s1 = 0, s2 = 0, sm = 0, sq = 0;
      for (int64_t k = 0; k < nblocks; k++) {
        const double* p = &part[((size_t)k * nclust + (c - 1)) * 5];
        cnt += p[0]; s1 += p[1]; s2 += p[2]; sm += p[3]; sq += p[4];
      }
      const int64_t npos = (int64_t)llroundl(cnt);
      if (npos < 2) return set_error(LDW_ERR_ARG, "ldw_sr_postprocess: cluster %d has fewer than two links above the fitted decay", (int)c);
      const double mean = (double)(sm / (long double)npos);
      const double v = (double)((sq - (long double)npos * (long double)mean * (long double)mean) / (long double)npos);  // biased variance
      double par[2];
      LDW_TRY(ldwpost::beta_fit(*S, c, npos, (double)s1, (double)s2, mean, v, par));
      const double lbeta = ldwpost::lbeta_fn(par[0], par[1]);
      PerC& P = pc[c];
      P.npos = npos;
      LDW_TRY(P.rows.alloc((size_t)npos * 8));
      LDW_TRY(P.srp.alloc((size_t)npos * 8));
      LDW_TRY(P.dup.alloc((size_t)npos));
      {
        thrust::counting_iterator<int64_t> it(0);
        size_t tb = 0;
        cub::DeviceSelect::Flagged(nullptr, tb, it, d_flags.as<uint8_t>() + (size_t)(c - 1) * N, P.rows.as<int64_t>(), d_nsel.as<int64_t>(), (int)N, st);
        LDW_TRY(d_tmp.ensure(tb));
        LDW_CUDA(cub::DeviceSelect::Flagged(d_tmp.p, tb, it, d_flags.as<uint8_t>() + (size_t)(c - 1) * N, P.rows.as<int64_t>(),
                                            d_nsel.as<int64_t>(), (int)N, st));
      }
      int64_t nsel = 0;
      LDW_CUDA(cudaMemcpyAsync(&nsel, d_nsel.p, 8, cudaMemcpyDeviceToHost, st));
      LDW_CUDA(cudaStreamSynchronize(st));
      if (nsel != npos) return set_error(LDW_ERR_INTERNAL, "ldw_sr_postprocess_dev: %lld flagged links but %lld counted", (long long)nsel, (long long)npos);
      sr_srp_kernel<<<(unsigned)((npos + 127) / 128), 128, 0, st>>>(FT, F, c, par[0], par[1], lbeta, srp_cutoff, P.rows.as<int64_t>(), npos,
                                                                   P.srp.as<double>(), P.dup.as<uint8_t>(), d_minkey.as<unsigned long long>());
      LDW_CUDA(cudaGetLastError());
    }
    lap("select + srp_max (all clusters)");
    // Only what sr_links_red / sr_links_ARACNE_check need comes to the host (sr_links_df itself -- every link above the fit,
    // 4.5e6 rows at 616 x 100k -- stays on the device): first the cross-cluster links, whose de-duplication decides which of
    // them are in sr_links_red and therefore min(sr_links_red$MI); then the same-cluster rows of the two sets.
    struct Got { std::vector<int64_t> row; std::vector<double> srp, mi; std::vector<int32_t> pos1, pos2, c1, c2, len, blk; };
    auto fetch = [&](int32_t c, int phase, double min_mi, Got& G) -> int {
      PerC& P = pc[c];
      const int64_t npos = P.npos;
      LDW_TRY(d_keep.ensure((size_t)npos));
      LDW_TRY(d_pick.ensure((size_t)npos * 8));
      sr_keep_kernel<<<(unsigned)((npos + 255) / 256), 256, 0, st>>>(FT, P.rows.as<int64_t>(), P.srp.as<double>(), P.dup.as<uint8_t>(), npos, srp_cutoff,
                                                                    min_mi, phase, d_keep.as<uint8_t>());
      thrust::counting_iterator<int64_t> it(0);
      size_t tb = 0;
      cub::DeviceSelect::Flagged(nullptr, tb, it, d_keep.as<uint8_t>(), d_pick.as<int64_t>(), d_nsel.as<int64_t>(), (int)npos, st);
      LDW_TRY(d_tmp.ensure(tb));
      LDW_CUDA(cub::DeviceSelect::Flagged(d_tmp.p, tb, it, d_keep.as<uint8_t>(), d_pick.as<int64_t>(), d_nsel.as<int64_t>(), (int)npos, st));
      int64_t m = 0;
      LDW_CUDA(cudaMemcpyAsync(&m, d_nsel.p, 8, cudaMemcpyDeviceToHost, st));
      LDW_CUDA(cudaStreamSynchronize(st));
      G.row.resize((size_t)m); G.srp.resize((size_t)m); G.mi.resize((size_t)m);
      for (auto* v32 : {&G.pos1, &G.pos2, &G.c1, &G.c2, &G.len, &G.blk}) v32->resize((size_t)m);
      if (m == 0) return 0;
      DevBuf o_row, o_srp, o1, o2, o3, o4, o5, o6, o7;
      LDW_TRY(o_row.alloc((size_t)m * 8)); LDW_TRY(o_srp.alloc((size_t)m * 8)); LDW_TRY(o7.alloc((size_t)m * 8));
      LDW_TRY(o1.alloc((size_t)m * 4)); LDW_TRY(o2.alloc((size_t)m * 4)); LDW_TRY(o3.alloc((size_t)m * 4));
      LDW_TRY(o4.alloc((size_t)m * 4)); LDW_TRY(o5.alloc((size_t)m * 4)); LDW_TRY(o6.alloc((size_t)m * 4));
      sr_gather_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(FT, P.rows.as<int64_t>(), P.srp.as<double>(), d_pick.as<int64_t>(), m, o_row.as<int64_t>(),
                                                                   o_srp.as<double>(), o1.as<int32_t>(), o2.as<int32_t>(), o3.as<int32_t>(), o4.as<int32_t>(),
                                                                   o5.as<int32_t>(), o6.as<int32_t>(), o7.as<double>());
      LDW_CUDA(cudaGetLastError());
      LDW_CUDA(cudaMemcpyAsync(G.row.data(), o_row.p, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
      LDW_CUDA(cudaMemcpyAsync(G.srp.data(), o_srp.p, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
      LDW_CUDA(cudaMemcpyAsync(G.mi.data(), o7.p, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
      LDW_CUDA(cudaMemcpyAsync(G.pos1.data(), o1.p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
      LDW_CUDA(cudaMemcpyAsync(G.pos2.data(), o2.p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
      LDW_CUDA(cudaMemcpyAsync(G.c1.data(), o3.p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
      LDW_CUDA(cudaMemcpyAsync(G.c2.data(), o4.p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
      LDW_CUDA(cudaMemcpyAsync(G.len.data(), o5.p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
      LDW_CUDA(cudaMemcpyAsync(G.blk.data(), o6.p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
      LDW_CUDA(cudaStreamSynchronize(st));
      return 0;
    };
    auto append = [&](const Got& G, int64_t t) -> int64_t {  // one row into the gathered table; returns its position
      S->g_pos1.push_back(G.pos1[t]); S->g_pos2.push_back(G.pos2[t]); S->g_c1.push_back(G.c1[t]); S->g_c2.push_back(G.c2[t]);
      S->g_len.push_back(G.len[t]); S->g_blk.push_back(G.blk[t]); S->g_mi.push_back(G.mi[t]);
      return (int64_t)S->g_mi.size() - 1;
    };
    // ---- phase 0: cross-cluster links -> de-duplicated (:474-483) ----
    std::vector<int64_t> dup_row, dup_at, dup_keep_row, dup_keep_at;
    std::vector<int32_t> dup_c, dup_keep_c;
    std::vector<double> dup_srp, dup_keep_srp;
    for (int32_t c = 1; c <= nclust; c++) {
      Got G;
      LDW_TRY(fetch(c, 0, 0.0, G));
      for (size_t t = 0; t < G.row.size(); t++) {
        if (std::isnan(G.srp[t])) continue;  // :458
        dup_row.push_back(G.row[t]); dup_at.push_back(append(G, (int64_t)t)); dup_c.push_back(c); dup_srp.push_back(G.srp[t]);
      }
    }
    {
      ldwpost::SrPostPriv tmp;  // run the shared de-duplication on the cross-cluster links alone
      ldwpost::LinkCols cols{S->g_pos1.data(), S->g_pos2.data(), S->g_c1.data(), S->g_c2.data(), S->g_len.data(), S->g_mi.data()};
      std::vector<int64_t> at;
      ldwpost::dedup_and_select(tmp, cols, dup_row, dup_at, dup_c, dup_srp, at, INFINITY);
      dup_keep_row = tmp.row; dup_keep_c = tmp.clust_c; dup_keep_srp = tmp.srp; dup_keep_at = at;
    }
    unsigned long long minkey = ~0ull;
    LDW_CUDA(cudaMemcpyAsync(&minkey, d_minkey.p, 8, cudaMemcpyDeviceToHost, st));
    LDW_CUDA(cudaStreamSynchronize(st));
    double min_mi = INFINITY;  // min(sr_links_red$MI), :495
    if (minkey != ~0ull) {
      const unsigned long long bts = (minkey >> 63) ? (minkey & 0x7FFFFFFFFFFFFFFFull) : ~minkey;
      memcpy(&min_mi, &bts, 8);
    }
    for (size_t k = 0; k < dup_keep_row.size(); k++)
      if (dup_keep_srp[k] > srp_cutoff) min_mi = std::min(min_mi, S->g_mi[(size_t)dup_keep_at[k]]);
    lap("cross-cluster links + min MI");
    // ---- phase 1: same-cluster rows of sr_links_red / sr_links_ARACNE_check, cluster by cluster in scan order ----
    std::vector<int64_t> df_at;
    for (int32_t c = 1; c <= nclust; c++) {
      Got G;
      LDW_TRY(fetch(c, 1, min_mi, G));
      for (size_t t = 0; t < G.row.size(); t++) {
        S->row.push_back(G.row[t]); S->clust_c.push_back(c); S->srp.push_back(G.srp[t]);
        df_at.push_back(append(G, (int64_t)t));
      }
    }
    for (size_t k = 0; k < dup_keep_row.size(); k++) {  // the de-duplicated cross-cluster links follow (:483), those in either set
      const double mi = S->g_mi[(size_t)dup_keep_at[k]];
      if (!(dup_keep_srp[k] > srp_cutoff || mi >= min_mi)) continue;
      S->row.push_back(dup_keep_row[k]); S->clust_c.push_back(dup_keep_c[k]); S->srp.push_back(dup_keep_srp[k]);
      df_at.push_back(dup_keep_at[k]);
    }
    lap("rows of red / check sets");
    const int64_t ndf = (int64_t)S->row.size();
    for (int64_t i = 0; i < ndf; i++) {
      if (S->srp[(size_t)i] > srp_cutoff) S->red.push_back(i);
      if (S->g_mi[(size_t)df_at[(size_t)i]] >= min_mi) S->chk.push_back(i);
    }
    {  // the gathered columns in df order (the gathered table also holds the losing copies of cross-cluster links)
      std::vector<int32_t> a1(ndf), a2(ndf), a3(ndf), a4(ndf), a5(ndf), a6(ndf);
      std::vector<double> a7(ndf);
      for (int64_t k = 0; k < ndf; k++) {
        const int64_t at = df_at[(size_t)k];
        a1[k] = S->g_pos1[at]; a2[k] = S->g_pos2[at]; a3[k] = S->g_c1[at]; a4[k] = S->g_c2[at]; a5[k] = S->g_len[at]; a6[k] = S->g_blk[at];
        a7[k] = S->g_mi[at];
      }
      S->g_pos1.swap(a1); S->g_pos2.swap(a2); S->g_c1.swap(a3); S->g_c2.swap(a4); S->g_len.swap(a5); S->g_blk.swap(a6); S->g_mi.swap(a7);
    }
    lap("red / chk + reorder (host)");
    ldwpost::publish(*S, nclust, out);
    if (df_rows_out) {
      df_rows_out->n = out->n_df;
      df_rows_out->pos1 = S->g_pos1.data(); df_rows_out->pos2 = S->g_pos2.data(); df_rows_out->clust1 = S->g_c1.data();
      df_rows_out->clust2 = S->g_c2.data(); df_rows_out->len = S->g_len.data(); df_rows_out->MI = S->g_mi.data();
      df_rows_out->block = S->g_blk.data();
    }
    out->priv = S.release();
    return 0;
  });
}
