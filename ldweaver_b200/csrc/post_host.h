// Pieces of mergeNsort_sr_links (R/computePairwiseMI.R:400-495) shared by the host implementation (post_host.cpp:
// ldw_sr_postprocess) and the device one (post_dev.cu: ldw_sr_postprocess_dev): everything that is O(groups) or
// O(links above the fit) stays on the host in both.
#pragma once
#include <stdint.h>

#include <vector>

#include "../../include/ldw.h"

namespace ldwpost {

struct SrPostPriv {
  std::vector<int32_t> clust_c;
  std::vector<int64_t> row;
  std::vector<double> srp;
  std::vector<int64_t> red, chk;
  std::vector<int64_t> fit_off;
  std::vector<int32_t> fit_len;
  std::vector<double> fit_q95, fit_val, coef, shape, start;
  std::vector<int64_t> n_pos;
  std::vector<int32_t> nm_evals, nm_fail;
  // device path: link columns of the df rows (the full table never reaches the host)
  std::vector<int32_t> g_pos1, g_pos2, g_c1, g_c2, g_len, g_blk;
  std::vector<double> g_mi;
};

struct LinkCols {
  const int32_t *pos1, *pos2, *clust1, *clust2, *len;
  const double* MI;
};

// glist: non-empty groups (cluster-major, ascending length; group id = (c - 1) * nl + len), gq: their 95th percentiles
int decay_fits(SrPostPriv& S, int32_t nclust, int64_t nl, const std::vector<int64_t>& glist, const std::vector<double>& gq);
int beta_fit(SrPostPriv& S, int32_t c, int64_t npos, double s1, double s2, double mean, double v, double par[2]);
double lbeta_fn(double a, double b);
void dedup_and_select(SrPostPriv& S, const LinkCols& cols, const std::vector<int64_t>& dup_row, const std::vector<int64_t>& dup_at,
                      const std::vector<int32_t>& dup_c, const std::vector<double>& dup_srp, std::vector<int64_t>& df_at, double srp_cutoff);
void publish(SrPostPriv& S, int32_t nclust, ldw_sr_post* out);

}  // namespace ldwpost
