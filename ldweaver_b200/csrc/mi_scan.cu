// Host driver + C ABI of the weighted pairwise-MI scan (see mi_kernel.cuh for the hot kernel).
//
// Reference semantics reproduced here (paths relative to the LDWeaver tree):
//   make_blocks                     R/computePairwiseMI.R:147-165   (block list, row-major i, j >= i)
//   perform_MI_computation_ACGTN    R/computePairwiseMI.R:167-386   (per block: MI, pair order Q2/Q5, len, sr/lr split,
//                                                                     per-block type-7 quantile Q3, SR-only drop Q12)
//   computeMI_Sprase/.fastHadamard  R/computePairwiseMI.R:390-398, src/computeMI.cpp:11-21 (incl. quirk Q1)
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cub/cub.cuh>
#include <vector>

#include "../../include/ldw.h"
#include "ctx.h"
#include "hdw.h"
#include "mi_aux.cuh"
#include "mi_kernel.cuh"
#include "mi_scan.h"

using namespace ldw;

namespace {

struct Group {
  int32_t slot0 = 0, nslots = 0, row0 = 0, count = 0;
};

struct DevLinks {
  DevBuf pos1, pos2, c1, c2, len, blk, mi;
  int ensure(int64_t m) {
    size_t b4 = (size_t)(m > 0 ? m : 1) * 4, b8 = (size_t)(m > 0 ? m : 1) * 8;
    LDW_TRY(pos1.ensure(b4)); LDW_TRY(pos2.ensure(b4)); LDW_TRY(c1.ensure(b4)); LDW_TRY(c2.ensure(b4));
    LDW_TRY(len.ensure(b4)); LDW_TRY(blk.ensure(b4)); LDW_TRY(mi.ensure(b8));
    return 0;
  }
};

// Everything one block needs on the device besides the static plan.
struct BlockDev {
  DevBuf rowdyn, coldyn, colinfo, tiles, from_idx, to_idx, rfl, rtl;
  PinnedBuf stage;
  cudaEvent_t uploaded = nullptr;               // the block's tables have arrived (upload stream)
  cudaEvent_t done = nullptr, done2 = nullptr;  // last use on the scan stream / on the select stream
  bool used = false, used2 = false;
};

struct BlockHost {  // pageable staging reused per ring entry (kept alive until the async copies are issued)
  std::vector<RowDyn> rowdyn;
  std::vector<ColDyn> coldyn;
  std::vector<ColInfo> colinfo;
  std::vector<TileDesc> tiles;
  std::vector<int32_t> from_idx, to_idx;
  std::vector<uint8_t> rfl, rtl;
  int32_t nf = 0, nt = 0;
  int32_t n_real_tiles = 0;  // tiles [n_real_tiles, tiles.size()) are the pilot sample (copies of real tiles)
  int diag = 0, ragged = 0;
  int64_t n_pairs = 0, n_sr = 0, n_lr = 0;
};

}  // namespace

// Scratch shared by all plans of one context (see ctx.h).
struct ScanWS {
  static constexpr int RING = 4;
  BlockDev ring[RING];
  BlockHost hring[RING];
  DevBuf d_dbg;
  // long-range collection state, triple-buffered: block b's candidates are pre-selected, refined and ranked on the
  // select stream while blocks b+1 and b+2 are being scanned into the other buffers
  struct LrBuf {
    DevBuf cand, vcand, mi64, state /*count, tcand, overflow, -, -, vcount*/, hist;
    cudaEvent_t scan_done = nullptr, sel_done = nullptr;
    bool used = false, chain_valid = false;
  } lr[3];
  static constexpr int NLR = 3;
  DevBuf d_state /*-, -, -, kept_overflow, -, -, -, -, chain[3]*/, d_sr_f32, d_dense;
  // written by the selection kernels / the publish kernel straight into host memory (pinned memory is device-
  // accessible under unified addressing): no small device->host copy has to queue behind the link columns
  PinnedBuf h_results, h_pub;
  DevBuf d_kept_key, d_kept_gi, d_kept_gj, d_kept_mi, d_kept_count, d_sort_tmp, d_keys_sorted, d_order_in, d_order_out;
  DevLinks d_sr, d_lr;
  bool events_ready = false;
  ~ScanWS() {
    for (auto& b : ring) {
      if (b.done) cudaEventDestroy(b.done);
      if (b.done2) cudaEventDestroy(b.done2);
      if (b.uploaded) cudaEventDestroy(b.uploaded);
    }
    for (auto& l : lr) {
      if (l.scan_done) cudaEventDestroy(l.scan_done);
      if (l.sel_done) cudaEventDestroy(l.sel_done);
    }
  }
};

static void scan_ws_free(void* p) { delete static_cast<ScanWS*>(p); }

static ScanWS* get_ws(ldw_ctx* ctx) {
  if (!ctx->scan_ws) {
    ctx->scan_ws = new ScanWS();
    ctx->scan_ws_free = scan_ws_free;
  }
  return static_cast<ScanWS*>(ctx->scan_ws);
}

struct ldw_mi_plan {
  ldw_ctx* ctx = nullptr;
  int64_t n = 0, S = 0, blk = 0, Kpad = 0;
  int nranges = 0;
  std::vector<int32_t> pos, paint;
  std::vector<uint8_t> r, mask;
  std::vector<double> w;
  std::vector<Group> groups;           // [nranges][4] (P = 1..4)
  std::vector<int32_t> range_slot0;    // [nranges + 1]
  std::vector<int32_t> slot_snp;       // global slot -> snp or -1
  int64_t nslots = 0, nrows = 0;
  double neff = 0, scale = 0;          // weight = W / scale
  int32_t neffH = 0, neffL = 0;
  uint32_t sa = 0, sb = 0;             // epilogue count unit: t = (H << sa) + (L >> sb)
  uint32_t M = 0;                      // pseudocount 0.5 in count units
  bool pos_sorted = true;
  // device, static
  DevBuf d_codes, d_w, d_p64, d_rec, d_r, d_mask, d_pos, d_paint, d_ops, d_dig;
  const uint8_t* codes_dev = nullptr;  // the class matrix on this device: d_codes, or a buffer the device group owns
  TmapSet tm;
  std::vector<BlockResult> results;
  double t_pack_ms = 0;
};

namespace {

// ---------------------------------------------------------------------------------------------- plan
// codes: host class matrix (uploaded here), or NULL when `codes_dev` already holds it on this device (multi-GPU: the
// group uploads once and broadcasts, ldw_group_load_codes)
int build_plan(ldw_mi_plan* P, const uint8_t* codes, const uint8_t* codes_dev, const double* hdw, const int32_t* pos,
               const int32_t* paint) {
  ldw_ctx* ctx = P->ctx;
  cudaStream_t st = ctx->stream;
  const int64_t n = P->n, S = P->S, blk = P->blk;
  cudaEvent_t e0, e1;
  LDW_CUDA(cudaEventCreate(&e0));
  LDW_CUDA(cudaEventCreate(&e1));
  LDW_CUDA(cudaEventRecord(e0, st));
  const bool dbg_t = getenv("LDW_DBG_TIMING") != nullptr;
  auto bt0 = std::chrono::steady_clock::now();
  auto blap = [&](const char* what) {
    if (!dbg_t) return;
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "ldw timing: plan %-22s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - bt0).count());
    bt0 = t;
  };
  P->pos.assign(pos, pos + n);
  P->paint.assign(paint, paint + n);
  P->w.assign(hdw, hdw + S);
  for (int64_t i = 1; i < n; i++)
    if (pos[i] < pos[i - 1]) P->pos_sorted = false;
  // ---- fixed-point weights: W_s = round(w_s * scale), scale = (2^28 - 1) / max(w), two 14-bit halves H, L.
  //      Each half h is written as 255 a - b with byte digits a = ceil(h / 255) <= 65, b = 255 a - h <= 254: the
  //      one-hot operand stores 0xFF where the allele matches, which the tensor core reads as +255 (unsigned pass,
  //      digit a) or -1 (signed pass, digit b) -- one operand array, no shifted copy, and the byte doubles as the
  //      select mask of the expander warps.  Partial sums stay below 65 * 255 * nseq < 2^31.
  double wmax = 0, neff = 0;
  for (int64_t s = 0; s < S; s++) {
    if (!(hdw[s] > 0) || !std::isfinite(hdw[s])) return set_error(LDW_ERR_ARG, "hdw[%lld] = %g is not a positive finite weight", (long long)s, hdw[s]);
    wmax = std::max(wmax, hdw[s]);
    neff += hdw[s];
  }
  P->neff = neff;
  {
    // count unit of the epilogue: t = (H << sa) + (L >> sb), sa + sb = 14, i.e. 2^sb weight units.  t must hold the total
    // of a joint table, (neff + r r' / 2) * scale <= (neff + 12.5) * 2^28 / max(w) weight units, in 32 bits -- sized from
    // the ACTUAL total weight, not from the worst case nseq * max(w): a population of a few large clusters (neff of a
    // handful at thousands of sequences) would otherwise lose 8-10 low bits of every cell to the floor below, which showed
    // as 1.1e-6 on MI with one clonal cluster of 5000 (tests/test_gpu_mi.py).
    const double total = (neff + 12.5) * 268435456.0 / wmax;
    int bits = 0;
    while (std::ldexp(1.0, bits) <= total) bits++;
    int sb_ = std::max(0, bits - 32);
    if (sb_ > 14) sb_ = 14;
    P->sb = (uint32_t)sb_;
    P->sa = (uint32_t)(14 - sb_);
  }
  // weights are W_s = round(w_s * scale) <= 2^28 - 1, with the scale chosen such that the pseudocount 0.5 is an
  // integer number M of count units (2^sb weight units each): scale = M * 2^(sb+1)
  {
    const double unit2 = std::ldexp(1.0, (int)P->sb + 1);
    double M = std::floor(268435455.0 / wmax / unit2);
    if (M < 1) return set_error(LDW_ERR_UNSUPPORTED, "largest weight %g too large for the fixed-point count unit", wmax);
    if (M > 1073741823.0) M = 1073741823.0;
    P->M = (uint32_t)M;
    P->scale = M * unit2;
    // all cells of a table plus their pseudocounts must fit 32 bits: (neff + 12.5) * scale / 2^sb < 2^32
    if ((neff + 12.5) * P->scale / std::ldexp(1.0, (int)P->sb) >= 4294967295.0) {
      // shrink the unit until it fits (loses weight bits only for pathological weight vectors, e.g. neff << 1)
      while (P->M > 1 && (neff + 12.5) * (P->M * unit2) / std::ldexp(1.0, (int)P->sb) >= 4294967295.0) P->M >>= 1;
      P->scale = P->M * unit2;
    }
  }
  P->Kpad = round_up(S, 128);
  std::vector<uint8_t> da(16384), db(16384);
  {
    for (int h = 0; h < 16384; h++) {
      int a = (h + 254) / 255, b = 255 * a - h;
      da[h] = (uint8_t)a;
      db[h] = (uint8_t)b;
    }
  }
  std::vector<uint8_t> dig((size_t)4 * P->Kpad, 0);  // rows: aH, aL, bH, bL (shared-memory slot order)
  std::vector<int32_t> wH(S), wL(S);
  int64_t sumH = 0, sumL = 0;
  // Rounding with error diffusion along the sequences ordered by weight: the fixed-point unit is set by the LARGEST
  // weight, so the members of a big clonal cluster (thousands of equal weights 1/(m+1), a singleton at 0.5 elsewhere) carry
  // only ~17 significant bits each, and round-to-nearest would give all of them the SAME rounding error -- it adds up
  // coherently in every joint count (measured: 1.1e-6 on MI at m = 5000).  Diffusing the residual to the next sequence of
  // the same (or the nearest) weight keeps every weight within one unit of its value, makes each cluster's total exact to
  // half a unit and turns the error of a subset sum into a zero-mean walk of ~0.3 sqrt(m) units instead of a bias of up to m / 2.
  std::vector<int64_t> Wq(S);
  {
    std::vector<int32_t> order(S);
    for (int64_t s = 0; s < S; s++) order[s] = (int32_t)s;
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return hdw[a] < hdw[b]; });
    double carry = 0.0;
    for (int64_t k = 0; k < S; k++) {
      const int32_t s = order[k];
      const double x = hdw[s] * P->scale + carry;
      int64_t W = (int64_t)std::llround(x);
      if (W < 0) W = 0;
      if (W > 268435455) W = 268435455;
      carry = x - (double)W;
      if (carry > 1.0) carry = 1.0;
      if (carry < -1.0) carry = -1.0;
      Wq[s] = W;
    }
  }
  for (int64_t s = 0; s < S; s++) {
    int64_t W = Wq[s];
    int32_t H = (int32_t)(W >> 14), L = (int32_t)(W & 16383);
    wH[s] = H; wL[s] = L;
    sumH += H; sumL += L;
    dig[0 * P->Kpad + s] = da[H];
    dig[1 * P->Kpad + s] = da[L];
    dig[2 * P->Kpad + s] = db[H];
    dig[3 * P->Kpad + s] = db[L];
  }
  if (sumH > 0x7fffffffLL || sumL > 0x7fffffffLL) return set_error(LDW_ERR_UNSUPPORTED, "too many sequences for int32 accumulation (nseq=%lld)", (long long)S);
  P->neffH = (int32_t)sumH;
  P->neffL = (int32_t)sumL;
  blap("weights/digits");
  // ---- upload codes, per-SNP allele statistics
  if (codes) {
    LDW_TRY(P->d_codes.alloc((size_t)n * S));
    P->codes_dev = static_cast<const uint8_t*>(P->d_codes.p);
  } else {
    P->codes_dev = codes_dev;
  }
  LDW_TRY(P->d_mask.alloc((size_t)n));
  LDW_TRY(P->d_r.alloc((size_t)n));
  DevBuf d_table;
  LDW_TRY(d_table.alloc((size_t)n * 5 * 4));
  if (codes) LDW_CUDA(cudaMemcpyAsync(P->d_codes.p, codes, (size_t)n * S, cudaMemcpyHostToDevice, st));
  LDW_TRY(snp_allele_stats(st, P->codes_dev, n, S, d_table.as<int32_t>(), P->d_mask.as<uint8_t>(), P->d_r.as<uint8_t>()));
  P->r.resize(n);
  P->mask.resize(n);
  LDW_CUDA(cudaMemcpyAsync(P->r.data(), P->d_r.p, (size_t)n, cudaMemcpyDeviceToHost, st));
  LDW_CUDA(cudaMemcpyAsync(P->mask.data(), P->d_mask.p, (size_t)n, cudaMemcpyDeviceToHost, st));
  LDW_CUDA(cudaStreamSynchronize(st));
  blap("upload + allele stats");
  for (int64_t i = 0; i < n; i++)
    if (P->r[i] < 2)
      return set_error(LDW_ERR_UNSUPPORTED, "SNP %lld is monomorphic (r=%d); the reference's filters never retain such sites "
                       "and this implementation does not support them", (long long)i, (int)P->r[i]);

  if (n >= (1 << 28)) return set_error(LDW_ERR_UNSUPPORTED, "nsnp >= 2^28 not supported");
  // ---- slots: per block range, SNPs grouped by plane count P = r - 1, each group padded to 128
  P->nranges = (int)((n + blk - 1) / blk);
  P->groups.assign((size_t)P->nranges * 4, Group());
  P->range_slot0.assign(P->nranges + 1, 0);
  int64_t slot = 0, row = 0;
  std::vector<int32_t> snp_slot(n);
  for (int b = 0; b < P->nranges; b++) {
    P->range_slot0[b] = (int32_t)slot;
    int64_t lo = (int64_t)b * blk, hi = std::min<int64_t>(n, lo + blk);
    for (int pc = 1; pc <= 4; pc++) {
      Group& G = P->groups[(size_t)b * 4 + pc - 1];
      G.slot0 = (int32_t)slot;
      G.row0 = (int32_t)row;
      int cnt = 0;
      for (int64_t i = lo; i < hi; i++)
        if (P->r[i] - 1 == pc) snp_slot[i] = (int32_t)(slot + cnt++);
      G.count = cnt;
      G.nslots = (int32_t)round_up(cnt, 128);
      slot += G.nslots;
      row += (int64_t)G.nslots * pc;
      if (slot > 0x3fffffff || row > 0x3fffffff) return set_error(LDW_ERR_UNSUPPORTED, "problem too large for 32-bit slot indices");
    }
  }
  P->range_slot0[P->nranges] = (int32_t)slot;
  P->nslots = slot;
  P->nrows = row;
  P->slot_snp.assign(slot, -1);
  for (int64_t i = 0; i < n; i++) P->slot_snp[snp_slot[i]] = (int32_t)i;
  // operand row -> (snp, allele)
  std::vector<uint32_t> row_info((size_t)std::max<int64_t>(row, 1), 0xFFFFFFFFu);
  for (int b = 0; b < P->nranges; b++)
    for (int pc = 1; pc <= 4; pc++) {
      const Group& G = P->groups[(size_t)b * 4 + pc - 1];
      for (int k = 0; k < G.count; k++) {
        int32_t snp = P->slot_snp[G.slot0 + k];
        int m = P->mask[snp], q = 0;
        for (int a = 0; a < 5 && q < pc; a++)
          if (m & (1 << a)) {
            row_info[(size_t)G.row0 + (size_t)q * G.nslots + k] = (uint32_t)snp | ((uint32_t)a << 28);
            q++;
          }
      }
    }
  blap("slots + row_info (host)");
  // ---- device: weights, records, operands
  DevBuf d_slot_snp, d_wH, d_wL, d_row_info;
  LDW_TRY(P->d_w.alloc((size_t)S * 8));
  LDW_TRY(d_wH.alloc((size_t)S * 4));
  LDW_TRY(d_wL.alloc((size_t)S * 4));
  LDW_TRY(d_slot_snp.alloc((size_t)std::max<int64_t>(slot, 1) * 4));
  LDW_TRY(d_row_info.alloc(row_info.size() * 4));
  LDW_TRY(P->d_p64.alloc((size_t)n * 5 * 8));
  LDW_TRY(P->d_rec.alloc((size_t)std::max<int64_t>(slot, 1) * 4 * sizeof(Rec)));
  LDW_TRY(P->d_pos.alloc((size_t)n * 4));
  LDW_TRY(P->d_paint.alloc((size_t)n * 4));
  LDW_CUDA(cudaMemcpyAsync(P->d_w.p, hdw, (size_t)S * 8, cudaMemcpyHostToDevice, st));
  LDW_CUDA(cudaMemcpyAsync(d_wH.p, wH.data(), (size_t)S * 4, cudaMemcpyHostToDevice, st));
  LDW_CUDA(cudaMemcpyAsync(d_wL.p, wL.data(), (size_t)S * 4, cudaMemcpyHostToDevice, st));
  LDW_CUDA(cudaMemcpyAsync(d_slot_snp.p, P->slot_snp.data(), (size_t)slot * 4, cudaMemcpyHostToDevice, st));
  LDW_CUDA(cudaMemcpyAsync(d_row_info.p, row_info.data(), row_info.size() * 4, cudaMemcpyHostToDevice, st));
  LDW_CUDA(cudaMemcpyAsync(P->d_pos.p, pos, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  LDW_CUDA(cudaMemcpyAsync(P->d_paint.p, paint, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  {
    int wpb = 8;
    mi_build_rec_kernel<<<(unsigned)((slot + wpb - 1) / wpb), wpb * 32, 0, st>>>(
        P->codes_dev, S, d_slot_snp.as<int32_t>(), slot, P->d_mask.as<uint8_t>(), P->d_w.as<double>(),
        d_wH.as<int32_t>(), d_wL.as<int32_t>(), P->d_rec.as<Rec>(), slot, P->d_p64.as<double>(), P->sa, P->sb, P->M);
    LDW_CUDA(cudaGetLastError());
  }
  size_t op_bytes = (size_t)std::max<int64_t>(row, 128) * P->Kpad;
  LDW_TRY(P->d_ops.alloc(op_bytes));
  LDW_TRY(P->d_dig.alloc(dig.size()));
  LDW_CUDA(cudaMemcpyAsync(P->d_dig.p, dig.data(), dig.size(), cudaMemcpyHostToDevice, st));
  {
    int64_t total = row * (P->Kpad / 16);
    if (total > 0) {
      mi_pack_operands_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(P->codes_dev, S, P->Kpad,
                                                                               d_row_info.as<uint32_t>(), row,
                                                                               P->d_ops.as<uint8_t>());
      LDW_CUDA(cudaGetLastError());
    }
  }
  uint64_t trows = (uint64_t)std::max<int64_t>(row, 128);
  LDW_TRY(make_tmap_u8_sw128(&P->tm.a, P->d_ops.p, trows, (uint64_t)P->Kpad, 128));
  for (int j = 0; j < 4; j++) LDW_TRY(make_tmap_u8_sw128(&P->tm.b[j], P->d_ops.p, trows, (uint64_t)P->Kpad, 64u >> j));
  {
    ScanWS* W = get_ws(ctx);
    if (!W->events_ready) {
      for (auto& b : W->ring) {
        LDW_CUDA(cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming));
        LDW_CUDA(cudaEventCreateWithFlags(&b.done2, cudaEventDisableTiming));
        LDW_CUDA(cudaEventCreateWithFlags(&b.uploaded, cudaEventDisableTiming));
      }
      for (auto& l : W->lr) {
        LDW_CUDA(cudaEventCreateWithFlags(&l.scan_done, cudaEventDisableTiming));
        LDW_CUDA(cudaEventCreateWithFlags(&l.sel_done, cudaEventDisableTiming));
      }
      W->events_ready = true;
    }
  }
  LDW_CUDA(cudaEventRecord(e1, st));
  LDW_CUDA(cudaStreamSynchronize(st));
  blap("records + pack + tmaps");
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  P->t_pack_ms = ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  LDW_CUDA(cudaFuncSetAttribute(mi_scan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MI_SMEM_BYTES));
  LDW_CUDA(cudaFuncSetAttribute(mi_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MI_SMEM_BYTES));
  return 0;
}

// ---------------------------------------------------------------------------------------------- per-block host prep
struct ScanCfg {
  double g, sr_dist, lr_retain, lr_approx;
  int flags;
  int dense = 0;  // debug: every (row, col) cell is wanted, so no tile is skipped
};

// R's circular distance (R/computePairwiseMI.R:330) for integer positions
inline double circ_len_d(double p1, double p2, double g) {
  double m = std::fmod(p1 - p2, g);
  if (m < 0) m += g;
  return 0.5 * g - std::fabs(m - 0.5 * g);
}

// Builds the block's from/to lists (Q12 drop in SR-only mode), local-index maps, short-range column structure
// and tile list.  Returns 1 if the block is empty (nothing to do).
int prepare_block(const ldw_mi_plan* P, int bf, int bt, const ScanCfg& cfg, BlockHost& H, bool sizes_only = false) {
  const int64_t n = P->n, blk = P->blk;
  const bool sr_only = (cfg.flags & LDW_SCAN_SR_ONLY) != 0;
  const double g = cfg.g, sr = cfg.sr_dist;
  int64_t flo = (int64_t)bf * blk, fhi = std::min<int64_t>(n, flo + blk);
  int64_t tlo = (int64_t)bt * blk, thi = std::min<int64_t>(n, tlo + blk);
  H.from_idx.clear();
  H.to_idx.clear();
  if (!sr_only) {
    for (int64_t i = flo; i < fhi; i++) H.from_idx.push_back((int32_t)i);
    for (int64_t i = tlo; i < thi; i++) H.to_idx.push_back((int32_t)i);
  } else {
    // R/computePairwiseMI.R:182-185: keep a SNP iff some SNP of the opposite list is strictly closer than sr_dist
    auto keep = [&](int64_t lo, int64_t hi, int64_t olo, int64_t ohi, std::vector<int32_t>& out) {
      for (int64_t i = lo; i < hi; i++) {
        double x = P->pos[i];
        bool k = false;
        // positions are sorted: nearest candidates are around lower_bound(x) and, circularly, the two ends
        auto it = std::lower_bound(P->pos.begin() + olo, P->pos.begin() + ohi, (int32_t)x);
        int64_t c = it - P->pos.begin();
        for (int64_t j : {c - 1, c, olo, ohi - 1})
          if (j >= olo && j < ohi && std::fabs(circ_len_d(P->pos[j], x, g)) < sr) k = true;
        if (k) out.push_back((int32_t)i);
      }
    };
    keep(flo, fhi, tlo, thi, H.from_idx);
    keep(tlo, thi, flo, fhi, H.to_idx);
  }
  const int32_t nf = (int32_t)H.from_idx.size(), nt = (int32_t)H.to_idx.size();
  H.nf = nf; H.nt = nt;
  H.diag = (nf == nt) && std::equal(H.from_idx.begin(), H.from_idx.end(), H.to_idx.begin());  // fromISto :198-202
  H.ragged = (nf != nt);
  H.n_pairs = H.n_sr = H.n_lr = 0;
  if (nf == 0 || nt == 0) return 1;
  H.n_pairs = H.diag ? (int64_t)nf * (nf - 1) / 2 : (int64_t)nf * nt - std::min(nf, nt);

  const int32_t fs0 = P->range_slot0[bf], fs1 = P->range_slot0[bf + 1];
  const int32_t ts0 = P->range_slot0[bt], ts1 = P->range_slot0[bt + 1];
  // ---- short-range structure per column (two-pointer sweeps; positions ascending within both lists)
  H.colinfo.assign(nt, ColInfo{0, 0, 0, 0, 0, 0});
  const bool all_sr = !(sr < 0.5 * g);  // len <= g/2 always
  {
    int64_t a0 = 0, a1 = 0, w1 = 0, w0 = 0;
    uint64_t accU = 0, accL = 0;
    std::vector<uint32_t> cu(nt), cl(nt);
    for (int j = 0; j < nt; j++) {
      const double pj = P->pos[H.to_idx[j]];
      int32_t ia0, ia1, ib0, ib1;
      if (all_sr) {
        ia0 = 0; ia1 = nf; ib0 = ib1 = 0;
      } else {
        while (a0 < nf && (double)P->pos[H.from_idx[a0]] < pj - sr) a0++;          // first row with pf >= pj - sr
        while (a1 < nf && (double)P->pos[H.from_idx[a1]] <= pj + sr) a1++;         // first row with pf >  pj + sr
        while (w1 < nf && (double)P->pos[H.from_idx[w1]] <= pj - (g - sr)) w1++;   // rows [0, w1): wrap from the left
        while (w0 < nf && (double)P->pos[H.from_idx[w0]] < pj + (g - sr)) w0++;    // rows [w0, nf): wrap from the right
        // intervals in ascending order: W1 = [0, w1), A = [a0, a1), W2 = [w0, nf); at most one wrap side is non-empty
        int32_t x0 = (int32_t)a0, x1 = (int32_t)std::max(a0, a1);
        if (w1 > 0) {
          if ((int32_t)w1 >= x0) { ia0 = 0; ia1 = std::max<int32_t>((int32_t)w1, x1); ib0 = ib1 = 0; }
          else { ia0 = 0; ia1 = (int32_t)w1; ib0 = x0; ib1 = x1; }
        } else if (w0 < nf) {
          if ((int32_t)w0 <= x1) { ia0 = x0; ia1 = nf; if ((int32_t)w0 < x0) ia0 = (int32_t)w0; ib0 = ib1 = 0; }
          else { ia0 = x0; ia1 = x1; ib0 = (int32_t)w0; ib1 = nf; }
        } else {
          ia0 = x0; ia1 = x1; ib0 = ib1 = 0;
        }
      }
      ColInfo& c = H.colinfo[j];
      c.a0 = ia0; c.a1 = ia1; c.b0 = ib0; c.b1 = ib1;
      auto below = [&](int x) {
        return std::min(std::max(x - ia0, 0), ia1 - ia0) + std::min(std::max(x - ib0, 0), ib1 - ib0);
      };
      int tot = (ia1 - ia0) + (ib1 - ib0);
      cu[j] = H.diag ? 0u : (uint32_t)below(std::min(j, nf));
      cl[j] = (uint32_t)(tot - below(std::min(j + 1, nf)));
      accU += cu[j];
      accL += cl[j];
    }
    if (accU + accL > 0xFFFFFFF0ull) return set_error(LDW_ERR_UNSUPPORTED, "more than 2^32 short-range links in one block");
    uint32_t ru = 0, rl = (uint32_t)accU;
    for (int j = 0; j < nt; j++) {
      H.colinfo[j].baseU = ru;
      H.colinfo[j].baseL = rl;
      ru += cu[j];
      rl += cl[j];
    }
    H.n_sr = (int64_t)(accU + accL);
    H.n_lr = H.n_pairs - H.n_sr;
  }
  if (sizes_only) return 0;
  // ---- local-index maps over the slot ranges
  H.rfl.resize(nf);
  H.rtl.resize(nt);
  for (int k = 0; k < nf; k++) H.rfl[k] = P->r[H.from_idx[k]];
  for (int k = 0; k < nt; k++) H.rtl[k] = P->r[H.to_idx[k]];
  std::vector<int32_t> il_of(fhi - flo, -1), jl_of(thi - tlo, -1);
  for (int k = 0; k < nf; k++) il_of[H.from_idx[k] - flo] = k;
  for (int k = 0; k < nt; k++) jl_of[H.to_idx[k] - tlo] = k;
  H.rowdyn.assign(fs1 - fs0, RowDyn{-1, 0.f});
  for (int32_t s = fs0; s < fs1; s++) {
    int32_t snp = P->slot_snp[s];
    if (snp < 0) continue;
    int32_t il = il_of[snp - flo];
    RowDyn rd;
    rd.il = il;
    rd.rtl = (il >= 0 && il < nt) ? (float)H.rtl[il] : 0.f;
    H.rowdyn[s - fs0] = rd;
  }
  H.coldyn.assign(ts1 - ts0, ColDyn{-1, 0.f, 0, 0, 0, 0, 0, 0});
  for (int32_t s = ts0; s < ts1; s++) {
    int32_t snp = P->slot_snp[s];
    if (snp < 0) continue;
    int32_t jl = jl_of[snp - tlo];
    ColDyn cd{-1, 0.f, 0, 0, 0, 0, 0, 0};
    cd.jl = jl;
    if (jl >= 0) {
      cd.rfl = jl < nf ? (float)H.rfl[jl] : 0.f;
      const ColInfo& c = H.colinfo[jl];
      cd.a0 = c.a0; cd.a1 = c.a1; cd.b0 = c.b0; cd.b1 = c.b1; cd.baseU = c.baseU; cd.baseL = c.baseL;
    }
    H.coldyn[s - ts0] = cd;
  }

  // ---- tiles.  Per 16-slot chunk: first / last valid slot (SNP index and local index ascend with the slot
  // inside a group), so a tile's extent costs a handful of look-ups.
  const int nfc = (fs1 - fs0) / 16, ntc = (ts1 - ts0) / 16;
  std::vector<int32_t> f_first(nfc, -1), f_last(nfc, -1), t_first(ntc, -1), t_last(ntc, -1);
  for (int c = 0; c < nfc; c++)
    for (int k = 0; k < 16; k++)
      if (H.rowdyn[c * 16 + k].il >= 0) { if (f_first[c] < 0) f_first[c] = c * 16 + k; f_last[c] = c * 16 + k; }
  for (int c = 0; c < ntc; c++)
    for (int k = 0; k < 16; k++)
      if (H.coldyn[c * 16 + k].jl >= 0) { if (t_first[c] < 0) t_first[c] = c * 16 + k; t_last[c] = c * 16 + k; }
  auto extent = [](const std::vector<int32_t>& first, const std::vector<int32_t>& last, int off0, int nsl, int& lo, int& hi) {
    lo = hi = -1;
    for (int c = off0 / 16; c < (off0 + nsl) / 16; c++)
      if (first[c] >= 0) { if (lo < 0) lo = first[c]; hi = last[c]; }
  };
  // One tile = 128 row SNPs x NJ column SNPs of one (PA, PB) kind; tiles that hold no wanted pair are skipped.
  H.tiles.clear();
  for (int pa = 1; pa <= 4; pa++) {
    const Group& Gi = P->groups[(size_t)bf * 4 + pa - 1];
    if (Gi.count == 0) continue;
    for (int pb = 1; pb <= 4; pb++) {
      const Group& Gj = P->groups[(size_t)bt * 4 + pb - 1];
      if (Gj.count == 0) continue;
      const int njl = mi_njlog2(pa, pb), NJ = 1 << njl;
      const int n_ti = (Gi.count + 127) / 128, n_tj = (Gj.count + NJ - 1) / NJ;
      // per row tile / column tile extents
      struct Ext { int lo, hi, lmin, lmax; double p0, p1; };
      std::vector<Ext> ei(n_ti), ej(n_tj);
      for (int ti = 0; ti < n_ti; ti++) {
        Ext& e = ei[ti];
        extent(f_first, f_last, Gi.slot0 - fs0 + ti * 128, 128, e.lo, e.hi);
        if (e.lo >= 0) {
          e.lmin = H.rowdyn[e.lo].il; e.lmax = H.rowdyn[e.hi].il;
          e.p0 = P->pos[P->slot_snp[fs0 + e.lo]]; e.p1 = P->pos[P->slot_snp[fs0 + e.hi]];
        }
      }
      for (int tj = 0; tj < n_tj; tj++) {
        Ext& e = ej[tj];
        extent(t_first, t_last, Gj.slot0 - ts0 + tj * NJ, NJ, e.lo, e.hi);
        if (e.lo >= 0) {
          e.lmin = H.coldyn[e.lo].jl; e.lmax = H.coldyn[e.hi].jl;
          e.p0 = P->pos[P->slot_snp[ts0 + e.lo]]; e.p1 = P->pos[P->slot_snp[ts0 + e.hi]];
        }
      }
      auto make_tile = [&](int ti, int tj, uint32_t flags) {
        TileDesc td;
        memset(&td, 0, sizeof(td));
        td.a_row0 = Gi.row0 + ti * 128;
        td.a_pstride = Gi.nslots;
        td.b_row0 = Gj.row0 + tj * NJ;
        td.b_pstride = Gj.nslots;
        td.i_slot0 = Gi.slot0 + ti * 128;
        td.j_slot0 = Gj.slot0 + tj * NJ;
        td.i_dyn0 = td.i_slot0 - fs0;
        td.j_dyn0 = td.j_slot0 - ts0;
        td.PA = (uint8_t)pa; td.PB = (uint8_t)pb; td.njlog2 = (uint8_t)njl;
        td.flags = (uint8_t)flags;
        return td;
      };
      // 0: not wanted, else TILE flags | 0x80 marker
      auto wanted = [&](int ti, int tj, uint32_t& flags) -> bool {
        if (ti >= n_ti || tj >= n_tj) return false;
        const Ext &a = ei[ti], &b = ej[tj];
        if (a.lo < 0 || b.lo < 0) return false;
        if (H.diag && !cfg.dense && a.lmax <= b.lmin) return false;  // no pair with row > col
        const double gap = std::max(0.0, std::max(a.p0, b.p0) - std::min(a.p1, b.p1));
        const double span = std::max(a.p1, b.p1) - std::min(a.p0, b.p0);
        const bool has_sr = all_sr || gap <= sr || (g - span) <= sr;
        if (sr_only && !has_sr) return false;
        flags = has_sr ? TILE_HAS_SR : 0;
        return true;
      };
      // row tiles two at a time: one tile of a CTA pair (cta_group::2 MMAs, M = 256)
      for (int ti = 0; ti < n_ti; ti += 2)
        for (int tj = 0; tj < n_tj; tj++) {
          uint32_t fl0 = 0, fl1 = 0;
          const bool w0 = wanted(ti, tj, fl0), w1 = wanted(ti + 1, tj, fl1);
          if (!w0 && !w1) continue;
          TileDesc td = make_tile(ti, tj, w0 ? fl0 : (uint32_t)TILE_NULL);
          td.flags1 = (uint8_t)(w1 ? fl1 : (uint32_t)TILE_NULL);
          H.tiles.push_back(td);
        }
    }
  }
  // Order of the kinds: the large-stage geometries first, inside a geometry the rare kinds first, the most numerous last.  The CTA pairs take tiles round-robin, so the
  // launch ends on many equal tiles instead of a handful of expensive odd ones (4 x 4 planes: both accumulator buffers, two
  // 96 KB stages) that leave most SMs idle behind the last of them.  Tiles of a kind stay contiguous and in their order
  // (neighbouring CTAs share a row tile through L2); results do not depend on the order.
  {
    size_t cnt[16] = {0}, beg[16];
    for (const TileDesc& t : H.tiles) cnt[(t.PA - 1) * 4 + (t.PB - 1)]++;
    int ord[16];
    for (int k = 0; k < 16; k++) ord[k] = k;
    // (kinds of one stage geometry stay together: a change of geometry drains the pipeline)
    auto geo = [](int k) { return mi_stage_geo(k / 4 + 1, k % 4 + 1, mi_njlog2(k / 4 + 1, k % 4 + 1)); };
    std::stable_sort(ord, ord + 16, [&](int a, int b) { return geo(a) != geo(b) ? geo(a) > geo(b) : cnt[a] < cnt[b]; });
    size_t off = 0;
    for (int k = 0; k < 16; k++) { beg[ord[k]] = off; off += cnt[ord[k]]; }
    std::vector<TileDesc> sorted(H.tiles.size());
    for (const TileDesc& t : H.tiles) sorted[beg[(t.PA - 1) * 4 + (t.PB - 1)]++] = t;
    H.tiles.swap(sorted);
  }
  // pilot sample: up to 64 tiles spread evenly over the list (kinds in proportion), appended as copies
  H.n_real_tiles = (int32_t)H.tiles.size();
  if (!cfg.dense && H.n_real_tiles >= 1024) {
    const int np = 64;
    for (int k = 0; k < np; k++) H.tiles.push_back(H.tiles[(size_t)((2 * k + 1) * (int64_t)H.n_real_tiles / (2 * np))]);
  }
  return 0;
}

// All per-block arrays go through one pinned staging buffer per ring entry, so the copies are truly
// asynchronous and the host can prepare the next block while the device works on this one.
int upload_block(cudaStream_t st, BlockDev& D, const BlockHost& H) {  // st: the stream the copies are issued on
  struct Part { DevBuf* d; const void* src; size_t bytes; };
  Part parts[8] = {{&D.rowdyn, H.rowdyn.data(), H.rowdyn.size() * sizeof(RowDyn)},
                   {&D.coldyn, H.coldyn.data(), H.coldyn.size() * sizeof(ColDyn)},
                   {&D.colinfo, H.colinfo.data(), H.colinfo.size() * sizeof(ColInfo)},
                   {&D.tiles, H.tiles.data(), H.tiles.size() * sizeof(TileDesc)},
                   {&D.from_idx, H.from_idx.data(), H.from_idx.size() * 4},
                   {&D.to_idx, H.to_idx.data(), H.to_idx.size() * 4},
                   {&D.rfl, H.rfl.data(), H.rfl.size()},
                   {&D.rtl, H.rtl.data(), H.rtl.size()}};
  size_t total = 0;
  for (auto& p : parts) total += (p.bytes + 255) & ~(size_t)255;
  LDW_TRY(D.stage.ensure(total + 256));
  size_t off = 0;
  for (auto& p : parts) {
    LDW_TRY(p.d->ensure(std::max<size_t>(p.bytes, 16)));
    if (p.bytes) {
      memcpy(D.stage.as<uint8_t>() + off, p.src, p.bytes);
      LDW_CUDA(cudaMemcpyAsync(p.d->p, D.stage.as<uint8_t>() + off, p.bytes, cudaMemcpyHostToDevice, st));
    }
    off += (p.bytes + 255) & ~(size_t)255;
  }
  return 0;
}

void fill_scan_params(const ldw_mi_plan* P, const BlockDev& D, const BlockHost& H, const ScanCfg& cfg, ScanParams& sp) {
  memset(&sp, 0, sizeof(sp));
  sp.tiles = D.tiles.as<TileDesc>();
  sp.n_tiles = H.n_real_tiles;
  sp.rec = P->d_rec.as<Rec>();
  sp.rec_vstride = P->nslots;
  sp.rowdyn = D.rowdyn.as<RowDyn>();
  sp.coldyn = D.coldyn.as<ColDyn>();
  sp.nkb = (int32_t)(P->Kpad / 128);
  sp.nf = H.nf; sp.nt = H.nt;
  sp.diag = H.diag; sp.ragged = H.ragged;
  sp.qcorr = (!H.diag && !(cfg.flags & LDW_SCAN_IDEAL_Q)) ? 1 : 0;
  sp.sr_only = (cfg.flags & LDW_SCAN_SR_ONLY) ? 1 : 0;
  sp.dig = P->d_dig.as<uint8_t>();
  sp.kpad = P->Kpad;
  sp.rfl_arr = D.rfl.as<uint8_t>();
  sp.rtl_arr = D.rtl.as<uint8_t>();
  sp.sa = P->sa;
  sp.sb = P->sb;
  sp.kT = (float)(std::ldexp(1.0, (int)P->sb) / P->scale);  // == 0.5 / M
  sp.M = P->M;
  for (int a = 0; a < 4; a++)
    for (int b = 0; b < 4; b++) {
      double ri = a + 2, rj = b + 2;
      double den = P->neff + 0.5 * ri * rj;
      const float den_f = (float)den, rkT = 1.0f / sp.kT, rden = 1.0f / den_f;
      sp.scale[a][b] = (float)(M_LN2 / den) * sp.kT;
      sp.q0s[a][b] = (float)(0.25 * ri * rj / den) * rkT;
      sp.qod[a][b] = (float)(0.25 / den) * rkT;
      sp.rp_qc[a][b] = rden * rkT;
      sp.rp_plain[a][b] = den_f * sp.kT;
    }
}

// Launch the scan kernel: persistent, one CTA pair (cluster of two) per TPC.
// `reserve`: SMs left free for the single-CTA selection kernels of earlier blocks (the persistent scan CTAs hold every
// SM they run on for the whole launch, so anything else would otherwise wait for the gap between two scans).
int launch_scan(const ldw_mi_plan* P, const ScanParams& sp, cudaStream_t st, int reserve = 0) {
  if (sp.n_tiles <= 0) return 0;
  // one CTA pair (a cluster of two: the two SMs of a TPC) per tile
  int grid = 2 * std::min<int>(sp.n_tiles, std::max(1, (P->ctx->num_sms - reserve) / 2));
  if (sp.dbg) mi_scan_kernel<true><<<grid, MI_THREADS, MI_SMEM_BYTES, st>>>(P->tm, sp);
  else mi_scan_kernel<false><<<grid, MI_THREADS, MI_SMEM_BYTES, st>>>(P->tm, sp);
  LDW_CUDA(cudaGetLastError());
  return 0;
}

RefineParams make_refine_params(const ldw_mi_plan* P, const BlockDev& D, const BlockHost& H, const ScanCfg& cfg) {
  RefineParams R;
  R.codes = P->codes_dev;
  R.S = P->S;
  R.w = P->d_w.as<double>();
  R.p64 = P->d_p64.as<double>();
  R.r = P->d_r.as<uint8_t>();
  R.mask = P->d_mask.as<uint8_t>();
  R.from_idx = D.from_idx.as<int32_t>();
  R.to_idx = D.to_idx.as<int32_t>();
  R.rfl_arr = D.rfl.as<uint8_t>();
  R.rtl_arr = D.rtl.as<uint8_t>();
  R.nf = H.nf; R.nt = H.nt;
  R.ideal_q = (cfg.flags & LDW_SCAN_IDEAL_Q) ? 1 : 0;
  R.neff = P->neff;
  return R;
}

int validate_scan(const ldw_mi_plan* P, const ScanCfg& cfg) {
  if (!(cfg.g > 0) || cfg.g != std::floor(cfg.g) || cfg.g > 9.0e15) return set_error(LDW_ERR_ARG, "genome length g=%g must be a positive integer", cfg.g);
  if (!(cfg.sr_dist >= 0)) return set_error(LDW_ERR_ARG, "sr_dist must be >= 0");
  if (!P->pos_sorted) return set_error(LDW_ERR_UNSUPPORTED, "POS must be non-decreasing (the short-range slot layout relies on it)");
  if (P->n > 0 && ((double)P->pos.back() - (double)P->pos.front()) >= cfg.g) return set_error(LDW_ERR_UNSUPPORTED, "POS span must be smaller than the genome length g");
  return 0;
}

}  // namespace

// ================================================================================================== C ABI
extern "C" {

int ldw_mi_plan_create(ldw_ctx* ctx, const uint8_t* codes, int64_t n_snp, int64_t nseq, const double* hdw,
                       const int32_t* pos, const int32_t* paint, int64_t blk, ldw_mi_plan** out) {
  if (!codes) return set_error(LDW_ERR_ARG, "ldw_mi_plan_create: null argument");
  return ldw::mi_plan_create_impl(ctx, codes, nullptr, n_snp, nseq, hdw, pos, paint, blk, out);
}

}  // extern "C"

int ldw::mi_plan_create_impl(ldw_ctx* ctx, const uint8_t* codes, const uint8_t* codes_dev, int64_t n_snp, int64_t nseq,
                             const double* hdw, const int32_t* pos, const int32_t* paint, int64_t blk, ldw_mi_plan** out) {
  LDW_TRY(ctx_bind(ctx));
  if (!out) return set_error(LDW_ERR_ARG, "ldw_mi_plan_create: null out");
  *out = nullptr;
  if ((!codes && !codes_dev) || !hdw || !pos || !paint) return set_error(LDW_ERR_ARG, "ldw_mi_plan_create: null argument");
  if (n_snp < 2 || nseq < 1) return set_error(LDW_ERR_ARG, "ldw_mi_plan_create: need at least 2 SNPs and 1 sequence");
  if (nseq > 65535) return set_error(LDW_ERR_UNSUPPORTED, "ldw_mi_plan_create: nseq > 65535 not supported by the 15-bit digit accumulation");
  if (blk < 128 || blk > 65535) return set_error(LDW_ERR_UNSUPPORTED, "ldw_mi_plan_create: block size %lld outside [128, 65535]", (long long)blk);
  if (codes) {
    // every code must be 0..4; eight bytes per step (a byte > 4 sets bit 7 of byte + 0x7B, or has it set already),
    // chunks spread over a few host threads
    const int64_t total = n_snp * nseq;
    const int64_t chunk = (int64_t)1 << 22;
    const int64_t nchunks = (total + chunk - 1) / chunk;
    std::vector<int64_t> bad_at(nchunks, -1);
    parallel_for(nchunks, 8, [&](int64_t c) {
      int64_t i = c * chunk;
      const int64_t end = std::min<int64_t>(total, i + chunk);
      for (; i + 8 <= end; i += 8) {
        uint64_t x;
        memcpy(&x, codes + i, 8);
        if ((((x & 0x7F7F7F7F7F7F7F7Full) + 0x7B7B7B7B7B7B7B7Bull) | x) & 0x8080808080808080ull) break;
      }
      for (; i < end; i++)
        if (codes[i] > 4) { bad_at[c] = i; break; }
    });
    for (int64_t c = 0; c < nchunks; c++)
      if (bad_at[c] >= 0)
        return set_error(LDW_ERR_ARG, "ldw_mi_plan_create: codes[%lld] = %d outside 0..4", (long long)bad_at[c], (int)codes[bad_at[c]]);
  }
  ldw_mi_plan* P = new ldw_mi_plan();
  P->ctx = ctx;
  P->n = n_snp; P->S = nseq; P->blk = blk;
  int rc = build_plan(P, codes, codes_dev, hdw, pos, paint);
  if (rc != 0) { delete P; return rc; }
  *out = P;
  return 0;
}

extern "C" {

void ldw_mi_plan_destroy(ldw_mi_plan* plan) {
  if (!plan) return;
  cudaSetDevice(plan->ctx->device);
  delete plan;
}

int ldw_mi_block_dense(ldw_mi_plan* P, int64_t block_index, double* mi_out, int64_t* nf_out, int64_t* nt_out) {
  if (!P) return set_error(LDW_ERR_ARG, "null plan");
  LDW_TRY(ctx_bind(P->ctx));
  cudaStream_t st = P->ctx->stream;
  ScanWS* W = get_ws(P->ctx);
  // block_index -> (bf, bt) in make_blocks order
  int bf = 0, bt = 0;
  {
    int64_t k = block_index;
    bool found = false;
    for (int i = 0; i < P->nranges && !found; i++)
      for (int j = i; j < P->nranges; j++) {
        if (k == 0) { bf = i; bt = j; found = true; break; }
        k--;
      }
    if (!found) return set_error(LDW_ERR_ARG, "block index out of range");
  }
  ScanCfg cfg{1e15, 0.0, 0, 0, 0};
  cfg.dense = 1;
  BlockHost& H = W->hring[0];
  BlockDev& D = W->ring[0];
  int e = prepare_block(P, bf, bt, cfg, H);
  if (e > 1) return e;
  if (nf_out) *nf_out = H.nf;
  if (nt_out) *nt_out = H.nt;
  if (!mi_out) return 0;
  LDW_TRY(upload_block(st, D, H));
  size_t cells = (size_t)H.nf * H.nt;
  LDW_TRY(W->d_dense.ensure(cells * 4));
  LDW_CUDA(cudaMemsetAsync(W->d_dense.p, 0, cells * 4, st));
  ScanParams sp;
  fill_scan_params(P, D, H, cfg, sp);
  sp.dense = 1;
  sp.dense_out = W->d_dense.as<float>();
  LDW_TRY(launch_scan(P, sp, st));
  DevBuf d64;
  LDW_TRY(d64.alloc(cells * 8));
  f32_to_f64_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(W->d_dense.as<float>(), (int64_t)cells, d64.as<double>());
  LDW_CUDA(cudaMemcpyAsync(mi_out, d64.p, cells * 8, cudaMemcpyDeviceToHost, st));
  LDW_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int ldw_mi_pairs_exact(ldw_mi_plan* P, int64_t block_index, const int32_t* from_local, const int32_t* to_local,
                       int64_t n_pairs, double* mi_out) {
  if (!P) return set_error(LDW_ERR_ARG, "null plan");
  LDW_TRY(ctx_bind(P->ctx));
  cudaStream_t st = P->ctx->stream;
  ScanWS* W = get_ws(P->ctx);
  int bf = 0, bt = 0;
  {
    int64_t k = block_index;
    bool found = false;
    for (int i = 0; i < P->nranges && !found; i++)
      for (int j = i; j < P->nranges; j++) {
        if (k == 0) { bf = i; bt = j; found = true; break; }
        k--;
      }
    if (!found) return set_error(LDW_ERR_ARG, "block index out of range");
  }
  if (n_pairs <= 0) return 0;
  ScanCfg cfg{1e15, 0.0, 0, 0, 0};
  BlockHost& H = W->hring[0];
  BlockDev& D = W->ring[0];
  int e = prepare_block(P, bf, bt, cfg, H);
  if (e != 0) return e == 1 ? set_error(LDW_ERR_ARG, "empty block") : e;
  for (int64_t k = 0; k < n_pairs; k++)
    if (from_local[k] < 0 || from_local[k] >= H.nf || to_local[k] < 0 || to_local[k] >= H.nt)
      return set_error(LDW_ERR_ARG, "pair %lld outside the block", (long long)k);
  LDW_TRY(upload_block(st, D, H));
  DevBuf di, dj, dout;
  LDW_TRY(di.alloc((size_t)n_pairs * 4));
  LDW_TRY(dj.alloc((size_t)n_pairs * 4));
  LDW_TRY(dout.alloc((size_t)n_pairs * 8));
  LDW_CUDA(cudaMemcpyAsync(di.p, from_local, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, st));
  LDW_CUDA(cudaMemcpyAsync(dj.p, to_local, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, st));
  RefineParams R = make_refine_params(P, D, H, cfg);
  int grid = (int)std::min<int64_t>((n_pairs + REFINE_WARPS - 1) / REFINE_WARPS, 148 * 8);
  mi_refine_pairs_kernel<<<grid, 32 * REFINE_WARPS, 0, st>>>(R, di.as<int32_t>(), dj.as<int32_t>(), n_pairs, dout.as<double>());
  LDW_CUDA(cudaGetLastError());
  LDW_CUDA(cudaMemcpyAsync(mi_out, dout.p, (size_t)n_pairs * 8, cudaMemcpyDeviceToHost, st));
  LDW_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int ldw_mi_scan(ldw_mi_plan* P, double g, double sr_dist, double lr_retain_links, double lr_links_approx, int flags,
                int n_parts, int part, ldw_links* sr_out, ldw_links* lr_out, ldw_links* borderline_out, double* thr_out,
                double* prob_out, ldw_scan_stats* stats_out) {
  return ldw::mi_scan_impl(P, g, sr_dist, lr_retain_links, lr_links_approx, flags, n_parts, part, nullptr, sr_out, lr_out,
                           borderline_out, thr_out, prob_out, stats_out);
}

}  // extern "C"

// Sizes of every make_blocks block (pairs, short-range, long-range) on host threads: positions only, no device work.
int ldw::mi_block_sizes(const ldw_mi_plan* P, double g, double sr_dist, int flags, std::vector<BlockSizes>& out) {
  if (!P) return set_error(LDW_ERR_ARG, "null plan");
  ScanCfg cfg{g, sr_dist, 0, 0, flags};
  LDW_TRY(validate_scan(P, cfg));
  std::vector<std::pair<int, int>> bl;
  for (int i = 0; i < P->nranges; i++)
    for (int j = i; j < P->nranges; j++) bl.push_back({i, j});
  out.assign(bl.size(), BlockSizes{0, 0, 0, 0});
  parallel_for((int64_t)bl.size(), 16, [&](int64_t b) {
    BlockHost tmp;
    int e = prepare_block(P, bl[b].first, bl[b].second, cfg, tmp, true);
    out[b] = BlockSizes{tmp.n_pairs, tmp.n_sr, tmp.n_lr, e};
  });
  for (size_t b = 0; b < bl.size(); b++)
    if (out[b].err > 1) {  // re-run on this thread so that the (thread-local) error message is the caller's
      BlockHost t2;
      return prepare_block(P, bl[b].first, bl[b].second, cfg, t2, true);
    }
  return 0;
}

// Blocks are dealt by cost (pairs), largest first, each to the least-loaded part (lowest part on ties): diagonal blocks
// cost half of the others, so plain round-robin would leave up to 12 % imbalance at 8 parts.  Deterministic: every rank
// derives the same assignment (ldweaver_b200/api.py:partition_blocks mirrors it).
void ldw::mi_block_owners(const ldw_mi_plan* P, int n_parts, std::vector<int>& owner) {
  struct Item { int64_t cost, idx; };
  std::vector<Item> items;
  int64_t idx = 0;
  for (int i = 0; i < P->nranges; i++)
    for (int j = i; j < P->nranges; j++, idx++) {
      const int64_t ni = std::min<int64_t>(P->n, (int64_t)(i + 1) * P->blk) - (int64_t)i * P->blk;
      const int64_t nj = std::min<int64_t>(P->n, (int64_t)(j + 1) * P->blk) - (int64_t)j * P->blk;
      items.push_back({i == j ? ni * (ni - 1) / 2 : ni * nj, idx});
    }
  std::vector<Item> order = items;
  std::stable_sort(order.begin(), order.end(), [](const Item& a, const Item& b) { return a.cost > b.cost; });
  std::vector<int64_t> load(n_parts, 0);
  owner.assign(items.size(), 0);
  for (const Item& it : order) {
    int best = 0;
    for (int p = 1; p < n_parts; p++)
      if (load[p] < load[best]) best = p;
    owner[it.idx] = best;
    load[best] += it.cost;
  }
}

int64_t ldw::mi_plan_nblocks(const ldw_mi_plan* P) { return (int64_t)P->nranges * (P->nranges + 1) / 2; }

int ldw::mi_scan_impl(ldw_mi_plan* P, double g, double sr_dist, double lr_retain_links, double lr_links_approx, int flags,
                      int n_parts, int part, const ScanShared* shared, ldw_links* sr_out, ldw_links* lr_out,
                      ldw_links* borderline_out, double* thr_out, double* prob_out, ldw_scan_stats* stats_out) {
  if (!P) return set_error(LDW_ERR_ARG, "null plan");
  const auto t_entry = std::chrono::steady_clock::now();
  LDW_TRY(ctx_bind(P->ctx));
  cudaStream_t st = P->ctx->stream;
  ScanWS* W = get_ws(P->ctx);
  ScanCfg cfg{g, sr_dist, lr_retain_links, lr_links_approx, flags};
  LDW_TRY(validate_scan(P, cfg));
  const bool sr_only = (flags & LDW_SCAN_SR_ONLY) != 0;
  if (!sr_only && !(lr_links_approx > 0)) return set_error(LDW_ERR_ARG, "lr_links_approx must be > 0");
  if (n_parts < 1 || part < 0 || part >= n_parts) return set_error(LDW_ERR_ARG, "bad partition %d of %d", part, n_parts);

  // ---- block list (make_blocks order), this part's share.  Blocks are dealt by cost (pairs), largest first, each to
  //      the least-loaded part (lowest part on ties): diagonal blocks cost half of the others, so plain round-robin
  //      would leave up to 12 % imbalance at 8 parts.  Deterministic: every rank derives the same assignment.
  struct Blk { int bf, bt; int64_t index; };
  std::vector<Blk> blocks;
  {
    std::vector<int> owner;
    mi_block_owners(P, n_parts, owner);
    int64_t idx = 0;
    for (int i = 0; i < P->nranges; i++)
      for (int j = i; j < P->nranges; j++, idx++)
        if (owner[idx] == part) blocks.push_back({i, j, idx});
  }
  const int64_t nblk_total = (int64_t)P->nranges * (P->nranges + 1) / 2;
  if (thr_out) for (int64_t b = 0; b < nblk_total; b++) thr_out[b] = NAN;
  if (prob_out) for (int64_t b = 0; b < nblk_total; b++) prob_out[b] = NAN;

  cudaEvent_t ev0, ev1, ev2, ev3;
  LDW_CUDA(cudaEventCreate(&ev0)); LDW_CUDA(cudaEventCreate(&ev1)); LDW_CUDA(cudaEventCreate(&ev2)); LDW_CUDA(cudaEventCreate(&ev3));

  // ---- pass 1 (host): per-block sizes -> output offsets and long-range selection ranks
  struct Sel { int64_t n_lr = 0, n_sr = 0, n_pairs = 0; double prob = NAN; uint64_t k_lo = 0, k_hi = 0; double h = 0; int interp = 0;
               int emit_all = 0; uint32_t kprime = 0, delta = 1, cap = 0; int64_t sr_base = 0, sr_hbase = 0; int skip = 0; };
  std::vector<Sel> sel(blocks.size());
  int64_t total_sr = 0, total_pairs = 0, total_lr = 0;
  uint64_t max_cap = 1, sum_keep = 0;
  int64_t first_guess = -1;
  int first_rc = -1;
  {
    // block sizes on a few host threads (they gate the first launch), then the serial prefix / rank arithmetic
    std::vector<int> perr(blocks.size(), 0);
    // The block that will run first (the first one holding short-range links: diagonal or next to it) gets its full
    // tables built alongside, into ring slot 0, so the first launch does not wait for them.
    for (size_t b = 0; b < blocks.size() && first_guess < 0; b++)
      if (blocks[b].bt - blocks[b].bf <= 1) first_guess = (int64_t)b;
    if (first_guess < 0 && !blocks.empty()) first_guess = 0;
    parallel_for((int64_t)blocks.size() + 1, 8, [&](int64_t b) {
      if (b == (int64_t)blocks.size()) {
        if (first_guess >= 0) first_rc = prepare_block(P, blocks[first_guess].bf, blocks[first_guess].bt, cfg, W->hring[0]);
        return;
      }
      Sel& s = sel[b];
      if (shared) {  // the device group sized every block once for all its ranks
        const BlockSizes& z = shared->sizes[blocks[b].index];
        perr[b] = z.err;
        s.n_lr = z.n_lr; s.n_sr = z.n_sr; s.n_pairs = z.n_pairs;
        return;
      }
      BlockHost tmp;
      int e = prepare_block(P, blocks[b].bf, blocks[b].bt, cfg, tmp, true);
      perr[b] = e;
      s.n_lr = tmp.n_lr; s.n_sr = tmp.n_sr; s.n_pairs = tmp.n_pairs;
    });
    for (size_t b = 0; b < blocks.size(); b++) {
      const int e = perr[b];
      if (e > 1) {  // re-run on this thread so that the (thread-local) error message is the caller's
        BlockHost t2;
        return prepare_block(P, blocks[b].bf, blocks[b].bt, cfg, t2, true);
      }
      Sel& s = sel[b];
      s.skip = (e == 1);
      struct { int64_t n_lr, n_sr, n_pairs; } tmp{s.n_lr, s.n_sr, s.n_pairs};
      s.sr_base = total_sr;
      s.sr_hbase = shared ? shared->sr_hbase[blocks[b].index] : total_sr;  // row of the block in the host table
      total_sr += tmp.n_sr; total_pairs += tmp.n_pairs; total_lr += tmp.n_lr;
      if (!sr_only && tmp.n_lr > 0) {
        const double m = (double)tmp.n_lr;
        double prob = 1 - ((lr_retain_links * (m / lr_links_approx)) / m);  // R/computePairwiseMI.R:352
        if (prob < 0) prob = 0;
        s.prob = prob;
        double index = 1 + (m - 1 > 0 ? m - 1 : 0) * prob;  // stats::quantile type 7
        double lo = std::floor(index), hi = std::ceil(index);
        s.k_lo = (uint64_t)(tmp.n_lr - (int64_t)lo + 1);
        s.k_hi = (uint64_t)(tmp.n_lr - (int64_t)hi + 1);
        s.h = index - lo;
        s.interp = index > lo;
        uint64_t K = s.k_lo;
        if (2 * K + 65536 >= (uint64_t)tmp.n_lr) {
          s.emit_all = 1;
          if ((uint64_t)tmp.n_lr > 0xFFFFFFF0ull) return set_error(LDW_ERR_UNSUPPORTED, "block keeps more than 2^32 long-range links");
          s.cap = (uint32_t)tmp.n_lr;
          s.kprime = (uint32_t)std::min<uint64_t>(K, 0xFFFFFFFFull);
        } else {
          uint64_t kp = K + std::max<uint64_t>(1024, K / 8);
          uint64_t delta = std::max<uint64_t>(2 * kp, 32768);
          // room for the first wave of tiles that run before any threshold exists (all resident CTAs x one tile)
          uint64_t cap = std::min<uint64_t>((uint64_t)tmp.n_lr, 16 * kp + (4u << 20));
          if (cap > 0xFFFFFFF0ull) return set_error(LDW_ERR_UNSUPPORTED, "long-range candidate buffer exceeds 2^32 entries");
          s.kprime = (uint32_t)kp; s.delta = (uint32_t)delta; s.cap = (uint32_t)cap;
        }
        max_cap = std::max<uint64_t>(max_cap, s.cap);
        sum_keep += K;
      }
    }
  }
  const uint64_t kept_cap = 2 * sum_keep + (1u << 20);

  // ---- device workspace
  for (auto& l : W->lr) {
    LDW_TRY(l.cand.ensure(max_cap * sizeof(Cand)));
    LDW_TRY(l.mi64.ensure(max_cap * 8));
    LDW_TRY(l.vcand.ensure(max_cap * sizeof(Cand)));
    LDW_TRY(l.state.ensure(64));
    LDW_TRY(l.hist.ensure(MI_HIST_BINS * 4));
    l.used = false; l.chain_valid = false;
  }
  LDW_TRY(W->d_state.ensure(64));
  LDW_TRY(W->h_results.ensure(std::max<size_t>(blocks.size(), 1) * sizeof(BlockResult)));
  LDW_TRY(W->h_pub.ensure(64));
  memset(W->h_results.p, 0, std::max<size_t>(blocks.size(), 1) * sizeof(BlockResult));
  LDW_TRY(W->d_sr_f32.ensure((size_t)std::max<int64_t>(total_sr, 1) * 4));
  LDW_TRY(W->d_kept_key.ensure(kept_cap * 8));
  LDW_TRY(W->d_kept_gi.ensure(kept_cap * 4));
  LDW_TRY(W->d_kept_gj.ensure(kept_cap * 4));
  LDW_TRY(W->d_kept_mi.ensure(kept_cap * 8));
  LDW_TRY(W->d_kept_count.ensure(16));
  const bool sr_rows = !(flags & (LDW_SCAN_NO_LINKS | LDW_SCAN_LR_ONLY));  // short-range link columns wanted at all
  if (sr_rows) LDW_TRY(W->d_sr.ensure(total_sr));
  LDW_CUDA(cudaMemsetAsync(W->d_kept_count.p, 0, 16, st));
  uint32_t* d_kept_overflow = W->d_state.as<uint32_t>() + 3;
  // Threshold seed for the next blocks: ONE word, overwritten by every selection and read by every block's begin
  // kernel, whichever selection finished last (usually one or two blocks back).  The value read only steers how many
  // candidates get collected -- results do not depend on it (the selection verifies completeness and is exact) -- so
  // the unordered read across the two streams is harmless; zero (nothing selected yet) means "collect from zero".
  uint32_t* d_chain2 = W->d_state.as<uint32_t>() + 8;
  LDW_CUDA(cudaMemsetAsync(W->d_state.p, 0, 64, st));
  cudaStream_t sst = P->ctx->select_stream;

  const bool want_host = !(flags & (LDW_SCAN_NO_LINKS | LDW_SCAN_NO_D2H));
  const bool sr_to_host = want_host && !(flags & LDW_SCAN_SR_ON_DEVICE);  // short-range rows cross PCIe at all?
  P->ctx->dev_sr.n = -1;
  cudaStream_t cst = P->ctx->copy_stream;
  cudaEvent_t ev_blk = nullptr;  // "this block's SR columns are materialised"
  // short-range rows go to this context's pinned table, or -- multi-GPU -- straight to the group's table, every rank
  // at the final offsets of its blocks (the table is complete when the last rank finishes: no merge pass)
  HostLinks& hsr = shared ? *shared->h_sr : P->ctx->h_sr;
  if (sr_to_host) {
    if (!shared && sr_rows) LDW_TRY(hsr.ensure(total_sr));
    LDW_CUDA(cudaEventCreateWithFlags(&ev_blk, cudaEventDisableTiming));
  }
  auto tpre = std::chrono::steady_clock::now();
  LDW_CUDA(cudaEventRecord(ev0, st));
  int64_t n_reruns = 0, n_launches = 0, n_scan_launches = 0, n_tiles = 0;
  double exec_ops = 0, exec_mufu = 0;
  std::vector<cudaEvent_t> kev;  // pairs of events around every scan-kernel launch
  auto kev_cleanup = [&]() { for (auto e : kev) cudaEventDestroy(e); kev.clear(); };
  int64_t dbg_block = -1;
  {
    const char* e = getenv("LDW_DBG_BLOCK");
    if (e) dbg_block = atoll(e);
  }
  double host_prep_ms = 0;
  int scan_reserve = 2;  // SMs left to the selection stream (presel, refine, select of earlier blocks); 1-3 within 1 % at C2
  if (const char* e = getenv("LDW_SCAN_RESERVE")) scan_reserve = atoi(e);
  // test hook: force every block's first attempt to start from this candidate threshold (a value above the true
  // threshold makes the selection fail its completeness check and exercises the re-run path)
  float force_seed = -1.f;
  if (const char* e = getenv("LDW_DBG_FORCE_SEED")) force_seed = (float)atof(e);
  // b: index into this rank's block list (output offsets, results); seq: position in the execution order (ring slots)
  // margin of the long-range selection (pre-selection window and completeness test): sound while twice the fp32
  // epilogue's error stays below it; the selection kernel measures that error per block and asks for a re-run with a
  // wider margin when it does not (BlockResult.bad & 8)
  double sel_margin0 = 4e-6;
  if (const char* e = getenv("LDW_DBG_SEL_MARGIN")) sel_margin0 = atof(e);  // test hook: a tiny margin exercises that re-run
  auto run_block = [&](size_t b, size_t seq, int force_emit_all, uint32_t cap_override, bool use_chain, double sel_margin) -> int {
    Sel& s = sel[b];
    if (s.skip) return 0;
    int slot = (int)(seq % ScanWS::RING);
    BlockDev& D = W->ring[slot];
    BlockHost& H = W->hring[slot];
    // host staging + device arrays of this ring entry are free again
    if (D.used) LDW_CUDA(cudaEventSynchronize(D.done));
    if (D.used2) { LDW_CUDA(cudaEventSynchronize(D.done2)); D.used2 = false; }
    ScanWS::LrBuf& L = W->lr[seq % ScanWS::NLR];
    uint32_t* d_count = L.state.as<uint32_t>();
    uint32_t* d_tcand = d_count + 1;
    uint32_t* d_overflow = d_count + 2;
    uint32_t* d_chain = d_chain2;
    auto hp0 = std::chrono::steady_clock::now();
    int e;
    if (seq == 0 && first_guess == (int64_t)b && first_rc >= 0 && first_rc <= 1 && !force_emit_all && !cap_override) {
      e = first_rc;  // built during pass 1
      first_rc = -1;
    } else {
      e = prepare_block(P, blocks[b].bf, blocks[b].bt, cfg, H);
    }
    if (e > 1) return e;
    // the tables travel on the upload stream while earlier blocks are still being scanned
    LDW_TRY(upload_block(P->ctx->upload_stream, D, H));
    LDW_CUDA(cudaEventRecord(D.uploaded, P->ctx->upload_stream));
    LDW_CUDA(cudaStreamWaitEvent(st, D.uploaded, 0));
    host_prep_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - hp0).count();
    D.used = true;
    ScanParams sp;
    fill_scan_params(P, D, H, cfg, sp);
    sp.sr_out = W->d_sr_f32.as<float>() + s.sr_base;
    if (dbg_block >= 0 && (int64_t)b == dbg_block) {
      LDW_TRY(W->d_dbg.ensure(4096 * 16 * 8));
      LDW_CUDA(cudaMemsetAsync(W->d_dbg.p, 0, 4096 * 16 * 8, st));
      sp.dbg = W->d_dbg.as<unsigned long long>();
    }
    const bool lr = !sr_only && s.n_lr > 0;
    const int emit_all = force_emit_all || s.emit_all;
    uint32_t cap = cap_override ? cap_override : s.cap;
    if (lr) {
      // this buffer's previous selection (three blocks earlier) must be done before its counters are cleared
      if (L.used) LDW_CUDA(cudaStreamWaitEvent(st, L.sel_done, 0));
      const int n_pilot = (int)H.tiles.size() - H.n_real_tiles;
      if (seq == 0 && use_chain && !emit_all && n_pilot > 0) {
        // first block of the call: nothing has been selected yet, so estimate its threshold from a pilot sample
        mi_block_begin_kernel<<<1, 256, 0, st>>>(d_count, L.hist.as<uint32_t>(), d_chain, 0);
        ScanParams pp = sp;
        pp.tiles = D.tiles.as<TileDesc>() + H.n_real_tiles;
        pp.n_tiles = n_pilot;
        pp.cand = L.cand.as<Cand>(); pp.cand_cap = cap; pp.cand_count = d_count; pp.tcand_bits = d_tcand;
        pp.hist = L.hist.as<uint32_t>(); pp.kprime = s.kprime; pp.delta = s.delta ? s.delta : 1; pp.overflow = d_overflow;
        pp.emit_all = 1;
        pp.dbg = nullptr;
        LDW_TRY(launch_scan(P, pp, st, 0));
        mi_pilot_seed_kernel<<<1, 1024, 0, st>>>(L.cand.as<Cand>(), d_count, cap, (unsigned long long)s.k_lo, (double)s.n_lr, d_chain);
        LDW_CUDA(cudaGetLastError());
        n_launches += 3;
      }
      if (force_seed >= 0.f && use_chain && !emit_all) {
        static_assert(sizeof(float) == 4, "float bits");
        LDW_CUDA(cudaMemcpyAsync(d_chain, &force_seed, 4, cudaMemcpyHostToDevice, st));
        LDW_CUDA(cudaStreamSynchronize(st));
      }
      mi_block_begin_kernel<<<1, 256, 0, st>>>(d_count, L.hist.as<uint32_t>(), d_chain, (use_chain && !emit_all) ? 1 : 0);
      LDW_CUDA(cudaGetLastError());
      n_launches++;
      sp.cand = L.cand.as<Cand>();
      sp.cand_cap = cap;
      sp.cand_count = d_count;
      sp.tcand_bits = d_tcand;
      sp.hist = L.hist.as<uint32_t>();
      sp.kprime = s.kprime;
      sp.delta = s.delta ? s.delta : 1;
      sp.overflow = d_overflow;
      sp.emit_all = emit_all;
    } else {
      sp.sr_only = 1;  // nothing long-range to collect in this block
    }
    if (sp.n_tiles > 0) {
      cudaEvent_t k0, k1;
      LDW_CUDA(cudaEventCreate(&k0));
      LDW_CUDA(cudaEventCreate(&k1));
      kev.push_back(k0);
      kev.push_back(k1);
      LDW_CUDA(cudaEventRecord(k0, st));
      LDW_TRY(launch_scan(P, sp, st, lr ? scan_reserve : 0));
      LDW_CUDA(cudaEventRecord(k1, st));
      n_launches++; n_scan_launches++;
      n_tiles += sp.n_tiles;
      for (int32_t ti = 0; ti < H.n_real_tiles; ti++) {  // 4 K-passes x 2 ops/MAC x 128 rows x (PA*PB*NJ) columns x Kpad
        const TileDesc& td = H.tiles[ti];
        exec_ops += 8.0 * 256.0 * (double)(td.PA * td.PB * (1 << td.njlog2)) * (double)P->Kpad;  // 2 passes x 2 halves, M = 256
        // MUFU instructions of the epilogue, per lane-pair of the tile (all 128 x NJ lanes of a non-null half execute): one
        // LG2 per cell of the (PA+1) x (PB+1) table, plus -- Q1 form, off-diagonal blocks -- one RCP per two cells of a row
        const int RA = td.PA + 1, RB = td.PB + 1;
        const int halves = ((td.flags & TILE_NULL) ? 0 : 1) + ((td.flags1 & TILE_NULL) ? 0 : 1);
        exec_mufu += halves * 128.0 * (double)(1 << td.njlog2) * (double)(RA * (RB + (sp.qcorr ? (RB + 1) / 2 : 0)));
      }
    }
    if (lr) {
      // fp64 refinement + exact selection on the select stream, overlapping the next block's scan
      n_launches += 2;
      LDW_CUDA(cudaEventRecord(L.scan_done, st));
      LDW_CUDA(cudaStreamWaitEvent(sst, L.scan_done, 0));
      RefineParams R = make_refine_params(P, D, H, cfg);
      // candidates that can reach the exact K-th largest value (fp32 pre-selection), then their fp64 refinement
      n_launches++;
      mi_presel_kernel<<<1, 1024, 0, sst>>>(L.cand.as<Cand>(), d_count, cap, d_tcand, emit_all, (unsigned long long)s.k_lo, (float)sel_margin,
                                           L.vcand.as<Cand>(), d_count + 5);
      LDW_CUDA(cudaGetLastError());
      mi_refine_list_kernel<<<P->ctx->num_sms * 4, 32 * REFINE_WARPS, 0, sst>>>(R, L.vcand.as<Cand>(), d_count + 5, L.mi64.as<double>());
      LDW_CUDA(cudaGetLastError());
      SelectParams q;
      memset(&q, 0, sizeof(q));
      q.cand = L.vcand.as<Cand>(); q.mi64 = L.mi64.as<double>(); q.vcount = d_count + 5; q.count = d_count; q.cap = cap;
      q.overflow = d_overflow; q.tcand_bits = d_tcand; q.emit_all = emit_all;
      q.k_lo = s.k_lo; q.k_hi = s.k_hi; q.h = s.h; q.interpolate = s.interp;
      q.tol_safe = sel_margin; q.tol_border = 1e-9;
      q.from_idx = D.from_idx.as<int32_t>(); q.to_idx = D.to_idx.as<int32_t>();
      q.nf = H.nf; q.nt = H.nt; q.diag = H.diag; q.block = (int32_t)blocks[b].index;
      q.kept_key = W->d_kept_key.as<uint64_t>(); q.kept_gi = W->d_kept_gi.as<int32_t>(); q.kept_gj = W->d_kept_gj.as<int32_t>();
      q.kept_mi = W->d_kept_mi.as<double>(); q.kept_count = W->d_kept_count.as<unsigned long long>(); q.kept_cap = kept_cap;
      q.kept_overflow = d_kept_overflow;
      q.chain_bits = d_chain;
      q.result = W->h_results.as<BlockResult>() + b;
      mi_select_kernel<<<1, 1024, 0, sst>>>(q);
      LDW_CUDA(cudaGetLastError());
      LDW_CUDA(cudaEventRecord(L.sel_done, sst));
      LDW_CUDA(cudaEventRecord(D.done2, sst));
      L.used = true;
      D.used2 = true;
    }
    if (s.n_sr > 0 && sr_rows) {
      n_launches++;
      SrMatParams m;
      m.col = D.colinfo.as<ColInfo>(); m.from_idx = D.from_idx.as<int32_t>(); m.to_idx = D.to_idx.as<int32_t>();
      m.pos = P->d_pos.as<int32_t>(); m.paint = P->d_paint.as<int32_t>();
      m.sr_mi = W->d_sr_f32.as<float>() + s.sr_base;
      m.nf = H.nf; m.nt = H.nt; m.diag = H.diag; m.block = (int32_t)blocks[b].index; m.g = (int64_t)g;
      m.o_pos1 = W->d_sr.pos1.as<int32_t>() + s.sr_base; m.o_pos2 = W->d_sr.pos2.as<int32_t>() + s.sr_base;
      m.o_c1 = W->d_sr.c1.as<int32_t>() + s.sr_base; m.o_c2 = W->d_sr.c2.as<int32_t>() + s.sr_base;
      m.o_len = W->d_sr.len.as<int32_t>() + s.sr_base; m.o_blk = W->d_sr.blk.as<int32_t>() + s.sr_base;
      m.o_mi = W->d_sr.mi.as<double>() + s.sr_base;
      mi_sr_materialize_kernel<<<H.nt, 128, 0, st>>>(m);
      LDW_CUDA(cudaGetLastError());
      if (flags & LDW_SCAN_SR_EXACT) {  // fp64 MI over the fp32-derived column, before the rows leave for the host
        n_launches++;
        mi_sr_exact_kernel<<<H.nt, 32 * REFINE_WARPS, 0, st>>>(m, make_refine_params(P, D, H, cfg));
        LDW_CUDA(cudaGetLastError());
      }
      if (sr_to_host) {  // (sr_rows holds here)
        // copy this block's finished rows to the host while the next blocks are being scanned
        LDW_CUDA(cudaEventRecord(ev_blk, st));
        LDW_CUDA(cudaStreamWaitEvent(cst, ev_blk, 0));
        HostLinks& h = hsr;
        const size_t o = (size_t)s.sr_hbase, nb4 = (size_t)s.n_sr * 4, nb8 = (size_t)s.n_sr * 8;
        LDW_CUDA(cudaMemcpyAsync(h.pos1.as<int32_t>() + o, m.o_pos1, nb4, cudaMemcpyDeviceToHost, cst));
        LDW_CUDA(cudaMemcpyAsync(h.pos2.as<int32_t>() + o, m.o_pos2, nb4, cudaMemcpyDeviceToHost, cst));
        LDW_CUDA(cudaMemcpyAsync(h.c1.as<int32_t>() + o, m.o_c1, nb4, cudaMemcpyDeviceToHost, cst));
        LDW_CUDA(cudaMemcpyAsync(h.c2.as<int32_t>() + o, m.o_c2, nb4, cudaMemcpyDeviceToHost, cst));
        LDW_CUDA(cudaMemcpyAsync(h.len.as<int32_t>() + o, m.o_len, nb4, cudaMemcpyDeviceToHost, cst));
        LDW_CUDA(cudaMemcpyAsync(h.blk.as<int32_t>() + o, m.o_blk, nb4, cudaMemcpyDeviceToHost, cst));
        LDW_CUDA(cudaMemcpyAsync(h.mi.as<double>() + o, m.o_mi, nb8, cudaMemcpyDeviceToHost, cst));
      }
    }
    LDW_CUDA(cudaEventRecord(D.done, st));
    return 0;
  };
  // Execution order: blocks that hold short-range links first, so their (large) link columns cross PCIe while the
  // long-range-only blocks are still being scanned.  Output order is unaffected: short-range rows have fixed
  // offsets and the kept long-range rows are sorted by (block, row rank) afterwards.
  std::vector<size_t> order;
  for (size_t b = 0; b < blocks.size(); b++) if (sel[b].n_sr > 0) order.push_back(b);
  for (size_t b = 0; b < blocks.size(); b++) if (!(sel[b].n_sr > 0)) order.push_back(b);
  for (size_t k = 0; k < order.size(); k++) LDW_TRY(run_block(order[k], k, 0, 0, true, sel_margin0));
  for (auto& l : W->lr)
    if (l.used) LDW_CUDA(cudaStreamWaitEvent(st, l.sel_done, 0));  // join the select stream
  LDW_CUDA(cudaEventRecord(ev1, st));
  const bool dbg_timing = getenv("LDW_DBG_TIMING") != nullptr;
  auto tp0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!dbg_timing) return;
    cudaStreamSynchronize(st);
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "ldw timing: %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - tp0).count());
    tp0 = t;
  };
  if (dbg_timing) fprintf(stderr, "ldw timing: %-28s %8.3f ms\n", "sizes + workspace (host)", std::chrono::duration<double, std::milli>(tpre - t_entry).count());
  if (dbg_timing) fprintf(stderr, "ldw timing: %-28s %8.3f ms\n", "submit blocks (host)", std::chrono::duration<double, std::milli>(tp0 - tpre).count());
  lap("drain blocks");

  // ---- verify the long-range selection of every block; re-run the (rare) blocks whose candidate set is not
  //      provably complete with a plain "collect everything" pass
  LDW_CUDA(cudaStreamSynchronize(st));
  P->results.assign(W->h_results.as<BlockResult>(), W->h_results.as<BlockResult>() + blocks.size());
  if (!sr_only) {
    for (size_t b = 0; b < blocks.size(); b++) {
      if (sel[b].skip || sel[b].n_lr == 0) continue;
      if (P->results[b].bad) {
        // second attempt: no chained seed (collect from zero, the histogram places the threshold); last resort:
        // collect every long-range pair of the block
        double margin = sel_margin0;
        for (int attempt = 0; attempt < 2 && P->results[b].bad; attempt++) {
          n_reruns++;
          // measured epilogue error too large for the margin: widen it to 8x the observation (the test wants 4x)
          if (P->results[b].bad & 8) margin = std::max(std::max(margin, 4e-6), 8.0 * (double)P->results[b].eps_obs);
          uint32_t cap = 0;
          if (attempt == 1) {
            if ((uint64_t)sel[b].n_lr > 0xFFFFFFF0ull) return set_error(LDW_ERR_UNSUPPORTED, "block %lld needs an exhaustive long-range pass over more than 2^32 links", (long long)blocks[b].index);
            cap = (uint32_t)sel[b].n_lr;
            ScanWS::LrBuf& L = W->lr[0];
            LDW_TRY(L.cand.ensure((size_t)cap * sizeof(Cand)));
            LDW_TRY(L.mi64.ensure((size_t)cap * 8));
            LDW_TRY(L.vcand.ensure((size_t)cap * sizeof(Cand)));
          }
          LDW_TRY(run_block(b, 0, attempt == 1, cap, false, margin));
          LDW_CUDA(cudaStreamWaitEvent(st, W->lr[0].sel_done, 0));
          LDW_CUDA(cudaStreamSynchronize(st));
          P->results[b] = W->h_results.as<BlockResult>()[b];
        }
        if (P->results[b].bad) return set_error(LDW_ERR_INTERNAL, "long-range selection failed for block %lld (flags %u)", (long long)blocks[b].index, P->results[b].bad);
      }
    }
  }
  lap("results + reruns");
  unsigned long long n_kept = 0;
  uint32_t state[4] = {0, 0, 0, 0};
  publish_kernel<<<1, 1, 0, st>>>(W->d_kept_count.as<unsigned long long>(), W->d_state.as<uint32_t>(), W->h_pub.as<unsigned long long>());
  LDW_CUDA(cudaGetLastError());
  LDW_CUDA(cudaStreamSynchronize(st));
  n_kept = W->h_pub.as<unsigned long long>()[0];
  state[3] = (uint32_t)W->h_pub.as<unsigned long long>()[1];
  if (state[3] || n_kept > kept_cap) return set_error(LDW_ERR_UNSUPPORTED, "more long-range links pass their block thresholds (%llu) than the output buffer holds (%llu): massive ties at the threshold", n_kept, (unsigned long long)kept_cap);

  // ---- long-range rows into reference order (block, then row order inside the block)
  int64_t n_border = 0;
  for (auto& r : P->results) n_border += r.n_border;
  if (n_kept > 0 && !(flags & LDW_SCAN_NO_LINKS)) {
    n_launches += 2 + 4;  // iota, materialise, radix-sort passes (library)
    LDW_TRY(W->d_keys_sorted.ensure(n_kept * 8));
    LDW_TRY(W->d_order_in.ensure(n_kept * 4));
    LDW_TRY(W->d_order_out.ensure(n_kept * 4));
    iota_u32_kernel<<<(unsigned)((n_kept + 255) / 256), 256, 0, st>>>(W->d_order_in.as<uint32_t>(), (int64_t)n_kept);
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, W->d_kept_key.as<uint64_t>(), W->d_keys_sorted.as<uint64_t>(),
                                    W->d_order_in.as<uint32_t>(), W->d_order_out.as<uint32_t>(), (int)n_kept, 0, 64, st);
    LDW_TRY(W->d_sort_tmp.ensure(tmp_bytes));
    LDW_CUDA(cub::DeviceRadixSort::SortPairs(W->d_sort_tmp.p, tmp_bytes, W->d_kept_key.as<uint64_t>(), W->d_keys_sorted.as<uint64_t>(),
                                             W->d_order_in.as<uint32_t>(), W->d_order_out.as<uint32_t>(), (int)n_kept, 0, 64, st));
    LDW_TRY(W->d_lr.ensure((int64_t)n_kept));
    mi_lr_materialize_kernel<<<(unsigned)((n_kept + 255) / 256), 256, 0, st>>>(
        W->d_order_out.as<uint32_t>(), W->d_keys_sorted.as<uint64_t>(), W->d_kept_gi.as<int32_t>(), W->d_kept_gj.as<int32_t>(),
        W->d_kept_mi.as<double>(), P->d_pos.as<int32_t>(), P->d_paint.as<int32_t>(), (int64_t)n_kept, (int64_t)g,
        W->d_lr.pos1.as<int32_t>(), W->d_lr.pos2.as<int32_t>(), W->d_lr.c1.as<int32_t>(), W->d_lr.c2.as<int32_t>(),
        W->d_lr.len.as<int32_t>(), W->d_lr.mi.as<double>(), W->d_lr.blk.as<int32_t>());
    LDW_CUDA(cudaGetLastError());
  }
  LDW_CUDA(cudaEventRecord(ev2, st));
  lap("lr sort + materialise");

  // ---- device -> pinned host
  P->ctx->h_sr.n = 0; P->ctx->h_lr.n = 0; P->ctx->h_border.n = 0;
  if (!(flags & (LDW_SCAN_NO_LINKS | LDW_SCAN_NO_D2H))) {
    auto d2h = [&](HostLinks& h, DevLinks& d, int64_t m) -> int {
      LDW_TRY(h.ensure(m));
      h.n = m;
      if (m == 0) return 0;
      LDW_CUDA(cudaMemcpyAsync(h.pos1.p, d.pos1.p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
      LDW_CUDA(cudaMemcpyAsync(h.pos2.p, d.pos2.p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
      LDW_CUDA(cudaMemcpyAsync(h.c1.p, d.c1.p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
      LDW_CUDA(cudaMemcpyAsync(h.c2.p, d.c2.p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
      LDW_CUDA(cudaMemcpyAsync(h.len.p, d.len.p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
      LDW_CUDA(cudaMemcpyAsync(h.blk.p, d.blk.p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
      LDW_CUDA(cudaMemcpyAsync(h.mi.p, d.mi.p, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
      return 0;
    };
    if (!shared) P->ctx->h_sr.n = (sr_rows && sr_to_host) ? total_sr : 0;  // short-range rows were streamed out block by block on the copy stream
    LDW_TRY(d2h(P->ctx->h_lr, W->d_lr, (int64_t)n_kept));
  }
  LDW_CUDA(cudaEventRecord(ev3, st));
  LDW_CUDA(cudaStreamSynchronize(st));
  lap("lr d2h");
  if (sr_to_host) LDW_CUDA(cudaStreamSynchronize(cst));
  lap("copy stream drain");
  if (ev_blk) cudaEventDestroy(ev_blk);

  // borderline list: kept-or-not candidates within tol of their block threshold are reported by count per block;
  // the explicit rows are the LR rows whose MI is within tol of the block threshold
  if (borderline_out) {
    std::vector<int64_t> idx;
    const int32_t* blkcol = P->ctx->h_lr.blk.as<int32_t>();
    const double* micol = P->ctx->h_lr.mi.as<double>();
    std::vector<double> thr_by_index(nblk_total, NAN);
    for (size_t b = 0; b < blocks.size(); b++)
      if (!sel[b].skip && sel[b].n_lr > 0 && !sr_only) thr_by_index[blocks[b].index] = P->results[b].thr;
    for (int64_t i = 0; i < P->ctx->h_lr.n; i++)
      if (std::fabs(micol[i] - thr_by_index[blkcol[i]]) <= 1e-9) idx.push_back(i);
    LDW_TRY(P->ctx->h_border.ensure((int64_t)idx.size()));
    P->ctx->h_border.n = (int64_t)idx.size();
    for (size_t k = 0; k < idx.size(); k++) {
      int64_t i = idx[k];
      P->ctx->h_border.pos1.as<int32_t>()[k] = P->ctx->h_lr.pos1.as<int32_t>()[i];
      P->ctx->h_border.pos2.as<int32_t>()[k] = P->ctx->h_lr.pos2.as<int32_t>()[i];
      P->ctx->h_border.c1.as<int32_t>()[k] = P->ctx->h_lr.c1.as<int32_t>()[i];
      P->ctx->h_border.c2.as<int32_t>()[k] = P->ctx->h_lr.c2.as<int32_t>()[i];
      P->ctx->h_border.len.as<int32_t>()[k] = P->ctx->h_lr.len.as<int32_t>()[i];
      P->ctx->h_border.blk.as<int32_t>()[k] = blkcol[i];
      P->ctx->h_border.mi.as<double>()[k] = micol[i];
    }
    P->ctx->h_border.fill(borderline_out);
  }
  if (flags & (LDW_SCAN_NO_LINKS | LDW_SCAN_NO_D2H)) {
    if (sr_out) { memset(sr_out, 0, sizeof(*sr_out)); sr_out->n = total_sr; }
    if (lr_out) { memset(lr_out, 0, sizeof(*lr_out)); lr_out->n = (int64_t)n_kept; }
  } else {
    if (shared || !sr_to_host) { if (sr_out) { memset(sr_out, 0, sizeof(*sr_out)); sr_out->n = sr_rows ? total_sr : 0; } }
    else P->ctx->h_sr.fill(sr_out);

    P->ctx->h_lr.fill(lr_out);
  }
  if (sr_rows && n_parts == 1) {  // the whole job's table is in this context's device memory, in reference order
    ldw_ctx::DevSr& d = P->ctx->dev_sr;
    d.pos1 = W->d_sr.pos1.as<int32_t>(); d.pos2 = W->d_sr.pos2.as<int32_t>(); d.c1 = W->d_sr.c1.as<int32_t>();
    d.c2 = W->d_sr.c2.as<int32_t>(); d.len = W->d_sr.len.as<int32_t>(); d.blk = W->d_sr.blk.as<int32_t>();
    d.mi = W->d_sr.mi.as<double>();
    d.n = total_sr;
  }
  for (size_t b = 0; b < blocks.size(); b++) {
    if (sel[b].skip || sel[b].n_lr == 0 || sr_only) continue;
    if (thr_out) thr_out[blocks[b].index] = P->results[b].thr;
    if (prob_out) prob_out[blocks[b].index] = sel[b].prob;
  }
  if (stats_out) {
    memset(stats_out, 0, sizeof(*stats_out));
    stats_out->n_blocks = (int64_t)blocks.size();
    stats_out->n_pairs = total_pairs;
    stats_out->n_sr = total_sr;
    stats_out->n_lr_total = sr_only ? 0 : total_lr;
    stats_out->n_lr_kept = (int64_t)n_kept;
    stats_out->n_borderline = n_border;
    {
      int64_t nc = 0;
      for (auto& r : P->results) nc += r.n_cand;
      stats_out->n_candidates = nc;
    }
    stats_out->n_reruns = n_reruns;
    float a = 0, b = 0, c = 0;
    cudaEventElapsedTime(&a, ev0, ev1);
    cudaEventElapsedTime(&b, ev1, ev2);
    cudaEventElapsedTime(&c, ev2, ev3);
    stats_out->t_pack_ms = P->t_pack_ms;
    stats_out->t_scan_ms = a;
    stats_out->t_select_ms = b;
    stats_out->t_d2h_ms = c;
    double tk = 0;
    for (size_t k = 0; k + 1 < kev.size(); k += 2) {
      float ms = 0;
      cudaEventElapsedTime(&ms, kev[k], kev[k + 1]);
      tk += ms;
    }
    stats_out->t_kernel_ms = tk;
    if (dbg_timing && kev.size() >= 4) {
      double gaps = 0, gmax = 0;
      std::string detail;
      for (size_t k = 1; k + 1 < kev.size(); k += 2) {
        float ms = 0;
        cudaEventElapsedTime(&ms, kev[k], kev[k + 1]);
        gaps += ms;
        gmax = std::max<double>(gmax, ms);
        char b[32];
        snprintf(b, sizeof(b), " %.0f", 1e3 * ms);
        detail += b;
      }
      float first = 0, last = 0;
      cudaEventElapsedTime(&first, ev0, kev[0]);
      cudaEventElapsedTime(&last, kev[kev.size() - 1], ev1);
      fprintf(stderr, "ldw timing: gaps between scan kernels: sum %.3f ms, max %.3f ms; before first %.3f ms, after last %.3f ms; us:%s\n",
              gaps, gmax, first, last, detail.c_str());
    }
    stats_out->n_scan_launches = n_scan_launches;
    stats_out->n_launches = n_launches;
    stats_out->n_tiles = n_tiles;
    stats_out->exec_int8_ops = exec_ops;
    stats_out->t_host_prep_ms = host_prep_ms;
    stats_out->exec_mufu_ops = exec_mufu;
    {
      double em = 0;
      for (size_t b = 0; b < blocks.size(); b++)
        if (!sel[b].skip && sel[b].n_lr > 0 && !sr_only) em = std::max(em, (double)P->results[b].eps_obs);
      stats_out->eps_obs_max = em;
    }
  }
  if (dbg_block >= 0 && W->d_dbg.p) {
    std::vector<unsigned long long> h(4096 * 16);
    cudaMemcpy(h.data(), W->d_dbg.p, h.size() * 8, cudaMemcpyDeviceToHost);
    double s[16] = {0};
    int nb = 0;
    for (int c = 0; c < 4096; c++)
      if (h[c * 16 + 3]) { nb++; for (int k = 0; k < 16; k++) s[k] += (double)h[c * 16 + k]; }
    if (nb) {
      fprintf(stderr, "ldw dbg block %lld (%d CTAs, kilo-cycles avg): producer total %.0f wait_jempty %.0f wait_empty %.0f | "
              "mma total %.0f wait_tempty %.0f wait_ready %.0f | expander total %.0f wait_full %.0f | epi(first) total %.0f "
              "wait_jfull %.0f wait_tfull %.0f | epi(last) total %.0f wait_jfull %.0f wait_tfull %.0f\n", (long long)dbg_block, nb,
              s[0] / nb / 1e3, s[1] / nb / 1e3, s[2] / nb / 1e3, s[3] / nb / 1e3, s[4] / nb / 1e3, s[5] / nb / 1e3, s[12] / nb / 1e3,
              s[13] / nb / 1e3, s[6] / nb / 1e3, s[7] / nb / 1e3, s[8] / nb / 1e3, s[9] / nb / 1e3, s[10] / nb / 1e3, s[11] / nb / 1e3);
    }
  }
  kev_cleanup();
  cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(ev2); cudaEventDestroy(ev3);
  lap("borderline + stats (host)");
  return 0;
}
