// The context behind the opaque ldw_ctx handle.
#pragma once
#include "host_util.h"

// Link columns in pinned host memory (what ldw_links points into).  Owned by the context and reused by every
// scan on it: allocating gigabytes of pinned memory costs far more than the scan itself.
struct HostLinks {
  ldw::PinnedBuf pos1, pos2, c1, c2, len, blk, mi;
  int64_t n = 0;
  int ensure(int64_t m) {
    size_t b4 = (size_t)(m > 0 ? m : 1) * 4, b8 = (size_t)(m > 0 ? m : 1) * 8;
    LDW_TRY(pos1.ensure(b4)); LDW_TRY(pos2.ensure(b4)); LDW_TRY(c1.ensure(b4)); LDW_TRY(c2.ensure(b4));
    LDW_TRY(len.ensure(b4)); LDW_TRY(blk.ensure(b4)); LDW_TRY(mi.ensure(b8));
    return 0;
  }
  void fill(ldw_links* o) const {
    if (!o) return;
    o->n = n;
    o->pos1 = pos1.as<int32_t>(); o->pos2 = pos2.as<int32_t>(); o->clust1 = c1.as<int32_t>(); o->clust2 = c2.as<int32_t>();
    o->len = len.as<int32_t>(); o->MI = mi.as<double>(); o->block = blk.as<int32_t>();
  }
};

struct ldw_ctx {
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // device -> host copies of finished link columns overlap the scan
  cudaStream_t upload_stream = nullptr;  // per-block tables host -> device, ahead of the block's scan
  cudaStream_t select_stream = nullptr;  // fp64 refinement + long-range selection of block b overlap the scan of block b+1
  HostLinks h_sr, h_lr, h_border;
  // scan workspace (device scratch, per-block ring, link columns): allocated on first use by mi_scan.cu, reused by
  // every plan on this context, released by ldw_destroy
  void* scan_ws = nullptr;
  void (*scan_ws_free)(void*) = nullptr;
  // The short-range table the last scan on this context left in device memory (columns inside scan_ws; valid until the
  // next scan): what ldw_sr_postprocess_dev works on.  n < 0: none (no scan yet, a partial scan, or no rows materialised).
  struct DevSr {
    const int32_t *pos1 = nullptr, *pos2 = nullptr, *c1 = nullptr, *c2 = nullptr, *len = nullptr, *blk = nullptr;
    const double* mi = nullptr;
    int64_t n = -1;
  } dev_sr;
};

namespace ldw {
// Binds the calling thread to the context's device.
int ctx_bind(ldw_ctx* ctx);
}  // namespace ldw
