// The context behind the opaque ldw_ctx handle.
#pragma once
#include "host_util.h"

struct ldw_ctx {
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
};

namespace ldw {
// Binds the calling thread to the context's device.
int ctx_bind(ldw_ctx* ctx);
}  // namespace ldw
