#include "host_util.h"

#include <map>
#include <mutex>

namespace ldw {

std::string& last_error_ref() {
  static thread_local std::string s;
  return s;
}

int set_error(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

// ------------------------------------------------------------------ cached device allocations
namespace {
struct DevCache {
  std::mutex mu;
  std::map<int, std::multimap<size_t, void*>> free_blocks;  // device -> size -> block
  std::map<int, size_t> cached_bytes;
};
DevCache& dev_cache() {
  static DevCache c;
  return c;
}
constexpr size_t kCacheLimit = (size_t)48 << 30;  // per device
size_t round_block(size_t n) {
  const size_t g = n >= ((size_t)1 << 20) ? ((size_t)2 << 20) : 512;  // 2 MiB granules for large blocks
  return (n + g - 1) / g * g;
}
}  // namespace

cudaError_t dev_alloc(void** p, size_t bytes, size_t* granted) {
  const size_t want = round_block(bytes);
  int dev = 0;
  cudaGetDevice(&dev);
  DevCache& c = dev_cache();
  {
    std::lock_guard<std::mutex> lk(c.mu);
    auto& fb = c.free_blocks[dev];
    auto it = fb.lower_bound(want);
    if (it != fb.end() && it->first <= want + want / 4 + 4096) {  // best fit, at most 25 % slack
      *p = it->second;
      *granted = it->first;
      c.cached_bytes[dev] -= it->first;
      fb.erase(it);
      return cudaSuccess;
    }
  }
  cudaError_t e = cudaMalloc(p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    dev_cache_trim();
    e = cudaMalloc(p, want);
  }
  if (e == cudaSuccess) *granted = want;
  return e;
}

void dev_free(void* p, size_t granted) {
  if (!p) return;
  int dev = 0;
  cudaGetDevice(&dev);
  DevCache& c = dev_cache();
  {
    std::lock_guard<std::mutex> lk(c.mu);
    if (c.cached_bytes[dev] + granted <= kCacheLimit) {
      c.free_blocks[dev].emplace(granted, p);
      c.cached_bytes[dev] += granted;
      return;
    }
  }
  cudaFree(p);
}

void dev_cache_trim() {
  int dev = 0;
  cudaGetDevice(&dev);
  DevCache& c = dev_cache();
  std::multimap<size_t, void*> blocks;
  {
    std::lock_guard<std::mutex> lk(c.mu);
    blocks.swap(c.free_blocks[dev]);
    c.cached_bytes[dev] = 0;
  }
  for (auto& kv : blocks) cudaFree(kv.second);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int make_tmap_u8_sw128(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t kbytes, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(LDW_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (kbytes % 128 != 0 || (reinterpret_cast<uintptr_t>(gptr) & 127) != 0 || box_rows == 0 || box_rows > 256)
    return set_error(LDW_ERR_INTERNAL, "tensor map: bad geometry rows=%llu kbytes=%llu box_rows=%u",
                     (unsigned long long)rows, (unsigned long long)kbytes, box_rows);
  cuuint64_t gdim[2] = {kbytes, rows};
  cuuint64_t gstride[1] = {kbytes};
  cuuint32_t box[2] = {128, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(gptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(LDW_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

}  // namespace ldw
