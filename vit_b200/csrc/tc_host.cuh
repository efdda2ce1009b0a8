// Host side of the TMA path: cuTensorMapEncodeTiled through the runtime's driver entry point (no libcuda
// link dependency) and a small cache of encoded maps keyed by (pointer, shape, box).  The engine calls the
// library with the same buffers every step, so after the first step every lookup is a hash hit; under CUDA
// graph replay the maps are baked into the captured kernel parameters.
#pragma once
#include <mutex>
#include <unordered_map>
#include <string.h>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/vit_b200.h"

namespace vb {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

struct TmKey {
  const void* p; uint64_t inner, outer; uint32_t box_inner, box_outer;
  bool operator==(const TmKey& o) const {
    return p == o.p && inner == o.inner && outer == o.outer && box_inner == o.box_inner && box_outer == o.box_outer;
  }
};
struct TmHash {
  size_t operator()(const TmKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.p);
    h = h * 1000003u ^ k.inner; h = h * 1000003u ^ k.outer; h = h * 1000003u ^ k.box_inner; h = h * 1000003u ^ k.box_outer;
    return h;
  }
};

// 2-D bf16 row-major matrix [outer rows, inner cols] (pitch = inner), 128B swizzle, zero fill out of bounds
static int get_tmap(const void* p, uint64_t inner, uint64_t outer, uint32_t box_inner, uint32_t box_outer,
                    CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<TmKey, CUtensorMap, TmHash> cache;
  TmKey key{p, inner, outer, box_inner, box_outer};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return VITB200_OK; }
  }
  PFN_encodeTiled enc = get_encode();
  if (!enc) return VITB200_ERR_DEVICE;
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {inner * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap tm;
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return VITB200_ERR_ARG;
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = tm;
  }
  *out = tm;
  return VITB200_OK;
}


}  // namespace vb
