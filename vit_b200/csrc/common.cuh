// Shared device helpers for the vit_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>

#include "../../include/vit_b200.h"

#define VB_CHECK_LAUNCH()                                   \
  do {                                                      \
    cudaError_t e__ = cudaPeekAtLastError();                \
    if (e__ != cudaSuccess) return vb_cuda_error(e__);      \
  } while (0)

int vb_cuda_error(cudaError_t e);  // records the message for vitb200_last_cuda_error(), returns VITB200_ERR_CUDA

#include <stdlib.h>

namespace vb {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  The step is a chain of ~20 short kernels; with a plain launch each one
// pays launch latency + block scheduling + its own prologue after the previous kernel has drained.  Kernels
// launched through vb_launch_pdl() may start while their predecessor is still running: they do their
// data-independent prologue (mbarrier init, descriptor prefetch, weight staging) and then block in pdl_wait()
// until every earlier kernel of the stream has completed and its writes are visible.
// Rules: (1) pdl_trigger() is called only AFTER pdl_wait() has returned.  Kernel k+1 can therefore start no earlier
//            than the moment kernel k got past its own wait, i.e. when kernels <= k-1 have completed: at most two
//            kernels of the chain are resident at once (k running, k+1 in its prologue).
//        (2) before pdl_wait() a kernel may only read data that its IMMEDIATE predecessor does not write.  Weights
//            are written only by the optimizer kernel; the only kernel that can directly follow it is the
//            embedding kernel, which therefore stages its weights after the wait.  Every other kernel stages
//            its weights / bias vectors before the wait.
//        (3) TMEM is allocated after pdl_wait(), so a waiting CTA never holds TMEM that a CTA of the still running
//            predecessor needs.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// Phase timeline (debug builds only: VITB200_TIMELINE=1 python -m vit_b200.build -> libvitb200_tl.so).
// Thread 0 of every CTA stamps clock64() at phase boundaries; tools/timeline.py prints the per-phase cycles.
// ---------------------------------------------------------------------------------------------
#ifdef VB_TIMELINE
#define VB_TL_SLOTS 32
#define VB_TL_CTAS 512
#define VB_TL_DECL(name) __device__ long long name[VB_TL_CTAS * VB_TL_SLOTS];
#define VB_TL(name, k)                                                                                        \
  do {                                                                                                        \
    const int cta_ = blockIdx.y * gridDim.x + blockIdx.x;                                                     \
    if (threadIdx.x == 0 && cta_ < VB_TL_CTAS) name[cta_ * VB_TL_SLOTS + (k)] = clock64();                    \
  } while (0)
// the same stamp taken by thread `t` (e.g. the first thread of a side warp group)
#define VB_TL_T(name, k, t)                                                                                   \
  do {                                                                                                        \
    const int cta_ = blockIdx.y * gridDim.x + blockIdx.x;                                                     \
    if (threadIdx.x == (t) && cta_ < VB_TL_CTAS) name[cta_ * VB_TL_SLOTS + (k)] = clock64();                  \
  } while (0)
#define VB_TL_EXPORT(fn, name)                                                                                \
  extern "C" int fn(long long* host) {                                                                        \
    return cudaMemcpyFromSymbol(host, name, sizeof(long long) * VB_TL_CTAS * VB_TL_SLOTS) == cudaSuccess ? 0 : -1; \
  }
#else
#define VB_TL_DECL(name)
#define VB_TL(name, k) do { } while (0)
#define VB_TL_T(name, k, t) do { } while (0)
#define VB_TL_EXPORT(fn, name)
#endif

// ---------------------------------------------------------------------------------------------
// 4-wide typed loads/stores (activations are float or bf16; accumulation is always fp32)
// ---------------------------------------------------------------------------------------------
template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ float4 ld(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void st(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct Vec4<bf16> {
  static __device__ __forceinline__ float4 ld(const bf16* p) {
    uint2 r = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  }
  static __device__ __forceinline__ void st(bf16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
  }
};

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// round-trip through the activation type (identity for float): mirrors autocast's bf16 op outputs
template <typename T> __device__ __forceinline__ float round_to(float v) { return to_f<T>(from_f<T>(v)); }

__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
__device__ __forceinline__ bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; }
template <typename T> __device__ __forceinline__ bool aligned_vec4(const T* p) {
  return sizeof(T) == 4 ? aligned16(p) : aligned8(p);
}

// ---------------------------------------------------------------------------------------------
// warp / block reductions (fixed order => deterministic)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int W> __device__ __forceinline__ float group_sum(float v) {  // sum over W consecutive lanes
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int W> __device__ __forceinline__ float group_max(float v) {
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-7 counter RNG for dropout (7 rounds is the Crush-resistant minimum of Salmon et al. 2011; the extra three
// rounds of the -10 default are safety margin that costs 30 % of the instructions of the attention epilogues).
// Masks are never stored: forward and backward regenerate them from (seed, step, site, element index).
//   key     = seed (64 bit)
//   counter = { blk.lo, blk.hi, site, step }   with blk = element_index / 8
// One 128-bit block decides EIGHT elements: element e uses the 16-bit half (e & 1) of word ((e >> 1) & 3) of the block
// at blk = e >> 3;  keep(e) <=> half >= round(p * 2^16)   (p is honoured to 2^-17: 0.1 -> 0.100006).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_7(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return make_uint4(c0, c1, c2, c3);
}

// attention-probability masks are indexed [b, head, query, key] with the key axis padded to a multiple of 8
__host__ __device__ __forceinline__ int attn_drop_tpad(int T) { return (T + 7) & ~7; }

struct DropCtx {
  uint32_t k0, k1, site, step;
  uint32_t thresh;  // keep <=> half >= thresh ; thresh = round(p * 2^16)
  float scale;      // 1 / (1 - p)
  bool on;
};
__device__ __forceinline__ DropCtx make_drop(float p, uint64_t seed, uint32_t step, uint32_t site) {
  DropCtx d;
  d.on = p > 0.f;
  d.k0 = (uint32_t)seed; d.k1 = (uint32_t)(seed >> 32);
  d.site = site; d.step = step;
  const float t = p * 65536.f + 0.5f;
  d.thresh = t >= 65536.f ? 65536u : (uint32_t)t;   // 65536 => nothing is kept (p = 1)
  d.scale = p < 1.f ? 1.f / (1.f - p) : 0.f;
  return d;
}
__device__ __forceinline__ float2 drop_word(const DropCtx& d, uint32_t w) {
  return make_float2((w & 0xffffu) >= d.thresh ? d.scale : 0.f, (w >> 16) >= d.thresh ? d.scale : 0.f);
}
// 8 keep-multipliers (0 or 1/(1-p)) for elements [8*idx8, 8*idx8+7]
__device__ __forceinline__ void drop8(const DropCtx& d, uint64_t idx8, float (&k)[8]) {
  if (!d.on) {
#pragma unroll
    for (int i = 0; i < 8; ++i) k[i] = 1.f;
    return;
  }
  const uint4 r = philox4x32_7((uint32_t)idx8, (uint32_t)(idx8 >> 32), d.site, d.step, d.k0, d.k1);
  const float2 a = drop_word(d, r.x), b = drop_word(d, r.y), c = drop_word(d, r.z), e = drop_word(d, r.w);
  k[0] = a.x; k[1] = a.y; k[2] = b.x; k[3] = b.y; k[4] = c.x; k[5] = c.y; k[6] = e.x; k[7] = e.y;
}
// 4 keep-multipliers for elements [4*idx4, 4*idx4+3]
__device__ __forceinline__ float4 drop4(const DropCtx& d, uint64_t idx4) {
  if (!d.on) return make_float4(1.f, 1.f, 1.f, 1.f);
  const uint64_t idx8 = idx4 >> 1;
  const uint4 r = philox4x32_7((uint32_t)idx8, (uint32_t)(idx8 >> 32), d.site, d.step, d.k0, d.k1);
  const float2 a = drop_word(d, (idx4 & 1) ? r.z : r.x), b = drop_word(d, (idx4 & 1) ? r.w : r.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float drop1(const DropCtx& d, uint64_t idx) {
  if (!d.on) return 1.f;
  const uint64_t idx8 = idx >> 3;
  const uint4 r = philox4x32_7((uint32_t)idx8, (uint32_t)(idx8 >> 32), d.site, d.step, d.k0, d.k1);
  const uint32_t q = (uint32_t)(idx >> 1) & 3u;
  const uint32_t w = q == 0 ? r.x : q == 1 ? r.y : q == 2 ? r.z : r.w;
  const uint32_t h = (idx & 1) ? (w >> 16) : (w & 0xffffu);
  return h >= d.thresh ? d.scale : 0.f;
}

// ---------------------------------------------------------------------------------------------
// erf-GELU (HF hidden_act='gelu' => torch.nn.functional.gelu, exact form): gelu(x) = x * Phi(x).
// Phi and phi share exp(-x^2/2); erf uses Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, i.e. below fp32 noise of
// the surrounding GEMMs): ~14 instructions for both, instead of ~50 for erff + expf.  The GELU epilogues are the
// longest serial instruction streams of the fused kernels (128 columns per thread), so this matters.
// ---------------------------------------------------------------------------------------------
// MUFU wrappers: one instruction each (expf / division without --use_fast_math expand to range-fixup sequences)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// ht = 0.5 * erfc(|x| / sqrt 2) = 1 - Phi(|x|)   and   e = exp(-x^2 / 2)
__device__ __forceinline__ float gelu_half_tail(float x, float& e) {
  e = ex2_approx(x * x * -0.72134752044448170f);                  // -0.5 * log2(e)
  const float t = rcp_approx(fmaf(0.3275911f * 0.70710678118654752f, fabsf(x), 1.f));
  float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);   // 0.5 folded into the coefficients
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  return p * t * e;
}
__device__ __forceinline__ void gelu_cdf_pdf(float x, float& cdf, float& pdf) {
  float e;
  const float ht = gelu_half_tail(x, e);
  cdf = 0.5f + copysignf(0.5f - ht, x);
  pdf = 0.39894228040143268f * e;
}
// gelu(x) = x Phi(x) = max(x, 0) - |x| (1 - Phi(|x|))
__device__ __forceinline__ float gelu_f(float x) {
  float e;
  const float ht = gelu_half_tail(x, e);
  return fmaf(-fabsf(x), ht, fmaxf(x, 0.f));
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  float c, d;
  gelu_cdf_pdf(x, c, d);
  return fmaf(x, d, c);
}

// ---------------------------------------------------------------------------------------------
// "last block done" ticket: returns true in exactly one block (the last to arrive), after all
// other blocks' global writes are visible.  The counter resets itself to 0 for the next launch.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool last_block_ticket(unsigned int* counter, unsigned int nblocks) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    unsigned int t = atomicAdd(counter, 1u);
    is_last = (t == nblocks - 1);
    if (is_last) *counter = 0u;
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// launch with the programmatic-stream-serialization attribute (see pdl_wait above); the kernel MUST call pdl_wait()
template <typename... KArgs, typename... Args>
static inline void vb_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool no_pdl = getenv("VITB200_NO_PDL") != nullptr;   // debugging aid: plain stream-ordered launches
  cfg.attrs = at;
  cfg.numAttrs = no_pdl ? 0 : 1;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// the same with a thread-block cluster of `cluster` CTAs along x (1 = no cluster attribute)
template <typename... KArgs, typename... Args>
static inline void vb_launch_pdl_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                         int cluster, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  static const bool no_pdl = getenv("VITB200_NO_PDL") != nullptr;
  if (!no_pdl) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = (unsigned)cluster;
    at[n].val.clusterDim.y = 1;
    at[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace vb
