// Gradient-norm clipping + AdamW over one flat fp32 parameter arena (HBM-bound streaming kernels),
// bf16 shadow refresh for BF16-mode GEMM operands, and the dropout-mask test helper.
#include <string.h>

#include "common.cuh"

namespace vb {

constexpr int OP_THREADS = 256;
constexpr int OP_MAX_BLOCKS = 592;

static inline int op_grid(size_t n) {
  size_t g = (n / 4 + OP_THREADS * 4 - 1) / (OP_THREADS * 4);
  if (g > OP_MAX_BLOCKS) g = OP_MAX_BLOCKS;
  if (g < 1) g = 1;
  return (int)g;
}

// hyper: {lr, b1, b2, eps, wd, max_norm, grad_scale, -} ; state: {step, norm, coef, bc1, bc2, -, -, -}
__global__ void __launch_bounds__(OP_THREADS)
grad_norm_kernel(const float* __restrict__ g, size_t n4, const float* __restrict__ hyper, float* __restrict__ state,
                 float* __restrict__ partial, unsigned int* counter) {
  __shared__ float red[OP_THREADS / 32];
  pdl_wait();
  pdl_trigger();
  float acc = 0.f;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (size_t i = (size_t)blockIdx.x * OP_THREADS + threadIdx.x; i < n4; i += (size_t)gridDim.x * OP_THREADS) {
    float4 v = g4[i];
    acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < OP_THREADS / 32; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
  if (!last_block_ticket(counter, gridDim.x)) return;
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (unsigned int b = 0; b < gridDim.x; ++b) tot += (double)__ldcg(&partial[b]);
    const float gs = hyper[6];
    const float norm = (float)sqrt(tot) * fabsf(gs);
    const float max_norm = hyper[5];
    float coef = 1.f;
    if (max_norm > 0.f) coef = fminf(1.f, max_norm / (norm + 1e-6f));
    const float step = state[0] + 1.f;
    state[0] = step;
    state[1] = norm;
    state[2] = coef;
    state[3] = (float)(1.0 - pow((double)hyper[1], (double)step));
    state[4] = (float)(1.0 - pow((double)hyper[2], (double)step));
  }
}

__global__ void __launch_bounds__(OP_THREADS)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             bf16* __restrict__ shadow, size_t n4, const float* __restrict__ hyper, const float* __restrict__ state,
             uint64_t* rng) {
  pdl_wait();
  pdl_trigger();
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
  const float gmul = hyper[6] * state[2];
  const float bc1 = state[3], bc2 = state[4];
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = 1.f / sqrtf(bc2);
  const float decay = 1.f - lr * wd;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (size_t i = (size_t)blockIdx.x * OP_THREADS + threadIdx.x; i < n4; i += (size_t)gridDim.x * OP_THREADS) {
    float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
    float pa[4] = {pp.x, pp.y, pp.z, pp.w}, ga[4] = {gg.x, gg.y, gg.z, gg.w};
    float ma[4] = {mm.x, mm.y, mm.z, mm.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = ga[k] * gmul;
      pa[k] *= decay;
      ma[k] = fmaf(1.f - b1, gr - ma[k], ma[k]);          // exp_avg.lerp_(grad, 1 - beta1)
      va[k] = fmaf(1.f - b2, gr * gr, b2 * va[k]);         // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
      const float denom = sqrtf(va[k]) * inv_sqrt_bc2 + eps;
      pa[k] -= step_size * (ma[k] / denom);
    }
    p4[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
    m4[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    v4[i] = make_float4(va[0], va[1], va[2], va[3]);
    if (shadow) Vec4<bf16>::st(shadow + i * 4, make_float4(pa[0], pa[1], pa[2], pa[3]));
  }
  if (rng && blockIdx.x == 0 && threadIdx.x == 0) rng[1] += 1ull;
}

// ------------------------------------------------------------------------------------------------
// One-launch optimizer tail:  [sum of the per-CTA gradient partials] -> global L2 norm -> clip -> AdamW.
// The norm is a grid-wide dependency; the grid is at most one CTA per SM (all co-resident), so a ticket + epoch flag
// in global memory is a safe grid barrier: block b publishes its partial sum, takes a ticket; the last arrival
// reduces the partials in block order (deterministic), writes the state and bumps the epoch that the others poll.
// Each thread keeps its (<= TAIL_KEEP) reduced gradient vectors in registers across the barrier.
// ------------------------------------------------------------------------------------------------
// ---- peer (NVLink) exchange buffer of the data-parallel optimizer tail ------------------------------------------
// layout: [1024 B header: uint32 flags[world]] [parity 0: xfloats fp32] [parity 1: xfloats fp32]
struct PeerX {
  void* const* bufs;   // DEVICE array [world]: every rank's exchange buffer (own entry = local memory)
  int rank, world;     // world <= 1: no exchange
  size_t xfloats;
};
__device__ __forceinline__ unsigned int* peer_flags(void* buf) { return reinterpret_cast<unsigned int*>(buf); }
__device__ __forceinline__ float* peer_data(void* buf, unsigned int parity, size_t xfloats) {
  return reinterpret_cast<float*>(reinterpret_cast<char*>(buf) + 1024) + (size_t)parity * xfloats;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_relaxed_sys_f4(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
constexpr int TAIL_THREADS = 256;
constexpr int TAIL_KEEP = 4;
constexpr int TAIL_MAX_BLOCKS = 148;

__global__ void __launch_bounds__(TAIL_THREADS)
clip_adamw_fused_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                        bf16* __restrict__ shadow, size_t n4, const float* __restrict__ hyper, float* __restrict__ state,
                        uint64_t* rng, const float* __restrict__ gpart, int slots, size_t stride4, size_t red_lo4,
                        size_t red_hi4, float* __restrict__ partial, unsigned int* sync, const PeerX X) {
  __shared__ float red[TAIL_THREADS / 32];
  __shared__ bool is_last;
  pdl_wait();
  pdl_trigger();
  unsigned int* ticket = sync;
  volatile unsigned int* epoch = sync + 1;
  const unsigned int my_epoch = *epoch;   // read BEFORE this block's ticket: the epoch cannot move until every block arrived
  const unsigned int seq = *(volatile unsigned int*)(sync + 2) + 1u;   // launch number (same on every rank): exchange tag
  float4* g4 = reinterpret_cast<float4*>(g);
  const size_t gstride = (size_t)gridDim.x * TAIL_THREADS;
  const size_t i0 = (size_t)blockIdx.x * TAIL_THREADS + threadIdx.x;
  float4 keep[TAIL_KEEP];
  float acc = 0.f;
  int k = 0;
  for (size_t i = i0; i < n4; i += gstride, ++k) {
    float4 s;
    if (slots > 0 && i >= red_lo4 && i < red_hi4) {   // gradient = sum over the CTA partial slots, in slot order
      const float4* src = reinterpret_cast<const float4*>(gpart) + i;
      s = make_float4(0.f, 0.f, 0.f, 0.f);
      int z = 0;
      for (; z + 16 <= slots; z += 16) {   // 16 independent L2 reads in flight per thread, summed in slot order
        float4 t[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) t[q] = __ldcg(src + (size_t)(z + q) * stride4);
#pragma unroll
        for (int q = 0; q < 16; ++q) { s.x += t[q].x; s.y += t[q].y; s.z += t[q].z; s.w += t[q].w; }
      }
      for (; z + 4 <= slots; z += 4) {
        float4 t[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) t[q] = __ldcg(src + (size_t)(z + q) * stride4);
#pragma unroll
        for (int q = 0; q < 4; ++q) { s.x += t[q].x; s.y += t[q].y; s.z += t[q].z; s.w += t[q].w; }
      }
      for (; z < slots; ++z) {
        const float4 t = __ldcg(src + (size_t)z * stride4);
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
      }
      g4[i] = s;   // the flat gradient arena stays the public result (p.grad views, DDP, tests)
    } else {
      s = g4[i];
    }
    if (X.world > 1) {   // data parallel: publish the local gradient, the sum over ranks is formed below
      reinterpret_cast<float4*>(peer_data(X.bufs[X.rank], seq & 1u, X.xfloats))[i] = s;
      continue;
    }
    if (k < TAIL_KEEP) keep[k] = s;
    acc += (s.x * s.x + s.y * s.y) + (s.z * s.z + s.w * s.w);
  }
  if (X.world > 1) {
    // ---- gradient all-reduce over NVLink peer memory, fused into this kernel (no NCCL launch on the step's critical
    // path).  1. every block fences its slice of the published gradient and takes a ticket; the last one raises this
    // rank's flag (= seq) in EVERY rank's buffer.  2. all blocks wait until all `world` flags in the local buffer
    // reached seq.  3. each element is summed over the ranks' buffers in rank order -- every rank computes bit-identical
    // sums, so the replicas cannot drift.  Buffers are double-buffered by launch parity: a rank can only overwrite
    // parity p two launches later, after a full flag round in between proved that every peer finished reading it.
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int tk = atomicAdd(ticket, 1u);
      if (tk == gridDim.x - 1) {
        *ticket = 0u;
        __threadfence_system();
        for (int q = 0; q < X.world; ++q) st_release_sys(peer_flags(X.bufs[q]) + X.rank, seq);
      }
    }
    if (threadIdx.x < X.world) {
      const unsigned int* f = peer_flags(X.bufs[X.rank]) + threadIdx.x;
      unsigned int spins = 0;
      while ((int)(ld_acquire_sys(f) - seq) < 0) {
        if (++spins > (1u << 28)) __trap();   // a dead peer traps instead of hanging the device
      }
      __threadfence_system();
    }
    __syncthreads();
    k = 0;
    for (size_t i = i0; i < n4; i += gstride, ++k) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int q = 0; q < X.world; ++q) {
        const float4 t = ld_relaxed_sys_f4(reinterpret_cast<const float4*>(peer_data(X.bufs[q], seq & 1u, X.xfloats)) + i);
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
      }
      g4[i] = s;   // summed gradient (the 1/world mean is hyper[6] = grad_scale)
      if (k < TAIL_KEEP) keep[k] = s;
      acc += (s.x * s.x + s.y * s.y) + (s.z * s.z + s.w * s.w);
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < TAIL_THREADS / 32; ++w) t += red[w];
    partial[blockIdx.x] = t;
    __threadfence();
    const unsigned int tk = atomicAdd(ticket, 1u);
    is_last = (tk == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    double tot = 0.0;
    for (unsigned int b = 0; b < gridDim.x; ++b) tot += (double)__ldcg(&partial[b]);
    const float gs = hyper[6];
    const float norm = (float)sqrt(tot) * fabsf(gs);
    const float max_norm = hyper[5];
    float coef = 1.f;
    if (max_norm > 0.f) coef = fminf(1.f, max_norm / (norm + 1e-6f));
    const float step = state[0] + 1.f;
    state[0] = step;
    state[1] = norm;
    state[2] = coef;
    state[3] = (float)(1.0 - pow((double)hyper[1], (double)step));
    state[4] = (float)(1.0 - pow((double)hyper[2], (double)step));
    if (rng) rng[1] += 1ull;
    *ticket = 0u;
    sync[2] = seq;
    __threadfence();
    *epoch = my_epoch + 1u;   // release
  }
  if (threadIdx.x == 0) {
    unsigned int spins = 0;
    while (*epoch == my_epoch) {
      if (++spins > (1u << 28)) __trap();   // a mis-sized grid traps instead of hanging the device
    }
    __threadfence();
  }
  __syncthreads();
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
  const float gmul = hyper[6] * __ldcg(&state[2]);
  const float bc1 = __ldcg(&state[3]), bc2 = __ldcg(&state[4]);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = 1.f / sqrtf(bc2);
  const float decay = 1.f - lr * wd;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  k = 0;
  for (size_t i = i0; i < n4; i += gstride, ++k) {
    const float4 gg = k < TAIL_KEEP ? keep[k] : __ldcg(&g4[i]);
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    float pa[4] = {pp.x, pp.y, pp.z, pp.w}, ga[4] = {gg.x, gg.y, gg.z, gg.w};
    float ma[4] = {mm.x, mm.y, mm.z, mm.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float gr = ga[q] * gmul;
      pa[q] *= decay;
      ma[q] = fmaf(1.f - b1, gr - ma[q], ma[q]);          // exp_avg.lerp_(grad, 1 - beta1)
      va[q] = fmaf(1.f - b2, gr * gr, b2 * va[q]);         // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
      const float denom = sqrtf(va[q]) * inv_sqrt_bc2 + eps;
      pa[q] -= step_size * (ma[q] / denom);
    }
    p4[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
    m4[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    v4[i] = make_float4(va[0], va[1], va[2], va[3]);
    if (shadow) Vec4<bf16>::st(shadow + i * 4, make_float4(pa[0], pa[1], pa[2], pa[3]));
  }
}

__global__ void __launch_bounds__(OP_THREADS)
cast_bf16_kernel(const float* __restrict__ p, bf16* __restrict__ shadow, size_t n4) {
  const float4* p4 = reinterpret_cast<const float4*>(p);
  for (size_t i = (size_t)blockIdx.x * OP_THREADS + threadIdx.x; i < n4; i += (size_t)gridDim.x * OP_THREADS)
    Vec4<bf16>::st(shadow + i * 4, p4[i]);
}

__global__ void __launch_bounds__(OP_THREADS)
dropout_mask_kernel(uint8_t* __restrict__ mask, size_t n, float p_drop, const uint64_t* __restrict__ rng, uint32_t site) {
  const DropCtx dc = make_drop(p_drop, rng ? rng[0] : 0ull, rng ? (uint32_t)rng[1] : 0u, site);
  for (size_t i = (size_t)blockIdx.x * OP_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * OP_THREADS)
    mask[i] = drop1(dc, i) > 0.f ? 1 : 0;
}

}  // namespace vb

using namespace vb;

extern "C" size_t vitb200_grad_norm_ws_bytes(size_t n) {
  (void)n;
  return 4096 + OP_MAX_BLOCKS * sizeof(float);
}

extern "C" int vitb200_grad_norm(const float* g, size_t n, const float* hyper, float* state, void* ws, void* stream) {
  if (!g || !hyper || !state || !ws) return VITB200_ERR_ARG;
  if (n % 4 != 0) return VITB200_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(g) & 15) != 0) return VITB200_ERR_ALIGN;
  unsigned int* counter = reinterpret_cast<unsigned int*>(ws);
  float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + 4096);
  vb_launch_pdl(grad_norm_kernel, dim3(op_grid(n)), dim3(OP_THREADS), 0, (cudaStream_t)stream, g, n / 4, hyper, state, partial, counter);
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

extern "C" int vitb200_adamw(float* p, const float* g, float* m, float* v, void* shadow, size_t n, const float* hyper,
                             const float* state, uint64_t* rng, void* stream) {
  if (!p || !g || !m || !v || !hyper || !state) return VITB200_ERR_ARG;
  if (n % 4 != 0) return VITB200_ERR_SHAPE;
  if (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
        reinterpret_cast<uintptr_t>(v)) & 15) != 0 || (reinterpret_cast<uintptr_t>(shadow) & 7) != 0)
    return VITB200_ERR_ALIGN;
  vb_launch_pdl(adamw_kernel, dim3(op_grid(n)), dim3(OP_THREADS), 0, (cudaStream_t)stream, p, g, m, v, (bf16*)shadow, n / 4, hyper, state, rng);
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

extern "C" size_t vitb200_clip_adamw_fused_ws_bytes(void) { return 4096 + TAIL_MAX_BLOCKS * sizeof(float); }

extern "C" int vitb200_clip_adamw_fused(float* p, float* g, float* m, float* v, void* shadow, size_t n, const float* hyper,
                                        float* state, uint64_t* rng, const float* gpart, int slots, size_t stride,
                                        size_t red_start, size_t red_end, void* ws, void* stream) {
  if (!p || !g || !m || !v || !hyper || !state || !ws) return VITB200_ERR_ARG;
  if (n % 4 != 0) return VITB200_ERR_SHAPE;
  if (slots > 0 && (!gpart || (stride | red_start | red_end) % 4 != 0 || red_end < red_start || red_end > n)) return VITB200_ERR_ARG;
  if (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
        reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(gpart)) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(shadow) & 7) != 0)
    return VITB200_ERR_ALIGN;
  const size_t n4 = n / 4;
  size_t grid = (n4 + TAIL_THREADS - 1) / TAIL_THREADS;
  if (grid > TAIL_MAX_BLOCKS) grid = TAIL_MAX_BLOCKS;   // <= one CTA per SM: the in-kernel grid barrier needs co-residency
  if (grid < 1) grid = 1;
  unsigned int* sync = reinterpret_cast<unsigned int*>(ws);   // {ticket, epoch}: zero-initialised by the caller once
  float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + 4096);
  vb_launch_pdl(clip_adamw_fused_kernel, dim3((unsigned)grid), dim3(TAIL_THREADS), 0, (cudaStream_t)stream, p, g, m, v,
                (bf16*)shadow, n4, hyper, state, rng, gpart, slots, stride / 4, red_start / 4, red_end / 4, partial, sync,
                PeerX{nullptr, 0, 1, 0});
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

// ---- data-parallel variant: the gradient all-reduce runs inside the kernel over peer memory ---------------------
extern "C" size_t vitb200_peer_buffer_bytes(size_t n) { return 1024 + 2 * n * sizeof(float); }

extern "C" int vitb200_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
  if (!ptr || !handle64 || bytes == 0) return VITB200_ERR_ARG;
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e != cudaSuccess) return vb_cuda_error(e);
  if ((e = cudaMemset(*ptr, 0, bytes)) != cudaSuccess) return vb_cuda_error(e);
  cudaIpcMemHandle_t h;
  if ((e = cudaIpcGetMemHandle(&h, *ptr)) != cudaSuccess) return vb_cuda_error(e);
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  if ((e = cudaDeviceSynchronize()) != cudaSuccess) return vb_cuda_error(e);
  return VITB200_OK;
}
extern "C" int vitb200_peer_open(const unsigned char* handle64, void** ptr) {
  if (!ptr || !handle64) return VITB200_ERR_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return vb_cuda_error(e);
  return VITB200_OK;
}
extern "C" int vitb200_peer_close(void* ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  return e == cudaSuccess ? VITB200_OK : vb_cuda_error(e);
}
extern "C" int vitb200_peer_free(void* ptr) {
  cudaError_t e = cudaFree(ptr);
  return e == cudaSuccess ? VITB200_OK : vb_cuda_error(e);
}

extern "C" int vitb200_clip_adamw_fused_dp(float* p, float* g, float* m, float* v, void* shadow, size_t n, const float* hyper,
                                           float* state, uint64_t* rng, const float* gpart, int slots, size_t stride,
                                           size_t red_start, size_t red_end, void* ws, void* const* peer_bufs, int rank,
                                           int world, void* stream) {
  if (!p || !g || !m || !v || !hyper || !state || !ws || !peer_bufs) return VITB200_ERR_ARG;
  if (world < 2 || world > TAIL_THREADS || rank < 0 || rank >= world) return VITB200_ERR_ARG;
  if (n % 4 != 0) return VITB200_ERR_SHAPE;
  if (slots > 0 && (!gpart || (stride | red_start | red_end) % 4 != 0 || red_end < red_start || red_end > n)) return VITB200_ERR_ARG;
  if (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
        reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(gpart)) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(shadow) & 7) != 0)
    return VITB200_ERR_ALIGN;
  const size_t n4 = n / 4;
  size_t grid = (n4 + TAIL_THREADS - 1) / TAIL_THREADS;
  if (grid > TAIL_MAX_BLOCKS) grid = TAIL_MAX_BLOCKS;
  if (grid < 1) grid = 1;
  unsigned int* sync = reinterpret_cast<unsigned int*>(ws);
  float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + 4096);
  vb_launch_pdl(clip_adamw_fused_kernel, dim3((unsigned)grid), dim3(TAIL_THREADS), 0, (cudaStream_t)stream, p, g, m, v,
                (bf16*)shadow, n4, hyper, state, rng, gpart, slots, stride / 4, red_start / 4, red_end / 4, partial, sync,
                PeerX{peer_bufs, rank, world, n});
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

extern "C" int vitb200_cast_bf16(const float* p, void* shadow, size_t n, void* stream) {
  if (!p || !shadow) return VITB200_ERR_ARG;
  if (n % 4 != 0) return VITB200_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(p) & 15) != 0 || (reinterpret_cast<uintptr_t>(shadow) & 7) != 0) return VITB200_ERR_ALIGN;
  if (n == 0) return VITB200_OK;
  cast_bf16_kernel<<<op_grid(n), OP_THREADS, 0, (cudaStream_t)stream>>>(p, (bf16*)shadow, n / 4);
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

extern "C" int vitb200_dropout_mask(uint8_t* mask, size_t n, float p_drop, const uint64_t* rng, uint32_t site,
                                    void* stream) {
  if (!mask) return VITB200_ERR_ARG;
  if (n == 0) return VITB200_OK;
  size_t g = (n + OP_THREADS - 1) / OP_THREADS;
  if (g > 4096) g = 4096;
  dropout_mask_kernel<<<(int)g, OP_THREADS, 0, (cudaStream_t)stream>>>(mask, n, p_drop, rng, site);
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}
