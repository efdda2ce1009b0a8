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

// acc[0] += sum g^2 (fixed-order, deterministic): the squared gradient norm of parameters that live OUTSIDE the arena
// (a trainable preprocessor matrix), handed to the one-launch optimizer tail through state[5]
__global__ void __launch_bounds__(OP_THREADS)
sumsq_accum_kernel(const float* __restrict__ g, size_t n4, float* __restrict__ acc_out, float* __restrict__ partial,
                   unsigned int* counter) {
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
    acc_out[0] += (float)tot;
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
// The norm is a grid-wide dependency; the grid is at most one CTA per SM (all co-resident), so spinning on global memory
// is a safe grid barrier: block b publishes {partial sum, launch number} in one 8-byte store, block 0 collects the slots
// in a fixed order (deterministic) and publishes {norm, launch number} the same way, everybody else polls that slot.
// Each thread keeps its (<= TAIL_KEEP) reduced gradient vectors in registers across the barrier.
// ------------------------------------------------------------------------------------------------
// ---- peer (NVLink) exchange buffer of the data-parallel optimizer tail ------------------------------------------
// layout: [16 KB header (unused words kept for alignment)] [parity 0: world_max x n slots] [parity 1: ...]; a slot is
// 8 bytes {gradient value bits, launch tag}.  Rank r PUSHES every reduced gradient element, tagged with the launch number,
// into slot [parity][r][i] of EVERY rank's buffer with ONE 8-byte store (value and tag travel together, so no flag, no
// release fence, no acknowledgement round trip -- the "LL" idea of NCCL); a rank polls its OWN buffer until the tags of
// all ranks' slots of an element equal the launch number and sums the values in rank order.
constexpr size_t PEER_HEADER = 16384;
constexpr int PEER_MAX_WORLD = 8;   // one NVLink / NVSwitch node
// streamed gradient groups (vitb200_grad_stream): counters raised by the backward kernel, arena ranges in float4 units
struct GradStream {
  unsigned int* done;
  unsigned int expect;
  int n;
  unsigned int lo4[VITB200_MAX_GROUPS], hi4[VITB200_MAX_GROUPS];
};
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct PeerX {
  void* const* bufs;   // DEVICE array [world]: every rank's exchange buffer (own entry = local memory)
  int rank, world;     // world <= 1: no exchange
  size_t xfloats;
};
__device__ __forceinline__ uint2* peer_slots(void* buf, unsigned int parity, int src, size_t xfloats) {
  return reinterpret_cast<uint2*>(reinterpret_cast<char*>(buf) + PEER_HEADER) + ((size_t)parity * PEER_MAX_WORLD + src) * xfloats;
}
__device__ __forceinline__ void st_relaxed_sys_v2(uint2* p, uint32_t a, uint32_t b) {
  asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint2 ld_relaxed_sys_v2(const uint2* p) {
  uint2 v;
  asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
VB_TL_DECL(tl_tail)
constexpr int TAIL_MAX_THREADS = 512;         // block size is chosen per arena: four lanes share one float4 element and
                                              // the launcher picks 4 * ceil(n4 / 148) threads (a multiple of 32), so the
                                              // 40 k-parameter arena of the configured model is ONE pass over 148 SMs
constexpr int TAIL_KEEP = 4;
constexpr int TAIL_MAX_BLOCKS = 148;

// Thread (e, j): float4 element e of the arena, lane j = tid & 3.
//   reduction of the gradient partials: lane j sums slots j, j+4, ... (independent 16-byte L2 reads), the four partial
//     float4s are combined by a two-level butterfly (fixed order) and lane j keeps component j;
//   everything after that (peer exchange, norm, AdamW) is scalar work on element 4 e + j: consecutive lanes touch
//     consecutive floats, so all accesses stay coalesced and four times more threads hide the L2 / NVLink latency.
__global__ void __launch_bounds__(TAIL_MAX_THREADS)
clip_adamw_fused_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                        bf16* __restrict__ shadow, size_t n4, const float* __restrict__ hyper, float* __restrict__ state,
                        uint64_t* rng, const float* __restrict__ gpart, int slots, size_t stride4, size_t red_lo4,
                        size_t red_hi4, float* __restrict__ partial, unsigned int* sync, const PeerX X, const GradStream GS) {
  __shared__ float red[TAIL_MAX_THREADS / 32];
  VB_TL(tl_tail, 0);
  // Streamed mode: the backward kernel is still running.  Nothing it reads is written before the pdl_wait() further down
  // (behind the grid barrier, which no block passes before every gradient group has been signalled), and the next
  // kernel of the stream (the forward of the next step) touches no memory before its own dependency wait, so it may be
  // made resident early.
  if (GS.n == 0) pdl_wait();
  pdl_trigger();
  VB_TL(tl_tail, 1);
  // read BEFORE this block's ticket: block 0 rewrites them only after every block of this launch took its ticket
  const unsigned int seq = *(volatile unsigned int*)(sync + 2) + 1u;   // launch number (same on every rank): barrier / exchange tag
  const float step_prev = *(volatile float*)state;
  const float extra_sq = *(volatile float*)(state + 5);   // squared gradient norm of parameters outside the arena (vitb200_sumsq_accum)
  const int lane4 = threadIdx.x & 3;
  const size_t epb = blockDim.x >> 2;   // float4 elements per block pass
  const size_t estride = (size_t)gridDim.x * epb;
  const size_t e0 = (size_t)blockIdx.x * epb + (threadIdx.x >> 2);
  float keep[TAIL_KEEP];
  float acc = 0.f;
  if (GS.n > 0) {
    // wait for the gradient groups this block's elements belong to (one pass: a contiguous slice; several passes: all)
    if (threadIdx.x == 0) {
      const bool one_pass = estride >= n4;
      const size_t blo = one_pass ? (size_t)blockIdx.x * epb : 0;
      const size_t bhi = one_pass ? min(n4, blo + epb) : n4;
      for (int gi = 0; gi < GS.n; ++gi) {
        if (GS.lo4[gi] >= bhi || GS.hi4[gi] <= blo) continue;
        unsigned int spins = 0;
        while (ld_acquire_gpu(GS.done + gi) < GS.expect) {
          __nanosleep(64);
          if (++spins > (1u << 26)) { state[7] = 2.f; break; }   // a backward launch that never signals must not hang the device
        }
      }
    }
    __syncthreads();
    VB_TL(tl_tail, 9);
  }
  int k = 0;
  for (size_t e = e0; e < n4; e += estride, ++k) {
    float gj;
    if (slots > 0 && e >= red_lo4 && e < red_hi4) {
      const float4* src = reinterpret_cast<const float4*>(gpart) + e;
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      int z = lane4;
      for (; z + 60 < slots; z += 64) {   // 16 independent reads in flight
        float4 t[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) t[q] = __ldcg(src + (size_t)(z + 4 * q) * stride4);
#pragma unroll
        for (int q = 0; q < 16; ++q) { s.x += t[q].x; s.y += t[q].y; s.z += t[q].z; s.w += t[q].w; }
      }
      for (; z + 12 < slots; z += 16) {
        float4 t[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) t[q] = __ldcg(src + (size_t)(z + 4 * q) * stride4);
#pragma unroll
        for (int q = 0; q < 4; ++q) { s.x += t[q].x; s.y += t[q].y; s.z += t[q].z; s.w += t[q].w; }
      }
      for (; z < slots; z += 4) {
        const float4 t = __ldcg(src + (size_t)z * stride4);
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
      }
      // the four lanes of an element are adjacent lanes of one warp and always take this branch together
      const unsigned gm = 0xFu << (threadIdx.x & 28);
      s.x += __shfl_xor_sync(gm, s.x, 1); s.y += __shfl_xor_sync(gm, s.y, 1);
      s.z += __shfl_xor_sync(gm, s.z, 1); s.w += __shfl_xor_sync(gm, s.w, 1);
      s.x += __shfl_xor_sync(gm, s.x, 2); s.y += __shfl_xor_sync(gm, s.y, 2);
      s.z += __shfl_xor_sync(gm, s.z, 2); s.w += __shfl_xor_sync(gm, s.w, 2);
      gj = lane4 == 0 ? s.x : lane4 == 1 ? s.y : lane4 == 2 ? s.z : s.w;
    } else {
      gj = g[4 * e + lane4];
    }
    if (X.world > 1) {   // data parallel: push the local gradient element to every rank (own buffer included)
      const size_t i = 4 * e + lane4;
      for (int q = 0; q < X.world; ++q)
        st_relaxed_sys_v2(peer_slots(X.bufs[q], seq & 1u, X.rank, X.xfloats) + i, __float_as_uint(gj), seq);
      continue;
    }
    g[4 * e + lane4] = gj;   // the flat gradient arena stays the public result (p.grad views, tests)
    if (k < TAIL_KEEP) keep[k] = gj;
    acc = fmaf(gj, gj, acc);
  }
  if (X.world > 1) {
    // ---- gradient all-reduce over NVLink peer memory, fused into this kernel (no NCCL launch on the step's critical
    // path).  Every element was pushed above as {value, launch tag} into slot [parity][rank][i] of every rank's buffer;
    // here each thread polls the `world` slots of ITS element in the local buffer until all tags equal this launch and
    // adds the values in rank order -- every rank computes bit-identical sums, so the replicas cannot drift.  No flag,
    // no fence, no block barrier: an 8-byte store is a single transaction.  Buffers are double-buffered by launch
    // parity: a rank overwrites parity p two launches later, after it received the peers' data of the launch in between,
    // which they only sent after they had finished reading this one.
    VB_TL(tl_tail, 6);
    k = 0;
    for (size_t e = e0; e < n4; e += estride, ++k) {
      const size_t i = 4 * e + lane4;
      const uint2* slot = peer_slots(X.bufs[X.rank], seq & 1u, 0, X.xfloats) + i;
      uint2 t[PEER_MAX_WORLD];
      unsigned int spins = 0;
      bool all = false;
      while (!all) {
        all = true;
#pragma unroll
        for (int q = 0; q < PEER_MAX_WORLD; ++q)
          if (q < X.world) t[q] = ld_relaxed_sys_v2(slot + (size_t)q * X.xfloats);
#pragma unroll
        for (int q = 0; q < PEER_MAX_WORLD; ++q)
          if (q < X.world && t[q].y != seq) all = false;
        // a peer that never arrives (crashed rank) must not hang the device, and must not poison the context of the
        // surviving ranks either: give up after ~minutes, record the failure in state[7] (the host checks it,
        // PeerExchange.check()) and carry on with whatever the buffer holds
        if (!all && ++spins > (1u << 28)) { state[7] = 1.f; break; }
      }
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < PEER_MAX_WORLD; ++q)
        if (q < X.world) s += __uint_as_float(t[q].x);   // rank order
      g[i] = s;   // summed gradient (the 1/world mean is hyper[6] = grad_scale)
      if (k < TAIL_KEEP) keep[k] = s;
      acc = fmaf(s, s, acc);
    }
    VB_TL(tl_tail, 8);
  }
  VB_TL(tl_tail, 2);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  // ---- grid barrier + norm.  Tickets only ever count up: launch number seq is complete when the counter reaches
  // seq * gridDim (the workspace belongs to one arena, so the grid is the same at every launch).  EVERY block then sums
  // the block partials itself (same fixed order => same bits) and derives coef / bias corrections: there is no serial
  // "last block computes, everybody polls a second flag" hop.  Block 0 alone records the results.
  if (threadIdx.x < 32) {
    // everything the norm step needs besides the block partials is fetched BEFORE the barrier (hyper-parameters and the
    // running beta powers were written by earlier launches), so only one L2 round trip follows it
    const float h_scale = hyper[6], max_norm = hyper[5];
    const double* cur = reinterpret_cast<const double*>(sync + 16) + ((seq - 1u) & 1u) * 8;
    const double b1d = (double)hyper[1], b2d = (double)hyper[2];
    const double c0 = cur[0], c1 = cur[1], c2 = cur[2], c3 = cur[3], c4 = cur[4];
    // Grid barrier without fences or atomics, two hops: block b publishes {partial sum, launch number} as ONE 8-byte
    // store (the value arrives with its tag); block 0's first warp polls the slots of all blocks, sums them in a fixed
    // order and publishes {sum of squares, launch number} the same way; every other block polls that single slot.  (All
    // blocks polling all slots would put 20 k loads on ten cache lines; the fence + ticket form costs two MEMBAR.ALL.GPU
    // and an atomic round trip on the critical path of every step.)
    uint2* slots = reinterpret_cast<uint2*>(sync + 256);   // slots of the blocks at byte 1024 of the workspace; [148] = the total
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
      partial[blockIdx.x] = t;   // (kept for inspection)
      asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(slots + blockIdx.x), "r"(__float_as_uint(t)), "r"(seq) : "memory");
    }
    VB_TL(tl_tail, 3);
    float totf;
    if (blockIdx.x == 0) {
      // lane l takes slots l, l + 32, ... (loads issued together), then a fixed shuffle tree (deterministic)
      float pv[5];
      unsigned int spins = 0;
      bool all = false;
      while (!all) {
        all = true;
        uint2 sv2[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          sv2[q] = make_uint2(0u, seq);
          if (threadIdx.x + 32 * q < gridDim.x)
            asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(sv2[q].x), "=r"(sv2[q].y) : "l"(slots + threadIdx.x + 32 * q) : "memory");
        }
#pragma unroll
        for (int q = 0; q < 5; ++q) { pv[q] = __uint_as_float(sv2[q].x); if (sv2[q].y != seq) all = false; }
        if (!all && ++spins > (1u << 28)) __trap();   // a mis-sized grid traps instead of hanging the device
      }
      double tot = 0.0;
#pragma unroll
      for (int q = 0; q < 5; ++q) tot += (double)pv[q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
      tot += (double)extra_sq;
      totf = (float)sqrt(tot);
      if (threadIdx.x == 0)
        asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(slots + TAIL_MAX_BLOCKS), "r"(__float_as_uint(totf)), "r"(seq) : "memory");
    } else {
      uint2 v = make_uint2(0u, 0u);
      if (threadIdx.x == 0) {
        unsigned int spins = 0;
        for (;;) {
          asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(slots + TAIL_MAX_BLOCKS) : "memory");
          if (v.y == seq) break;
          if (++spins > (1u << 28)) __trap();
        }
      }
      totf = __uint_as_float(__shfl_sync(0xffffffffu, v.x, 0));
    }
    if (threadIdx.x == 0) {
      const float norm = totf * fabsf(h_scale);
      float coef = 1.f;
      if (max_norm > 0.f) coef = fminf(1.f, max_norm / (norm + 1e-6f));
      const float step = step_prev + 1.f;
      // beta^step: a double pow() is ~2 us of dependent FP64; the running products {beta1^t, beta2^t, t, beta1, beta2} are
      // kept next to the barrier words (two copies, by launch parity, so block 0 can write while the others read) and
      // only recomputed when the step counter was changed behind our back (restore / load_state)
      double p1, p2;
      if (c2 == (double)step_prev && c3 == b1d && c4 == b2d && step_prev > 0.f) {
        p1 = c0 * b1d; p2 = c1 * b2d;
      } else {
        p1 = pow(b1d, (double)step); p2 = pow(b2d, (double)step);
      }
      red[0] = coef;
      red[1] = (float)(1.0 - p1);
      red[2] = (float)(1.0 - p2);
      if (blockIdx.x == 0) {
        double* nxt = reinterpret_cast<double*>(sync + 16) + (seq & 1u) * 8;
        nxt[0] = p1; nxt[1] = p2; nxt[2] = (double)step; nxt[3] = b1d; nxt[4] = b2d;
        state[0] = step;
        state[1] = norm;
        state[2] = coef;
        state[3] = red[1];
        state[4] = red[2];
        state[5] = 0.f;   // consumed
        if (rng) rng[1] += 1ull;
        sync[2] = seq;
        for (int gi = 0; gi < GS.n; ++gi) GS.done[gi] = 0u;   // every block is past its waits (grid barrier above)
      }
    }
  }
  __syncthreads();
  if (GS.n > 0) pdl_wait();   // the backward kernel has signalled everything; this makes its completion formal before the
                              // parameters it was reading are rewritten
  VB_TL(tl_tail, 4);
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
  const float gmul = hyper[6] * red[0];
  const float bc1 = red[1], bc2 = red[2];
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = 1.f / sqrtf(bc2);
  const float decay = 1.f - lr * wd;
  k = 0;
  for (size_t e = e0; e < n4; e += estride, ++k) {
    const size_t i = 4 * e + lane4;
    const float gr = (k < TAIL_KEEP ? keep[k] : __ldcg(&g[i])) * gmul;
    float pa = p[i] * decay, ma = m[i], va = v[i];
    ma = fmaf(1.f - b1, gr - ma, ma);            // exp_avg.lerp_(grad, 1 - beta1)
    va = fmaf(1.f - b2, gr * gr, b2 * va);       // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
    const float denom = sqrtf(va) * inv_sqrt_bc2 + eps;
    pa -= step_size * (ma / denom);
    p[i] = pa; m[i] = ma; v[i] = va;
    if (shadow) shadow[i] = __float2bfloat16_rn(pa);
  }
  VB_TL(tl_tail, 5);
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

extern "C" int vitb200_sumsq_accum(const float* g, size_t n, float* acc, void* ws, void* stream) {
  if (!g || !acc || !ws) return VITB200_ERR_ARG;
  if (n % 4 != 0) return VITB200_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(g) & 15) != 0) return VITB200_ERR_ALIGN;
  unsigned int* counter = reinterpret_cast<unsigned int*>(ws);
  float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + 4096);
  vb_launch_pdl(sumsq_accum_kernel, dim3(op_grid(n)), dim3(OP_THREADS), 0, (cudaStream_t)stream, g, n / 4, acc, partial, counter);
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

// block size / grid of the tail kernel: <= one CTA per SM (the in-kernel grid barrier needs co-residency), and, when the
// arena is small enough, exactly one pass (no block does two rounds while the others wait at the barrier)
static inline void tail_launch_shape(size_t n4, int* threads, int* grid) {
  size_t per = (n4 + TAIL_MAX_BLOCKS - 1) / TAIL_MAX_BLOCKS;   // float4 elements per block for one pass
  size_t t = ((per * 4 + 31) / 32) * 32;
  if (t < 128) t = 128;
  if (t > TAIL_MAX_THREADS) t = TAIL_MAX_THREADS;
  size_t g = (n4 + t / 4 - 1) / (t / 4);
  if (g > TAIL_MAX_BLOCKS) g = TAIL_MAX_BLOCKS;
  if (g < 1) g = 1;
  *threads = (int)t;
  *grid = (int)g;
}

VB_TL_EXPORT(vitb200_tl_tail, vb::tl_tail)

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
  int threads, grid;
  tail_launch_shape(n4, &threads, &grid);
  unsigned int* sync = reinterpret_cast<unsigned int*>(ws);   // {ticket, -, launch count, ..., beta powers}: zeroed by the caller once
  float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + 4096);
  vb_launch_pdl(clip_adamw_fused_kernel, dim3((unsigned)grid), dim3((unsigned)threads), 0, (cudaStream_t)stream, p, g, m, v,
                (bf16*)shadow, n4, hyper, state, rng, gpart, slots, stride / 4, red_start / 4, red_end / 4, partial, sync,
                PeerX{nullptr, 0, 1, 0}, GradStream{nullptr, 0u, 0, {}, {}});
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

extern "C" int vitb200_clip_adamw_fused_streamed(float* p, float* g, float* m, float* v, void* shadow, size_t n, const float* hyper,
                                                 float* state, uint64_t* rng, const float* gpart, int slots, size_t stride,
                                                 size_t red_start, size_t red_end, void* ws, const vitb200_grad_stream* gs,
                                                 void* const* peer_bufs, int rank, int world, void* stream) {
  if (!p || !g || !m || !v || !hyper || !state || !ws || !gs || !gs->done) return VITB200_ERR_ARG;
  if (gs->n_groups < 1 || gs->n_groups > VITB200_MAX_GROUPS || gs->expect == 0 || slots <= 0) return VITB200_ERR_ARG;
  if (peer_bufs ? (world < 2 || world > PEER_MAX_WORLD || rank < 0 || rank >= world) : (world != 1)) return VITB200_ERR_ARG;
  if (n % 4 != 0) return VITB200_ERR_SHAPE;
  if (!gpart || (stride | red_start | red_end) % 4 != 0 || red_end < red_start || red_end > n) return VITB200_ERR_ARG;
  if (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
        reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(gpart)) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(shadow) & 7) != 0)
    return VITB200_ERR_ALIGN;
  GradStream GS{gs->done, gs->expect, gs->n_groups, {}, {}};
  size_t covered = 0;
  for (int i = 0; i < gs->n_groups; ++i) {   // ascending, contiguous, float4-aligned, covering [0, n)
    if (gs->lo[i] != covered || gs->hi[i] < gs->lo[i] || (gs->lo[i] | gs->hi[i]) % 4 != 0) return VITB200_ERR_ARG;
    covered = gs->hi[i];
    GS.lo4[i] = gs->lo[i] / 4; GS.hi4[i] = gs->hi[i] / 4;
  }
  if (covered != n) return VITB200_ERR_ARG;
  const size_t n4 = n / 4;
  int threads, grid;
  tail_launch_shape(n4, &threads, &grid);
  unsigned int* sync = reinterpret_cast<unsigned int*>(ws);
  float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + 4096);
  vb_launch_pdl(clip_adamw_fused_kernel, dim3((unsigned)grid), dim3((unsigned)threads), 0, (cudaStream_t)stream, p, g, m, v,
                (bf16*)shadow, n4, hyper, state, rng, gpart, slots, stride / 4, red_start / 4, red_end / 4, partial, sync,
                peer_bufs ? PeerX{peer_bufs, rank, world, n} : PeerX{nullptr, 0, 1, 0}, GS);
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

// ---- data-parallel variant: the gradient all-reduce runs inside the kernel over peer memory ---------------------
extern "C" size_t vitb200_peer_buffer_bytes(size_t n) { return PEER_HEADER + 2 * (size_t)PEER_MAX_WORLD * n * sizeof(uint2); }

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
  if (world < 2 || world > PEER_MAX_WORLD || rank < 0 || rank >= world) return VITB200_ERR_ARG;
  if (n % 4 != 0) return VITB200_ERR_SHAPE;
  if (slots > 0 && (!gpart || (stride | red_start | red_end) % 4 != 0 || red_end < red_start || red_end > n)) return VITB200_ERR_ARG;
  if (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
        reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(gpart)) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(shadow) & 7) != 0)
    return VITB200_ERR_ALIGN;
  const size_t n4 = n / 4;
  int threads, grid;
  tail_launch_shape(n4, &threads, &grid);
  unsigned int* sync = reinterpret_cast<unsigned int*>(ws);
  float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + 4096);
  vb_launch_pdl(clip_adamw_fused_kernel, dim3((unsigned)grid), dim3((unsigned)threads), 0, (cudaStream_t)stream, p, g, m, v,
                (bf16*)shadow, n4, hyper, state, rng, gpart, slots, stride / 4, red_start / 4, red_end / 4, partial, sync,
                PeerX{peer_bufs, rank, world, n}, GradStream{nullptr, 0u, 0, {}, {}});
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
