// Whole-network backward kernel (bf16, H = 32, 2 heads of 16, T <= 129).  See include/vit_b200.h and mega_common.cuh.
//
// One CTA (or CTA pair: one attention head each, everything else computed by both) owns ONE sample and walks the network
// backwards with the gradient of the residual row in registers (8 columns per thread):
//   head + final LayerNorm of the CLS row (warp 16)
//   per layer, top down
//     upper : dz -> dropout' -> [W2 dgrad / wgrad] -> gelu' -> [W1 dgrad / wgrad] -> LN2 backward (+residual) -> dh
//             -> dropout' -> [Wo dgrad / wgrad] -> dctx tile (the dO operand of attention backward)
//     attn  : S = Q K^T, dP = dO V^T (UMMA) -> P, dS in registers -> dS, P~ tiles -> dQ = dS K, dK = dS^T Q, dV = P~^T dO
//             -> dqkv tile (shared memory; in a CTA pair each CTA writes its head's columns into both tiles)
//     lower : [Wqkv dgrad / wgrad] -> LN1 backward (+ dh) -> dz of the layer below
//   embedding : dropout'(dz0), dWp = dtok^T patches, dbp, dcls, dpos
// Saved activations arrive by TMA into buffers that are re-used phase by phase (loads are issued as soon as a buffer's
// last reader -- an MMA or an epilogue -- is done); weight tiles of the layer below are staged the same way.
// dgrad and wgrad run on tcgen05 from the SAME shared-memory tile (K-major view / MN-major view, see fused_bwd.cu).
// Parameter gradients of the sample leave as one partial set gpart[sample][n_opt]; the optimizer kernel sums the
// samples in order (deterministic, no atomics).  Warp 16 carries the 129th token with FMAs and contributes rank-1
// terms to the weight gradients.
#include "mega_common.cuh"

namespace vb {

// shared-memory plan (byte offsets from the 1 KB-aligned base)
constexpr uint32_t MB_R1 = 0;                    // 96 KB  upper: m @0 (32K), a @32K (32K), u2 @64K, hmid @80K;  gradient staging
                                                 //        attention: dS @0 (2 x 16K), P~ @32K (2 x 16K), dqkv tile @64K (2 x 16K)
                                                 //        lower: u @0, z @16K, dqkv tile @64K
constexpr uint32_t MB_CTX = 98304;               // 16 KB  ctx (load) -> dctx = dO tile
constexpr uint32_t MB_R3 = MB_CTX + 16384;       // 48 KB  upper: ddelta @0 (16K), da @16K (32K);  attention: q|k @0, v @18K
                                                 //        embedding: dtok @0, patches @16K
constexpr uint32_t MB_W2 = MB_R3 + 49152;        // 8 KB
constexpr uint32_t MB_W1 = MB_W2 + 8192;         // 16 KB
constexpr uint32_t MB_WO = MB_W1 + 16384;        // 4 KB
constexpr uint32_t MB_WQ = MB_WO + 4096;         // 12 KB
constexpr uint32_t MB_BAR = MB_WQ + 12288;       // barriers, TMEM slot
constexpr uint32_t MB_EX = MB_BAR + 256;         // 4 KB  float2 [4][128] LayerNorm-backward sums | float [4][128] D_i partials
constexpr uint32_t MB_RED = MB_EX + 4096;        // 4 KB  float [2][512] bias-gradient partials
constexpr uint32_t MB_SIDE = MB_RED + 4096;      // 8 KB  side-row vectors (floats), see SV_*
constexpr uint32_t MB_PRM = MB_SIDE + 8192 + 1024;   // LayerNorm gammas: [layers][2][32] + final [32]
// The side row (token 128) is carried by FOUR warps (128 lanes, sid = lane index in the group): 128-wide vectors have one
// element per lane, 32-wide vectors are replicated in every warp (lane = column), contractions are split over the warps
// and their partial sums are combined through shared memory in warp order (deterministic).
constexpr int MB_SIDE_WARPS = 4;
constexpr int MB_THREADS = MG_MAIN + 32 * MB_SIDE_WARPS;   // 640
constexpr uint32_t MB_QKV_BLK = 18432;
// side-row vectors (float offsets into MB_SIDE)
constexpr int SV_DD2 = 0, SV_M = 32, SV_DA = 160, SV_U2 = 288, SV_DD1 = 320, SV_CTX = 352, SV_DQKV = 384, SV_U = 480,
              SV_XDS = 512, SV_XPT = 672, SV_XQ = 832, SV_XDO = 848, SV_DZC = 864, SV_LN = 896, SV_LN1 = 960,
              SV_PART = 1024 /* [4][32] */, SV_DQP = 1152 /* [4][16] */, SV_DQC = 1216 /* [32] dq of the CLS row (CLS-only top layer) */;
// per-row terms of the side KEY (token 128): dS[i, 128] and P~[i, 128] of the 128 tensor-core query rows
constexpr uint32_t MB_K128 = MB_SIDE + 8192;     // float [2][128]
// TMEM columns (phases re-use them)
constexpr uint32_t UB_W2 = 0, UB_WO = 128, UB_DM = 160, UB_DU2 = 288, UB_DCTX = 320, UB_W1 = 352;   // upper
constexpr uint32_t AB_S = 0, AB_DP = 160, AB_DQ = 320, AB_DK = 352, AB_DV = 384;   // attention
constexpr uint32_t LB_WQ = 0, LB_DU = 64;       // lower
constexpr uint32_t EB_WP = 0;                    // embedding

VB_TL_DECL(tl_mega_bwd)

struct MegaBwdMaps { CUtensorMap wq, wo, w1, w2, z, hmid, u, u2, qkv, ctx, a, m; };

__device__ __forceinline__ uint32_t mb_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mb_peer_addr(const void* p, uint32_t peer) {
  uint32_t a;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(smem_u32(p)), "r"(peer));
  return a;
}
__device__ __forceinline__ void mb_st_peer16(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void mb_st_peer4(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// DSMEM pushes that signal the receiver's own mbarrier (see mega_fwd.cu)
__device__ __forceinline__ void mb_st_async16(uint32_t addr, uint4 v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void mb_st_async4(uint32_t addr, float v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(addr), "r"(__float_as_uint(v)),
               "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void mb_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
// write-after-read hand-off ("my tile may be overwritten"): nothing is published, so no release fence (MEMBAR.ALL.GPU)
__device__ __forceinline__ void mb_cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void mb_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void mb_bar_side() { asm volatile("bar.sync 2, 128;" ::: "memory"); }   // the 4 side-row warps

// sum over rows [rb, rb + 8) of column `col` (< 64) of a swizzled bf16 block (rows in order => deterministic)
__device__ __forceinline__ float mb_colsum8(const uint8_t* blk, int col, int rb) {
  float s = 0.f;
#pragma unroll
  for (int r = rb; r < rb + 8; ++r) s += __bfloat162float(*mg_elem(blk, r, col));
  return s;
}
// LayerNorm backward of a row spread over 4 threads: the two row sums are exchanged through shared memory
__device__ __forceinline__ void mb_ln_bwd_sums(const float (&du)[MG_HC], const float (&xh)[MG_HC], const float* gamma, int hc0,
                                               float2* s_ex, int r, int cg, float (&g)[MG_HC], float& c1, float& c2) {
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < MG_HC; ++j) {
    g[j] = du[j] * gamma[hc0 + j];
    s1 += g[j];
    s2 = fmaf(g[j], xh[j], s2);
  }
  s_ex[cg * 128 + r] = make_float2(s1, s2);
  mg_bar_main();
  float t1 = 0.f, t2 = 0.f;
#pragma unroll
  for (int q = 0; q < MG_CG; ++q) { const float2 p = s_ex[q * 128 + r]; t1 += p.x; t2 += p.y; }
  c1 = t1 * (1.f / MG_H);
  c2 = t2 * (1.f / MG_H);
}
// red[e * 128 + row] (e < NE) -> out(e) = sum over the 128 rows in a fixed order (8 threads x 16 rows, shuffle tree).
// Called by the 512 main threads after a barrier; NE * 8 <= 512.
template <typename F>
__device__ __forceinline__ void mb_reduce_rows(const float* red, int NE, F&& out) {
  const int e = threadIdx.x >> 3, part = threadIdx.x & 7;
  float s = 0.f;
  if (e < NE) {
#pragma unroll
    for (int k = 0; k < 16; ++k) s += red[e * 128 + part * 16 + k];
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (e < NE && part == 0) out(e, s);
}
// d loss / d logits of one sample (specvit.py:81-89; the gradient of the bf16 logits is bf16 under autocast)
__device__ __forceinline__ void mb_dlogits(const float* lg, const void* labels, int b, int B, int C, int kind, float g, float (&dl)[4]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) dl[c] = 0.f;
  if (kind == VITB200_LOSS_GIVEN) {
#pragma unroll
    for (int c = 0; c < 4; ++c) if (c < C) dl[c] = bf16_round(reinterpret_cast<const float*>(labels)[(size_t)b * C + c]);
    return;
  }
  if (kind == VITB200_LOSS_CE) {
    const long long y = reinterpret_cast<const long long*>(labels)[b];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < 4; ++c) if (c < C) mx = fmaxf(mx, lg[c]);
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) if (c < C) se += expf(lg[c] - mx);
#pragma unroll
    for (int c = 0; c < 4; ++c) if (c < C) dl[c] = bf16_round((expf(lg[c] - mx) / se - (c == (int)y ? 1.f : 0.f)) / (float)B * g);
    return;
  }
  const float n = (float)B * (float)C;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (c < C) {
      const float e = lg[c] - reinterpret_cast<const float*>(labels)[(size_t)b * C + c];
      const float d = kind == VITB200_LOSS_L1 ? ((e > 0.f) - (e < 0.f)) / n : 2.f * e / n;
      dl[c] = bf16_round(d * g);
    }
  }
}

// "this CTA's partial gradients of group g are written": one thread, after a CTA barrier that follows the stores (the
// release is cumulative over what the thread observed through the barrier).  The optimizer kernel of the step acquires
// the counter and starts summing that group's partials while the layers below are still running.
__device__ __forceinline__ void mb_signal(unsigned int* done, int g) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(done + g) : "memory");
}

__global__ void __launch_bounds__(MB_THREADS, 1)
mega_bwd_kernel(const __grid_constant__ MegaBwdMaps TM, const vitb200_mega_bwd_args PA) {
  constexpr int H = MG_H, I = MG_I, HC = MG_HC, D = MG_D;
  const vitb200_mega_fwd_args& P = PA.f;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t *sM = base + MB_R1, *sAct = base + MB_R1 + 32768, *sU2 = base + MB_R1 + 65536, *sHm = base + MB_R1 + 81920;
  uint8_t *sDS = base + MB_R1, *sPT = base + MB_R1 + 32768, *sDQ = base + MB_R1 + 65536;
  uint8_t* sCtx = base + MB_CTX;
  uint8_t *sD = base + MB_R3, *sDA = base + MB_R3 + 16384, *sQ0 = base + MB_R3, *sQ1 = base + MB_R3 + MB_QKV_BLK;
  uint8_t *sU = base + MB_R1, *sZ = base + MB_R1 + 16384, *sX = base + MB_R3 + 16384;   // u / z of the lower stage: R1 @0 / @16K
  uint8_t *sW2 = base + MB_W2, *sW1 = base + MB_W1, *sWo = base + MB_WO, *sWq = base + MB_WQ;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + MB_BAR);
  uint64_t *b_w2 = bars, *b_w1 = bars + 1, *b_wo = bars + 2, *b_wq = bars + 3, *b_up = bars + 4, *b_ctx = bars + 5,
           *b_qkv = bars + 6, *b_low = bars + 7, *b_mma = bars + 8, *b_wg = bars + 9, *b_a = bars + 10, *b_x = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  float2* s_ex = reinterpret_cast<float2*>(base + MB_EX);
  float* s_dp = reinterpret_cast<float*>(base + MB_EX);      // [4][128] D_i partials (attention phase)
  float* bsum = reinterpret_cast<float*>(base + MB_RED);     // [2][512]
  float* sv = reinterpret_cast<float*>(base + MB_SIDE);
  float* k128 = reinterpret_cast<float*>(base + MB_K128);     // [0..127]: dS[i,128] / scale, [128..255]: P~[i,128]
  float* s_gam = reinterpret_cast<float*>(base + MB_PRM);    // [layers][2][32]: ln1 gamma, ln2 gamma ; then final gamma [32]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  VB_TL(tl_mega_bwd, 0);
  const int csz = P.cluster == 2 ? 2 : 1;
  const uint32_t crank = csz == 2 ? mb_cluster_rank() : 0u;
  const bool lead = crank == 0;
  const int hd_lo = csz == 2 ? (int)crank : 0, hd_hi = csz == 2 ? (int)crank + 1 : MG_NH;
  const bool is_side = warp >= 16;
  const int sw = warp - 16, sid = tid - MG_MAIN;   // side group: warp / lane index inside the group
  const bool s0 = sw == 0;                         // the side warp that publishes replicated vectors
  float* sp = reinterpret_cast<float*>(base + MB_SIDE) + SV_PART;
  float* sdqp = reinterpret_cast<float*>(base + MB_SIDE) + SV_DQP;
  const int r = ((warp & 3) << 5) | lane;
  const int cg = warp >> 2, hc0 = cg * HC;
  const int T = P.Np + 1, L = P.layers, B = P.B, C = P.C;
  const int Tm = T < 128 ? T : 128;
  const bool has_side = T > 128;
  const int KP = (T + 15) & ~15, nch = KP >> 4;
  const int KM = KP < 128 ? KP : 128;            // keys covered by the dS / P~ tiles (key 128 = the side row: FMAs)
  const size_t M = (size_t)B * T;
  const bool valid = !is_side && r < Tm;
  const int b = (int)blockIdx.x / csz;          // this CTA's sample (grid = B * csz)
  const size_t grow = (size_t)b * T + (is_side ? 128 : (valid ? r : 0));
  float* gp = PA.gpart + (size_t)b * PA.n_opt;
  const bf16* shadow = reinterpret_cast<const bf16*>(P.shadow);

  if (tid == 0) {
    tma_prefetch_desc(&TM.wq); tma_prefetch_desc(&TM.wo); tma_prefetch_desc(&TM.w1); tma_prefetch_desc(&TM.w2);
    tma_prefetch_desc(&TM.z); tma_prefetch_desc(&TM.hmid); tma_prefetch_desc(&TM.u); tma_prefetch_desc(&TM.u2);
    tma_prefetch_desc(&TM.qkv); tma_prefetch_desc(&TM.ctx); tma_prefetch_desc(&TM.a); tma_prefetch_desc(&TM.m);
    for (int i = 0; i < 12; ++i) mbar_init(bars + i, 1);
    fence_barrier_init();
  }
  for (int j = tid; j < L * 64 + 32; j += MB_THREADS) {   // gammas are parameters: the optimizer ran two kernels ago
    if (j < L * 64) {
      const int l = j >> 6, e = j & 63;
      const float* lp = P.params + P.off_layer0 + (size_t)l * P.layer_stride;
      s_gam[j] = e < 32 ? lp[P.o_ln1g + e] : lp[P.o_ln2g + (e - 32)];
    } else {
      s_gam[j] = P.params[P.off_lnfg + (j - L * 64)];
    }
  }
  __syncthreads();
  auto load_w2 = [&](int l) { mbar_expect_tx(b_w2, 8192); tma_load_3d(sW2, &TM.w2, b_w2, 0, 0, l); tma_load_3d(sW2 + 4096, &TM.w2, b_w2, 64, 0, l); };
  auto load_w1 = [&](int l) { mbar_expect_tx(b_w1, 16384); tma_load_3d(sW1, &TM.w1, b_w1, 0, 0, l); };
  auto load_wo = [&](int l) { mbar_expect_tx(b_wo, 4096); tma_load_3d(sWo, &TM.wo, b_wo, 0, 0, l); };
  auto load_wq = [&](int l) { mbar_expect_tx(b_wq, 12288); tma_load_3d(sWq, &TM.wq, b_wq, 0, 0, l); };
  auto load_ctx = [&](int l) { mbar_expect_tx(b_ctx, 16384); tma_load_4d(sCtx, &TM.ctx, b_ctx, 0, 0, b, l); };
  // R1 is re-used phase by phase, so the saved activations of a layer arrive in two batches:
  //   a (pre-GELU, R1 @32K)            : as soon as the attention MMAs of the layer above are done        -> b_a
  //   m (@0), u2 (@64K), hmid (@80K)   : once the lower stage above has consumed its u / z / dqkv tiles   -> b_up
  auto load_a = [&](int l) {
    mbar_expect_tx(b_a, 32768);
    tma_load_4d(sAct, &TM.a, b_a, 0, 0, b, l); tma_load_4d(sAct + 16384, &TM.a, b_a, 64, 0, b, l);
  };
  auto load_upper = [&](int l) {
    mbar_expect_tx(b_up, 32768 + 16384 + 16384);
    tma_load_4d(sM, &TM.m, b_up, 0, 0, b, l); tma_load_4d(sM + 16384, &TM.m, b_up, 64, 0, b, l);
    tma_load_4d(sU2, &TM.u2, b_up, 0, 0, b, l);
    tma_load_4d(sHm, &TM.hmid, b_up, 0, 0, b, l);
  };
  if (tid == 0) { load_w2(L - 1); load_w1(L - 1); load_wo(L - 1); load_wq(L - 1); }   // weights: not written since the optimizer
  VB_TL(tl_mega_bwd, 1);
  pdl_wait();     // everything below reads what the forward kernel produced
  pdl_trigger();
  VB_TL(tl_mega_bwd, 2);
  const bool cls_only = P.cls_only != 0;
  if (P.defer_loss && blockIdx.x == 0 && warp == 1 && P.labels) {   // the forward kernel left the loss as per-CTA terms
    const float t = mg_loss_sum(reinterpret_cast<const float*>(reinterpret_cast<const char*>(P.ws) + 256), gridDim.x, lane);
    if (lane == 0) {
      const float v = t / (P.loss_kind == VITB200_LOSS_CE ? (float)B : (float)B * (float)C);
      P.loss[0] = v;
      if (P.loss_log) P.loss_log[P.rows ? (size_t)(P.rng[1] - P.rows_base[0]) : 0] = v;
    }
  }
  if (tid == 0 && !cls_only) { load_a(L - 1); load_upper(L - 1); load_ctx(L - 1); }   // (a CLS-only top layer loads no tiles)
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t my_tmem = tmem + ((uint32_t)((warp & 3) * 32) << 16);

  const uint32_t aD = smem_u32(sD), aDA = smem_u32(sDA), aM = smem_u32(sM), aU2 = smem_u32(sU2), aCtx = smem_u32(sCtx);
  const uint32_t aQ0 = smem_u32(sQ0), aQ1 = smem_u32(sQ1), aDS = smem_u32(sDS), aPT = smem_u32(sPT), aDQ = smem_u32(sDQ);
  const MgOp D_k{aD, 16, 16384, 0}, D_mn{aD, 16384, 0, 1};          // MN group 1 of the ddelta tile aliases the da tile: the
                                                                    // product rows 64..127 are never read (H = 32 weight rows)
  const MgOp DA_k{aDA, 16, 16384, 0}, DA_mn{aDA, 16384, 0, 1};
  const MgOp M_mn{aM, 16384, 0, 1}, U2_mn{aU2, 16384, 0, 1}, CTX_mn{aCtx, 16384, 0, 1};
  const MgOp W2_mn{smem_u32(sW2), 4096, 0, 1}, W1_mn{smem_u32(sW1), 16384, 0, 1}, WO_mn{smem_u32(sWo), 4096, 0, 1};
  const MgOp WQ_mn{smem_u32(sWq), 16384, 0, 1};
  const MgOp DQ_k{aDQ, 16, 16384, 0}, DQ_mn{aDQ, 16384, 0, 1}, U_mn{smem_u32(sU), 16384, 0, 1};
  const MgOp DS_k{aDS, 16, 16384, 0};

  const uint64_t seed = P.rng ? P.rng[0] : 0ull;
  const uint32_t step = P.rng ? (uint32_t)P.rng[1] : 0u;
  // dataset row of this sample (device-resident dataset mode, see vitb200_mega_fwd_args.rows)
  const size_t bsrc = P.rows ? (size_t)P.rows[(size_t)(P.rng[1] - P.rows_base[0]) * (size_t)B + b] : (size_t)b;
  const float scale = rsqrtf((float)D), sl2 = scale * MG_LOG2E;
  const int Tpad = attn_drop_tpad(T);
  uint32_t ph_mma = 0, ph_x = 0;

  // =============================== head + final LayerNorm of the CLS row (HF:455, specvit.py:78-89) ===============================
  float dz[HC];      // main: gradient of the residual row, columns hc0 .. hc0 + 7
  float dzs = 0.f;   // side: gradient of the residual row of token 128, column = lane
#pragma unroll
  for (int j = 0; j < HC; ++j) dz[j] = 0.f;
  if (is_side) {
    float lg[4], dl[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) lg[c] = c < C ? P.logits[(size_t)b * C + c] : 0.f;
    const float gl = PA.gloss ? PA.gloss[0] : 1.f;
    mb_dlogits(lg, PA.labels, (int)bsrc, B, C, PA.loss_kind, gl, dl);
    const float s = __bfloat162float(reinterpret_cast<const bf16*>(P.s_cls)[(size_t)b * H + lane]);
    float ds = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) if (c < C) ds = fmaf(dl[c], __bfloat162float(shadow[P.off_wh + c * H + lane]), ds);
    ds = bf16_round(ds);
    const float mu = P.stats[(size_t)(4 * L) * M + b], rs = P.stats[(size_t)(4 * L + 1) * M + b];
    const float xh = (P.z[((size_t)L * M + (size_t)b * T) * H + lane] - mu) * rs;
    const float gg = ds * s_gam[L * 64 + lane];
    const float c1 = mg_wsum(gg) * (1.f / H), c2 = mg_wsum(gg * xh) * (1.f / H);
    if (s0) sv[SV_DZC + lane] = rs * (gg - c1 - xh * c2);
    if (lead && s0) {
      gp[P.off_lnfg + lane] = ds * xh;
      gp[P.off_lnfb + lane] = ds;
#pragma unroll
      for (int c = 0; c < 4; ++c) if (c < C) gp[P.off_wh + c * H + lane] = dl[c] * s;
      if (lane < C) gp[P.off_bh + lane] = lane == 0 ? dl[0] : lane == 1 ? dl[1] : lane == 2 ? dl[2] : dl[3];
    }
  }
  __syncthreads();
  if (!is_side && r == 0 && !cls_only) {
#pragma unroll
    for (int j = 0; j < HC; ++j) dz[j] = sv[SV_DZC + hc0 + j];
  }

  // =============================== layers, top down ===============================
#ifdef VB_TL_TOP
  const int tl_l = L - 1;                    // (debug build, -DVB_TL_TOP: stamp the top layer -- CLS-only in a training step)
#else
  const int tl_l = L >= 2 ? L - 2 : L - 1;   // (debug build: the layer whose phases are time-stamped -- a steady-state one)
#endif
  (void)tl_l;
  for (int l = L - 1; l >= 0; --l) {
    const uint32_t par = (uint32_t)((L - 1 - l) & 1);   // every per-layer barrier completes once per layer
    // CLS-only top layer (see mega_fwd.cu): only the CLS row carries a gradient into the last layer (specvit.py:78), so its
    // MLP / out-proj / attention backward is ONE row: the side group runs it with FMAs (token 0 instead of token 128), the
    // weight gradients of that half are rank-1, dK / dV are rank-1 in the CLS query, and the tensor-core rows join again
    // at the QKV projection (lower stage).
    const bool ct = cls_only && l == L - 1;
    const bool s_on = has_side || ct;
    const int stok = ct ? 0 : 128;
    const size_t sgrow = (size_t)b * T + stok;
    const uint32_t parU = (uint32_t)((L - 1 - l - (cls_only ? 1 : 0)) & 1);   // b_up / b_a / b_ctx: first used one layer later
    const float* g1 = s_gam + l * 64;
    const float* g2 = s_gam + l * 64 + 32;
    const DropCtx dc_mlp = make_drop(P.p_hidden, seed, step, VITB200_SITE_MLP(l));
    const DropCtx dc_proj = make_drop(P.p_hidden, seed, step, VITB200_SITE_PROJ(l));
    const DropCtx dc_att = make_drop(P.p_attn, seed, step, VITB200_SITE_ATTN(l));
    // bytes the peer pushes into this CTA per layer: its head's dQ | dK | dV pieces (6 x 16 B x 128 rows) and, with a side
    // row, that row's 48 values
    if (csz == 2 && tid == 0) mbar_expect_tx(b_x, 12288u + (has_side ? 192u : 0u));
    const uint32_t x_peer = csz == 2 ? mb_peer_addr(b_x, crank ^ 1u) : 0u;
    float acc_b2 = 0.f, acc_bo = 0.f;
    float gam2[HC], bet2[HC];   // LN2 gamma / beta gradient terms of this thread's row
    float mu2 = 0.f, rs2 = 0.f;
    if (!is_side) { mu2 = P.stats[(size_t)(4 * l + 2) * M + grow]; rs2 = P.stats[(size_t)(4 * l + 3) * M + grow]; }
#pragma unroll
    for (int j = 0; j < HC; ++j) { gam2[j] = 0.f; bet2[j] = 0.f; }
    const float dzin = ct ? sv[SV_DZC + lane] : dzs;   // side chain: incoming gradient of its row (CLS: from the head)

    // ---------------- upper, stage 1: ddelta2 = dropout'(dz) -> sD ; dm = ddelta2 W2 ; dW2 += ddelta2^T m ----------------
    float dd2s = 0.f, dms = 0.f, das = 0.f, as_ = 0.f;   // side row (128-wide vectors: element sid)
    float mu2s = 0.f, rs2s = 0.f, hms = 0.f, u2s = 0.f, ctxs = 0.f;
    if (!is_side && !ct) {
      float kp[8], d2[HC];
      drop8(dc_mlp, (grow * H + hc0) >> 3, kp);
#pragma unroll
      for (int j = 0; j < HC; ++j) d2[j] = valid ? bf16_round(dz[j]) * kp[j] : 0.f;
      *reinterpret_cast<uint4*>(mg_chunk(sD, r, cg)) = mg_pack8(d2);
      *reinterpret_cast<uint4*>(mg_chunk(sD, r, 4 + cg)) = make_uint4(0u, 0u, 0u, 0u);   // columns 32..63: read by the MN-major view
    } else if (is_side && s_on) {
      // loads of the side row's saved activations (global, L2-resident): issued first, consumed along the chain
      const size_t lr = (size_t)l * M + sgrow;
      as_ = __bfloat162float(reinterpret_cast<const bf16*>(P.a)[lr * I + sid]);
      sv[SV_M + sid] = __bfloat162float(reinterpret_cast<const bf16*>(P.m)[lr * I + sid]);
      mu2s = P.stats[(size_t)(4 * l + 2) * M + sgrow]; rs2s = P.stats[(size_t)(4 * l + 3) * M + sgrow];
      hms = P.hmid[lr * H + lane];
      u2s = __bfloat162float(reinterpret_cast<const bf16*>(P.u2)[lr * H + lane]);
      ctxs = __bfloat162float(reinterpret_cast<const bf16*>(P.ctx)[lr * H + lane]);
      dd2s = bf16_round(dzin) * drop1(dc_mlp, sgrow * H + lane);
      if (s0) { sv[SV_U2 + lane] = u2s; sv[SV_CTX + lane] = ctxs; sv[SV_DD2 + lane] = dd2s; }
      mbar_wait(b_w2, par);
      // dm[i] = sum_h ddelta2[h] W2[h][i],  i = sid
      const uint8_t* wcol = sW2 + (sid >> 6) * 4096;
#pragma unroll 8
      for (int hh = 0; hh < H; ++hh)
        dms = fmaf(__shfl_sync(0xffffffffu, dd2s, hh), __bfloat162float(*mg_elem(wcol, hh, sid & 63)), dms);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (l == tl_l) VB_TL(tl_mega_bwd, 3);
    if (tid == 0 && !ct) {
      tc_fence_after();
      mbar_wait(b_w2, par);
      mg_issue(tmem + UB_DM, D_k, W2_mn, I, H / 16, false);      // dm[row, i] = sum_h ddelta2[row,h] W2[h,i]
      umma_commit(b_mma);                                        // the gelu' stage only needs dm
      mbar_wait(b_up, parU);                                     // (m, u2, hmid were requested a whole stage ago)
      tc_fence_after();
      mg_issue(tmem + UB_W2, D_mn, M_mn, I, 8, false);           // dW2[h, i] = sum_rows ddelta2[row,h] m[row,i]
    }                                                            // (completion is covered by the next commit)
    if (!is_side && !ct) {
      acc_b2 = mb_colsum8(sD, tid & 31, (tid >> 5) * 8);          // bias gradient partial: column tid & 31, rows 8 (tid >> 5) ..
      mbar_wait(b_up, parU);
      mbar_wait(b_a, parU);
      // a column of ones next to u2 (columns H..63 are TMA zero fill): the dW1 MMA then also emits db1
      if (cg == 0) *reinterpret_cast<uint32_t*>(mg_chunk(sU2, r, H / 8)) = valid ? 0x00003F80u : 0u;   // bf16 {1, 0}
      mbar_wait(b_mma, ph_mma); ph_mma ^= 1;
      tc_fence_after();
      // ---- da = dm * gelu'(a) -> sDA ----
      const int c0 = cg * 32;
      float v[32];
      tmem_ld_32x32(my_tmem + UB_DM + c0, v);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 av = *reinterpret_cast<const uint4*>(mg_swz(sAct, r, (c0 >> 3) + q));
        const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&av);
        float o[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(a2[e]);
          const float2 dm = __bfloat1622float2(__floats2bfloat162_rn(v[q * 8 + 2 * e], v[q * 8 + 2 * e + 1]));
          o[2 * e] = valid ? dm.x * gelu_grad_f(f.x) : 0.f;
          o[2 * e + 1] = valid ? dm.y * gelu_grad_f(f.y) : 0.f;
        }
        *reinterpret_cast<uint4*>(mg_swz(sDA, r, (c0 >> 3) + q)) = mg_pack8(o);
      }
    } else if (is_side && s_on) {
      das = bf16_round(bf16_round(dms) * gelu_grad_f(as_));
      sv[SV_DA + sid] = das;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (l == tl_l) VB_TL(tl_mega_bwd, 4);
    // ---------------- stage 2: du2 = da W1 ; dW1 (+ db1) += da^T [u2 | 1] ; LN2 backward -> dh ; ddelta1 -> sD ----------------
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(b_w1, par);
      if (!ct) {
        mg_issue(tmem + UB_DU2, DA_k, W1_mn, H, I / 16, false);    // du2[row, h] = sum_i da[row,i] W1[i,h]
        mg_issue(tmem + UB_W1, DA_mn, U2_mn, H + 16, 8, false);    // dW1[i, h] = sum_rows da[row,i] u2[row,h] ; column H: db1[i]
        umma_commit(b_mma);
      }
      if (l > 0) load_w2(l - 1);                                 // W2: its GEMM completed a stage ago, the side warp is past it
    }
    float dd1s = 0.f, dhs = 0.f, xh2s = 0.f;
    if (!is_side && !ct) {
      mbar_wait(b_mma, ph_mma); ph_mma ^= 1;
      tc_fence_after();
      float du[HC], xh[HC], g[HC];
      tmem_ld_32x8(my_tmem + UB_DU2 + hc0, du);
      {
        const float4 x0 = *mg_f32(sHm, r, hc0), x1 = *mg_f32(sHm, r, hc0 + 4);
        xh[0] = x0.x; xh[1] = x0.y; xh[2] = x0.z; xh[3] = x0.w; xh[4] = x1.x; xh[5] = x1.y; xh[6] = x1.z; xh[7] = x1.w;
      }
#pragma unroll
      for (int j = 0; j < HC; ++j) {
        xh[j] = (xh[j] - mu2) * rs2;
        du[j] = valid ? bf16_round(du[j]) : 0.f;
        gam2[j] = du[j] * xh[j];
        bet2[j] = du[j];
      }
      float c1, c2;
      mb_ln_bwd_sums(du, xh, g2, hc0, s_ex, r, cg, g, c1, c2);
#pragma unroll
      for (int j = 0; j < HC; ++j) dz[j] += rs2 * (g[j] - c1 - xh[j] * c2);   // dz now holds dh
      float kp[8], d1[HC];
      drop8(dc_proj, (grow * H + hc0) >> 3, kp);
#pragma unroll
      for (int j = 0; j < HC; ++j) d1[j] = valid ? bf16_round(dz[j]) * kp[j] : 0.f;
      *reinterpret_cast<uint4*>(mg_chunk(sD, r, cg)) = mg_pack8(d1);
    } else if (is_side && s_on) {
      mbar_wait(b_w1, par);
      // du2[h] = sum_i da[i] W1[i][h],  h = lane; warp w contracts i in [32 w, 32 w + 32)
      float part = 0.f;
#pragma unroll 8
      for (int ii = 0; ii < 32; ++ii)
        part = fmaf(__shfl_sync(0xffffffffu, das, ii), __bfloat162float(*mg_elem(sW1, 32 * sw + ii, lane)), part);
      sp[sw * 32 + lane] = part;
      mb_bar_side();
      const float du2 = bf16_round((sp[lane] + sp[32 + lane]) + (sp[64 + lane] + sp[96 + lane]));
      xh2s = (hms - mu2s) * rs2s;
      const float gg = du2 * g2[lane];
      const float c1 = mg_wsum(gg) * (1.f / H), c2 = mg_wsum(gg * xh2s) * (1.f / H);
      dhs = dzin + rs2s * (gg - c1 - xh2s * c2);
      dd1s = bf16_round(dhs) * drop1(dc_proj, sgrow * H + lane);
      if (s0) { sv[SV_DD1 + lane] = dd1s; sv[SV_LN + lane] = du2 * xh2s; sv[SV_LN + 32 + lane] = du2; }   // dgamma2 / dbeta2 terms
      if (ct && s0) sv[SV_DZC + lane] = dhs;   // CLS-only top layer: dh of the CLS row goes back to tile row 0 (lower stage)
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (l == tl_l) VB_TL(tl_mega_bwd, 5);
    // ---------------- stage 3: dctx = ddelta1 Wo ; dWo += ddelta1^T ctx ----------------
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(b_wo, par);
      if (!ct) {
        mbar_wait(b_ctx, parU);
        mg_issue(tmem + UB_DCTX, D_k, WO_mn, H, H / 16, false);    // dctx[row, k] = sum_n ddelta1[row,n] Wo[n,k]
        mg_issue(tmem + UB_WO, D_mn, CTX_mn, H, 8, false);         // dWo[n, k] = sum_rows ddelta1[row,n] ctx[row,k]
        umma_commit(b_mma);
      }
      if (l > 0) load_w1(l - 1);
    }
    if (l == tl_l) VB_TL(tl_mega_bwd, 16);
    float dctxs = 0.f;   // side: dctx row, column = lane
    if (!is_side && !ct) {
      acc_bo = mb_colsum8(sD, tid & 31, (tid >> 5) * 8);
      mbar_wait(b_ctx, parU);
      mbar_wait(b_mma, ph_mma); ph_mma ^= 1;
      tc_fence_after();
      // dctx -> bf16 -> the ctx tile (dead: its wgrad MMA is done), which becomes the dO operand of attention backward;
      // D_i = dO_i . O_i (flash-attention backward's row statistic): partial over this thread's 8 columns
      float v[HC], o8[8];
      tmem_ld_32x8(my_tmem + UB_DCTX + hc0, v);
      uint8_t* cp = mg_chunk(sCtx, r, cg);
      mg_unpack8(*reinterpret_cast<const uint4*>(cp), o8);
      float dpart = 0.f;
#pragma unroll
      for (int j = 0; j < HC; ++j) { v[j] = valid ? bf16_round(v[j]) : 0.f; dpart = fmaf(v[j], o8[j], dpart); }
      *reinterpret_cast<uint4*>(cp) = mg_pack8(v);
      s_dp[cg * 128 + r] = dpart;
    } else if (is_side && s_on) {
      mbar_wait(b_wo, par);
#pragma unroll 8
      for (int n = 0; n < H; ++n) {
        const float dv = __shfl_sync(0xffffffffu, dd1s, n);
        dctxs = fmaf(dv, __bfloat162float(*mg_elem(sWo, n, lane)), dctxs);
      }
      dctxs = bf16_round(dctxs);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (l == tl_l) VB_TL(tl_mega_bwd, 17);
    // CTA pair: from here on this CTA's dqkv tile (R1 @64K, until now u2 / hmid) and side-row dqkv vector may be written
    // by the peer (barrier B: arrive here, the peer waits right before it pushes)
    if (csz == 2) mb_cluster_arrive_relaxed();
    if (l == tl_l) VB_TL(tl_mega_bwd, 6);
    // ---------------- upper: parameter gradients of the layer's MLP / out-proj half -> gpart ----------------
    // The ddelta / da tiles (R3) are dead: the q|k|v rows of the layer are fetched while the gradients drain.
    if (tid == 0) {
      mbar_expect_tx(b_qkv, 32768);
      tma_load_4d(sQ0, &TM.qkv, b_qkv, 0, 0, b, l);
      tma_load_4d(sQ1, &TM.qkv, b_qkv, 64, 0, b, l);
      if (l > 0) load_wo(l - 1);
    }
    {
      const size_t lo = (size_t)P.off_layer0 + (size_t)l * P.layer_stride;
      tc_fence_after();
      // TMEM rows are parameter rows: the tiles go through shared memory (16-byte pieces, XOR-swizzled so that neither
      // side has bank conflicts) and leave as coalesced float4s.  R1 (m / a / u2 / hmid) is dead after the MMAs above.
      float4* t_w1 = reinterpret_cast<float4*>(sM);            // [128 rows][8 pieces]   dW1 (16 KB)
      float4* t_w2 = reinterpret_cast<float4*>(sM + 16384);    // [32 rows][32 pieces]   dW2 (16 KB)
      float4* t_wo = reinterpret_cast<float4*>(sAct);          // [32 rows][8 pieces]    dWo (4 KB)
      float* t_b1 = reinterpret_cast<float*>(sAct + 4096);     // [128]                  db1
      float* red = reinterpret_cast<float*>(sAct + 8192);      // [2 * H][128]           LN2 gamma / beta terms (32 KB)
      const bool mm = !ct;                                     // tensor-core terms exist (CLS-only top layer: rank-1 terms only)
      if (!is_side && mm) {
        if ((warp & 3) == 0) {  // dW2 / dWo rows h = lanes 0..31: column group cg of each
          float v[32];
          tmem_ld_32x32(my_tmem + UB_W2 + cg * 32, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) t_w2[r * 32 + ((cg * 8 + j) ^ (r & 31))] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          float w[HC];
          tmem_ld_32x8(my_tmem + UB_WO + hc0, w);
          t_wo[r * 8 + ((2 * cg) ^ (r & 7))] = make_float4(w[0], w[1], w[2], w[3]);
          t_wo[r * 8 + ((2 * cg + 1) ^ (r & 7))] = make_float4(w[4], w[5], w[6], w[7]);
        }
        {  // dW1 rows i = lanes 0..127, H columns (+ db1 in column H)
          float v[HC];
          tmem_ld_32x8(my_tmem + UB_W1 + hc0, v);
          t_w1[r * 8 + ((2 * cg) ^ (r & 7))] = make_float4(v[0], v[1], v[2], v[3]);
          t_w1[r * 8 + ((2 * cg + 1) ^ (r & 7))] = make_float4(v[4], v[5], v[6], v[7]);
          if (cg == 0) {
            float b8[8];
            tmem_ld_32x8(my_tmem + UB_W1 + H, b8);
            t_b1[r] = b8[0];
          }
        }
#pragma unroll
        for (int j = 0; j < HC; ++j) { red[(hc0 + j) * 128 + r] = gam2[j]; red[(H + hc0 + j) * 128 + r] = bet2[j]; }
        bsum[tid] = acc_b2; bsum[512 + tid] = acc_bo;
      }
      tc_fence_before();
      __syncthreads();
      if (!is_side && lead) {
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        // side-row terms (rank 1): dW[n][k] += dy[n] x[k]; zero vectors when the sample has no side row
        const float* sdd2 = sv + SV_DD2; const float* sm = sv + SV_M; const float* sda = sv + SV_DA;
        const float* su2 = sv + SV_U2; const float* sdd1 = sv + SV_DD1; const float* sctx = sv + SV_CTX;
        float4* o_w1 = reinterpret_cast<float4*>(gp + lo + P.o_w1);
        float4* o_w2 = reinterpret_cast<float4*>(gp + lo + P.o_w2);
        float4* o_wo = reinterpret_cast<float4*>(gp + lo + P.o_wo);
        for (int e = tid; e < 128 * 8; e += MG_MAIN) {
          const int rr = e >> 3, c = e & 7;
          float4 t = mm ? t_w1[rr * 8 + (c ^ (rr & 7))] : z4;
          if (s_on) { const float d = sda[rr]; t.x = fmaf(d, su2[4 * c], t.x); t.y = fmaf(d, su2[4 * c + 1], t.y); t.z = fmaf(d, su2[4 * c + 2], t.z); t.w = fmaf(d, su2[4 * c + 3], t.w); }
          o_w1[e] = t;
        }
        for (int e = tid; e < 32 * 32; e += MG_MAIN) {
          const int rr = e >> 5, c = e & 31;
          float4 t = mm ? t_w2[rr * 32 + (c ^ (rr & 31))] : z4;
          if (s_on) { const float d = sdd2[rr]; t.x = fmaf(d, sm[4 * c], t.x); t.y = fmaf(d, sm[4 * c + 1], t.y); t.z = fmaf(d, sm[4 * c + 2], t.z); t.w = fmaf(d, sm[4 * c + 3], t.w); }
          o_w2[e] = t;
        }
        for (int e = tid; e < 32 * 8; e += MG_MAIN) {
          const int rr = e >> 3, c = e & 7;
          float4 t = mm ? t_wo[rr * 8 + (c ^ (rr & 7))] : z4;
          if (s_on) { const float d = sdd1[rr]; t.x = fmaf(d, sctx[4 * c], t.x); t.y = fmaf(d, sctx[4 * c + 1], t.y); t.z = fmaf(d, sctx[4 * c + 2], t.z); t.w = fmaf(d, sctx[4 * c + 3], t.w); }
          o_wo[e] = t;
        }
        if (tid < I) gp[lo + P.o_b1 + tid] = (mm ? t_b1[tid] : 0.f) + (s_on ? sda[tid] : 0.f);
        mb_reduce_rows(red, 2 * H, [&](int e, float s) {
          if (!mm) s = 0.f;
          if (e < H) gp[lo + P.o_ln2g + e] = s + (s_on ? sv[SV_LN + e] : 0.f);
          else gp[lo + P.o_ln2b + e - H] = s + (s_on ? sv[SV_LN + 32 + e - H] : 0.f);
        });
        if (tid >= 128 && tid < 128 + H) {
          const int c = tid - 128;
          float s2 = 0.f, so = 0.f;
          if (mm) for (int k = 0; k < 16; ++k) { s2 += bsum[k * 32 + c]; so += bsum[512 + k * 32 + c]; }
          gp[lo + P.o_b2 + c] = s2 + (s_on ? sdd2[c] : 0.f);
          gp[lo + P.o_bo + c] = so + (s_on ? sdd1[c] : 0.f);
        }
      }
    }
    // ---------------- attention backward ----------------
    // side row of the q|k|v tile (token 128): k, v go into key row 128; q / dO / O of the row feed the rank-1 terms
    float qss = 0.f, kss = 0.f, vss = 0.f;
    if (is_side && has_side) {
      const bf16* qrow = reinterpret_cast<const bf16*>(P.qkv) + ((size_t)l * M + grow) * MG_Q;
      qss = __bfloat162float(qrow[lane]); kss = __bfloat162float(qrow[32 + lane]); vss = __bfloat162float(qrow[64 + lane]);
      if (P.rope_cos) {
        const int c = lane & 7;
        const float cs = P.rope_cos[(size_t)128 * 8 + c], sn = P.rope_sin[(size_t)128 * 8 + c];
        const float qo = __shfl_xor_sync(0xffffffffu, qss, 8), ko = __shfl_xor_sync(0xffffffffu, kss, 8);
        qss = bf16_round((lane & 8) ? qss * cs + qo * sn : qss * cs - qo * sn);
        kss = bf16_round((lane & 8) ? kss * cs + ko * sn : kss * cs - ko * sn);
      }
    }
    __syncthreads();   // the gradient staging in R1 has been read: R1 becomes dS / P~ / dqkv
    if (l == tl_l) VB_TL(tl_mega_bwd, 7);
    if (!is_side) {
      mbar_wait(b_qkv, par);
      if (P.rope_cos) {   // rotate q / k rows in place (the saved rows are un-rotated): cg 0/1: q head 0/1, cg 2/3: k head 0/1
        float x[16];
        mg_unpack8(*reinterpret_cast<const uint4*>(mg_chunk(sQ0, r, 2 * cg)), &x[0]);
        mg_unpack8(*reinterpret_cast<const uint4*>(mg_chunk(sQ0, r, 2 * cg + 1)), &x[8]);
        const int t = valid ? r : 0;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float cs = P.rope_cos[(size_t)t * 8 + c], sn = P.rope_sin[(size_t)t * 8 + c];
          const float lo_ = x[c], hi = x[c + 8];
          x[c] = lo_ * cs - hi * sn;
          x[c + 8] = hi * cs + lo_ * sn;
        }
        *reinterpret_cast<uint4*>(mg_chunk(sQ0, r, 2 * cg)) = mg_pack8(&x[0]);
        *reinterpret_cast<uint4*>(mg_chunk(sQ0, r, 2 * cg + 1)) = mg_pack8(&x[8]);
      }
      // key rows 129 .. 143 only have to be finite (their scores are masked): zero them (R3 held other tiles before)
      if (has_side && tid < 240) {
        const int blk = tid / 120, rem = tid - blk * 120;
        *reinterpret_cast<uint4*>(base + MB_R3 + blk * MB_QKV_BLK + (129 + (rem >> 3)) * 128 + (rem & 7) * 16) = make_uint4(0u, 0u, 0u, 0u);
      }
    } else if (has_side && s0) {
      *mg_elem(sQ0, 128, 32 + lane) = __float2bfloat16_rn(kss);
      *mg_elem(sQ1, 128, lane) = __float2bfloat16_rn(vss);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    float* xds = sv + SV_XDS; float* xpt = sv + SV_XPT; float* xq = sv + SV_XQ; float* xdo = sv + SV_XDO;
    // the query row the side group carries: token 128, or the CLS row (tile row 0, already rotated) in a CLS-only top layer
    const float qc = (is_side && ct) ? __bfloat162float(*mg_elem(sQ0, 0, lane)) : qss;
    for (int hd = hd_lo; hd < hd_hi; ++hd) {
      const MgOp Qk{aQ0 + hd * 32, 16, 16384, 0}, Qmn{aQ0 + hd * 32, 16384, 0, 1};
      const MgOp Kk{aQ0 + 64 + hd * 32, 16, 16384, 0}, Kmn{aQ0 + 64 + hd * 32, 16384, 0, 1};
      const MgOp Vk{aQ1 + hd * 32, 16, 16384, 0};
      const MgOp DOk{aCtx + hd * 32, 16, 16384, 0}, DOmn{aCtx + hd * 32, 16384, 0, 1};
      const MgOp DS_mn{aDS, 16384, 0, 1}, PT_mn{aPT, 16384, 0, 1};
      if (tid == 0 && !ct) {
        tc_fence_after();
        mg_issue(tmem + AB_S, Qk, Kk, KP, 1, false);     // S  = Q K^T   (all keys, incl. the side row as key 128)
        mg_issue(tmem + AB_DP, DOk, Vk, KP, 1, false);   // dP = dO V^T
        umma_commit(b_mma);
      }
      float dq_side = 0.f;   // side: dq of token 128 for this head, lane = hd * 16 + c
      float side_dq = 0.f, side_dk = 0.f, side_dv = 0.f;   // side: finished dqkv entries of token 128 (see below)
      if (is_side) {
        if (s_on) {
          // token 128 (or the CLS row) as a QUERY with plain FMAs: dq directly; its rank-1 contributions to dK / dV of every key are handed
          // to the key-row epilogue through shared memory (xds, xpt, xq, xdo)
          float qf[D], dof[D];
          float Di = 0.f;
#pragma unroll
          for (int c = 0; c < D; ++c) {
            qf[c] = __shfl_sync(0xffffffffu, qc, hd * D + c);
            dof[c] = __shfl_sync(0xffffffffu, dctxs, hd * D + c);
            Di = fmaf(dof[c], __shfl_sync(0xffffffffu, ctxs, hd * D + c), Di);
          }
          const float lse2 = P.lse[(((size_t)l * B + b) * MG_NH + hd) * T + stok] * MG_LOG2E;
          const uint64_t drow = ((uint64_t)(b * MG_NH + hd) * T + stok) * (uint64_t)Tpad;
          float dq[D];
#pragma unroll
          for (int c = 0; c < D; ++c) dq[c] = 0.f;
          // key j = sid (one key per lane of the group); key 128 -- the side row itself -- is lane 0's second key
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int j = sid + 128 * jj;
            if (j < T && (jj == 0 || sid == 0)) {
              float kr[D], vr[D];
              mg_unpack8(*reinterpret_cast<const uint4*>(mg_chunk(sQ0, j, 4 + 2 * hd)), &kr[0]);
              mg_unpack8(*reinterpret_cast<const uint4*>(mg_chunk(sQ0, j, 5 + 2 * hd)), &kr[8]);
              mg_unpack8(*reinterpret_cast<const uint4*>(mg_chunk(sQ1, j, 2 * hd)), &vr[0]);
              mg_unpack8(*reinterpret_cast<const uint4*>(mg_chunk(sQ1, j, 2 * hd + 1)), &vr[8]);
              float sdot = 0.f, dp = 0.f;
#pragma unroll
              for (int c = 0; c < D; ++c) { sdot = fmaf(qf[c], kr[c], sdot); dp = fmaf(dof[c], vr[c], dp); }
              const float p = exp2f(sdot * sl2 - lse2);
              const float keep = drop1(dc_att, drow + (uint64_t)j);
              const float ds = bf16_round(p * (dp * keep - Di) * scale);
              xds[j] = ds;
              xpt[j] = bf16_round(p * keep);
#pragma unroll
              for (int c = 0; c < D; ++c) dq[c] = fmaf(ds, kr[c], dq[c]);
            }
          }
          // dq: warp sums, then the four warps' partials in warp order
#pragma unroll
          for (int c = 0; c < D; ++c) {
            const float t = mg_wsum(dq[c]);
            if (lane == c) sdqp[sw * D + c] = t;
          }
          mb_bar_side();
          if ((lane >> 4) == hd) {
            const int c = lane & 15;
            dq_side = (sdqp[c] + sdqp[D + c]) + (sdqp[2 * D + c] + sdqp[3 * D + c]);
            if (ct && s0) sv[SV_DQC + lane] = dq_side;   // CLS-only top layer: dq of the CLS row -> tile row 0 (epilogue below)
          }
          if (s0) {
#pragma unroll
            for (int c = 0; c < D; ++c) if (lane == c) { xq[c] = qf[c]; xdo[c] = dof[c]; }
          }
        }
      } else if (!ct) {
        const float lse2 = (valid ? P.lse[(((size_t)l * B + b) * MG_NH + hd) * T + r] : 0.f) * MG_LOG2E;
        const float Di = s_dp[(2 * hd) * 128 + r] + s_dp[(2 * hd + 1) * 128 + r];
        const uint64_t drow = ((uint64_t)(b * MG_NH + hd) * T + (valid ? r : 0)) * (uint64_t)Tpad;
        mbar_wait(b_mma, ph_mma); ph_mma ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int ch = cg; ch < nch; ch += MG_CG) {
          const int c0 = ch * 16;
          if (c0 == 128 && T - c0 == 1) {   // key 128 is the only live key of the last chunk: one element (its dS / P~ go to k128)
            float s8[8], dp8[8], kpa[8];
            tmem_ld_32x8(my_tmem + AB_S + c0, s8);
            tmem_ld_32x8(my_tmem + AB_DP + c0, dp8);
            drop8(dc_att, (drow + (uint64_t)c0) >> 3, kpa);
            const float p = valid ? ex2_approx(fmaf(s8[0], sl2, -lse2)) : 0.f;
            const float pk = p * kpa[0];
            k128[r] = bf16_round(valid ? fmaf(dp8[0], pk, -p * Di) : 0.f);
            k128[128 + r] = bf16_round(pk);
            continue;
          }
          float s[16], dp[16];
          tmem_ld_32x16(my_tmem + AB_S + c0, s);
          tmem_ld_32x16(my_tmem + AB_DP + c0, dp);
          if (valid && c0 + 16 <= T) {   // full chunk of a live query row: no masking
#pragma unroll
            for (int j = 0; j < 16; j += 8) {
              float kpa[8];
              drop8(dc_att, (drow + (uint64_t)(c0 + j)) >> 3, kpa);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float p = ex2_approx(fmaf(s[j + q], sl2, -lse2));
                const float pk = p * kpa[q];                                  // dropped probability
                s[j + q] = fmaf(dp[j + q], pk, -p * Di);                      // dS / scale (the scale is applied to dQ, dK)
                dp[j + q] = pk;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; j += 8) {
              float kpa[8];
              if (c0 + j < T) drop8(dc_att, (drow + (uint64_t)(c0 + j)) >> 3, kpa);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const bool on = valid && (c0 + j + q < T);
                const float kq = (c0 + j < T) ? kpa[q] : 0.f;
                const float p = on ? ex2_approx(fmaf(s[j + q], sl2, -lse2)) : 0.f;
                const float pk = p * kq;
                s[j + q] = on ? fmaf(dp[j + q], pk, -p * Di) : 0.f;
                dp[j + q] = pk;
              }
            }
          }
          if (c0 < 128) {
#pragma unroll
            for (int j = 0; j < 16; j += 8) {
              *reinterpret_cast<uint4*>(mg_swz(sDS, r, (c0 + j) >> 3)) = mg_pack8(&s[j]);
              *reinterpret_cast<uint4*>(mg_swz(sPT, r, (c0 + j) >> 3)) = mg_pack8(&dp[j]);
            }
          } else {   // key 128 (the side row): its column of dS / P~ stays out of the tiles, FMAs use it (bf16 like the tiles)
            k128[r] = bf16_round(s[0]);
            k128[128 + r] = bf16_round(dp[0]);
          }
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      if (l == tl_l && hd == hd_lo) VB_TL(tl_mega_bwd, 8);
      if (tid == 0 && !ct) {
        tc_fence_after();
        mg_issue(tmem + AB_DQ, DS_k, Kmn, D, KM / 16, false);   // dQ[i,:] = sum_{j<128} dS[i,j] k_j
        mg_issue(tmem + AB_DK, DS_mn, Qmn, D, 8, false);        // dK[j,:] = sum_i dS[i,j] q_i
        mg_issue(tmem + AB_DV, PT_mn, DOmn, D, 8, false);       // dV[j,:] = sum_i P~[i,j] dO_i
        umma_commit(b_mma);
      }
      if (l == tl_l && hd == hd_lo) VB_TL(tl_mega_bwd, 14);
      // side warp meanwhile: the side row as a KEY.  dk_128 = scale * sum_i dS[i,128] q_i + (own query term),
      // dv_128 = sum_i P~[i,128] dO_i + (own query term); lanes 0..15: dk column c, lanes 16..31: dv column c
      if (is_side && has_side) {
        const int c = lane & 15;
        const bool isv = lane >= 16;
        float acc = 0.f;
        if (!ct) {   // (CLS-only top layer: the tensor-core query rows carry no gradient)
          const uint8_t* tile = isv ? sCtx : sQ0;
          const float* wv = isv ? k128 + 128 : k128;
          float part = 0.f;
#pragma unroll 8
          for (int ii = 0; ii < 32; ++ii) {
            const int i = 32 * sw + ii;
            part = fmaf(wv[i], __bfloat162float(*mg_elem(tile, i, hd * D + c)), part);
          }
          sp[sw * 32 + lane] = part;
          mb_bar_side();
          acc = (sp[lane] + sp[32 + lane]) + (sp[64 + lane] + sp[96 + lane]);
        }
        float dk = acc * scale + xds[128] * __shfl_sync(0xffffffffu, qc, hd * D + c);
        float dv = acc + xpt[128] * __shfl_sync(0xffffffffu, dctxs, hd * D + c);
        // gather the row: lane n of this head's 16 columns holds dq; dk / dv come from lanes c / 16 + c
        float dq = ct ? 0.f : dq_side;   // (CLS-only: token 128 is no query of the top layer)
        const float dk_n = __shfl_sync(0xffffffffu, dk, lane & 15), dv_n = __shfl_sync(0xffffffffu, dv, 16 + (lane & 15));
        float dqn = dq, dkn = dk_n, dvn = dv_n;
        if (P.rope_cos) {   // inverse rotation of dq / dk (bf16-rounded first, like the tensor-core rows)
          dqn = bf16_round(dqn); dkn = bf16_round(dkn);
          const int c8 = lane & 7;
          const float cs = P.rope_cos[(size_t)128 * 8 + c8], sn = -P.rope_sin[(size_t)128 * 8 + c8];
          const float qo = __shfl_xor_sync(0xffffffffu, dqn, 8), ko = __shfl_xor_sync(0xffffffffu, dkn, 8);
          dqn = (lane & 8) ? dqn * cs + qo * sn : dqn * cs - qo * sn;
          dkn = (lane & 8) ? dkn * cs + ko * sn : dkn * cs - ko * sn;
        }
        side_dq = bf16_round(dqn); side_dk = bf16_round(dkn); side_dv = bf16_round(dvn);
      }
      if (!is_side && !ct) { mbar_wait(b_mma, ph_mma); ph_mma ^= 1; tc_fence_after(); }
      if (l == tl_l && hd == hd_lo) VB_TL(tl_mega_bwd, 15);
      if (tid == 0 && hd == hd_hi - 1) {
        // dS / P~ (R1 @0 .. 64K) are dead: the u / z rows of this layer and the pre-GELU rows of the layer below arrive there
        mbar_expect_tx(b_low, 32768);
        tma_load_4d(sU, &TM.u, b_low, 0, 0, b, l);
        tma_load_4d(sZ, &TM.z, b_low, 0, 0, b, l);
        if (l > 0) load_a(l - 1);
      }
      if (l == tl_l && hd == hd_lo) VB_TL(tl_mega_bwd, 9);
      // ---- dQ / dK / dV epilogue -> dqkv tile (R1 @64K).  In a CTA pair the pieces also go into the peer's tile, which
      //      is free once the peer has passed stage 2 of its upper half (its u2 / hmid tiles sat there): barrier B. ----
      if (csz == 2) mb_cluster_wait();
      if (l == tl_l && hd == hd_lo) VB_TL(tl_mega_bwd, 18);
      if (is_side && has_side && s0) {
        // dqkv row of token 128: dq in lanes hd * 16 + c; (dk, dv) of column c = lane & 15 in every lane
        if ((lane >> 4) == hd) {
          sv[SV_DQKV + lane] = side_dq;
          if (csz == 2) mb_st_async4(mb_peer_addr(&sv[SV_DQKV + lane], crank ^ 1u), side_dq, x_peer);
        }
        if (lane < 16) {
          sv[SV_DQKV + 32 + hd * D + lane] = side_dk;
          sv[SV_DQKV + 64 + hd * D + lane] = side_dv;
          if (csz == 2) {
            mb_st_async4(mb_peer_addr(&sv[SV_DQKV + 32 + hd * D + lane], crank ^ 1u), side_dk, x_peer);
            mb_st_async4(mb_peer_addr(&sv[SV_DQKV + 64 + hd * D + lane], crank ^ 1u), side_dv, x_peer);
          }
        }
      }
      if (!is_side) {
        if (P.rope_cos) {
          // the inverse rotation pairs columns c and c + 8: cg 0 takes the dQ row, cg 1 the dK row, cg 2 the dV row
          if (cg < 3) {
            float x[16];
            if (!ct) {
              tmem_ld_32x16(my_tmem + (cg == 0 ? AB_DQ : cg == 1 ? AB_DK : AB_DV), x);
#pragma unroll
              for (int c = 0; c < 16; ++c) x[c] = valid ? x[c] * (cg < 2 ? scale : 1.f) : 0.f;
            } else {   // CLS-only top layer: no tensor-core terms; dq exists for the CLS row (tile row 0) only
#pragma unroll
              for (int c = 0; c < 16; ++c) x[c] = (cg == 0 && r == 0) ? sv[SV_DQC + hd * D + c] : 0.f;
            }
            if (s_on && valid) {
              if (cg == 0) {   // key 128 of the tensor-core query rows
                if (!ct && has_side) {
                const float a = k128[r] * scale;
#pragma unroll
                for (int c = 0; c < 16; ++c) x[c] = fmaf(a, __bfloat162float(*mg_elem(sQ0, 128, 32 + hd * D + c)), x[c]);
                }
              } else {         // the side QUERY's rank-1 term
                const float a = cg == 1 ? xds[r] : xpt[r];
                const float* xr = cg == 1 ? xq : xdo;
#pragma unroll
                for (int c = 0; c < 16; ++c) x[c] = fmaf(a, xr[c], x[c]);
              }
            }
            if (cg < 2) {
#pragma unroll
              for (int c = 0; c < 16; ++c) x[c] = bf16_round(x[c]);
              const int t = valid ? r : 0;
#pragma unroll
              for (int c = 0; c < 8; ++c) {   // transpose of the rotation
                const float cs = P.rope_cos[(size_t)t * 8 + c], sn = -P.rope_sin[(size_t)t * 8 + c];
                const float lo_ = x[c], hi = x[c + 8];
                x[c] = lo_ * cs - hi * sn;
                x[c + 8] = hi * cs + lo_ * sn;
              }
            }
            uint8_t* blk = cg < 2 ? sDQ : sDQ + 16384;
            const int ch0 = cg == 0 ? 2 * hd : cg == 1 ? 4 + 2 * hd : 2 * hd;
            const uint4 p0 = mg_pack8(&x[0]), p1 = mg_pack8(&x[8]);
            uint8_t* d0 = mg_chunk(blk, r, ch0); uint8_t* d1 = mg_chunk(blk, r, ch0 + 1);
            *reinterpret_cast<uint4*>(d0) = p0; *reinterpret_cast<uint4*>(d1) = p1;
            if (csz == 2) { mb_st_async16(mb_peer_addr(d0, crank ^ 1u), p0, x_peer); mb_st_async16(mb_peer_addr(d1, crank ^ 1u), p1, x_peer); }
          }
        } else {
          // 6 pieces of 8 columns per row: dQ0 dQ1 | dK0 dK1 | dV0 dV1 ; group cg takes pieces cg and cg + 4
#pragma unroll 1
          for (int pc = cg; pc < 6; pc += MG_CG) {
            const int kind = pc >> 1, c8 = (pc & 1) * 8;   // 0 = dQ, 1 = dK, 2 = dV
            float o[8];
            if (!ct) {
              tmem_ld_32x8(my_tmem + (kind == 0 ? AB_DQ : kind == 1 ? AB_DK : AB_DV) + c8, o);
#pragma unroll
              for (int c = 0; c < 8; ++c) o[c] = valid ? o[c] * (kind < 2 ? scale : 1.f) : 0.f;
            } else {   // CLS-only top layer: no tensor-core terms; dq exists for the CLS row (tile row 0) only
#pragma unroll
              for (int c = 0; c < 8; ++c) o[c] = (kind == 0 && r == 0) ? sv[SV_DQC + hd * D + c8 + c] : 0.f;
            }
            if (s_on && valid) {
              if (kind == 0) {
                if (!ct && has_side) {
                  const float a = k128[r] * scale;
#pragma unroll
                  for (int c = 0; c < 8; ++c) o[c] = fmaf(a, __bfloat162float(*mg_elem(sQ0, 128, 32 + hd * D + c8 + c)), o[c]);
                }
              } else {
                const float a = kind == 1 ? xds[r] : xpt[r];
                const float* xr = kind == 1 ? xq : xdo;
#pragma unroll
                for (int c = 0; c < 8; ++c) o[c] = fmaf(a, xr[c8 + c], o[c]);
              }
            }
            uint8_t* blk = kind < 2 ? sDQ : sDQ + 16384;
            const int ch = (kind == 0 ? 2 * hd : kind == 1 ? 4 + 2 * hd : 2 * hd) + (pc & 1);
            const uint4 pk = mg_pack8(o);
            uint8_t* dst = mg_chunk(blk, r, ch);
            *reinterpret_cast<uint4*>(dst) = pk;
            if (csz == 2) mb_st_async16(mb_peer_addr(dst, crank ^ 1u), pk, x_peer);
          }
        }
      }
      fence_proxy_async();
      tc_fence_before();
      if (l == tl_l && hd == hd_lo) VB_TL(tl_mega_bwd, 19);
      if (csz == 2) { mbar_wait(b_x, ph_x); ph_x ^= 1; }   // the peer's pieces have landed
      __syncthreads();                                     // (and this CTA's own)
      if (csz == 2) fence_proxy_async();
      tc_fence_after();
    }
    if (ct && !is_side && r == 0) {
#pragma unroll
      for (int j = 0; j < HC; ++j) dz[j] = sv[SV_DZC + hc0 + j];   // dh of the CLS row (side chain) rejoins the tile
    }
    if (l == tl_l) VB_TL(tl_mega_bwd, 10);
    // ---------------- lower: du = dqkv Wqkv ; dWqkv (+ dbqkv) += dqkv^T [u | 1] ; LN1 backward (+ dh) -> dz ----------------
    float mu1 = 0.f, rs1 = 0.f;
    float us_ = 0.f, zs_ = 0.f, mu1s = 0.f, rs1s = 0.f;
    if (!is_side) {
      mu1 = P.stats[(size_t)(4 * l) * M + grow]; rs1 = P.stats[(size_t)(4 * l + 1) * M + grow];
      mbar_wait(b_low, par);
      if (cg == 0) *reinterpret_cast<uint32_t*>(mg_chunk(sU, r, H / 8)) = valid ? 0x00003F80u : 0u;   // ones column: dbqkv
    } else if (has_side) {
      const size_t lr = (size_t)l * M + grow;
      us_ = __bfloat162float(reinterpret_cast<const bf16*>(P.u)[lr * H + lane]);
      zs_ = P.z[lr * H + lane];
      mu1s = P.stats[(size_t)(4 * l) * M + grow]; rs1s = P.stats[(size_t)(4 * l + 1) * M + grow];
      if (s0) sv[SV_U + lane] = us_;
    }
    if (tid == 0 && l > 0) load_ctx(l - 1);   // the dO tile is dead
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(b_wq, par);
      mg_issue(tmem + LB_DU, DQ_k, WQ_mn, H, MG_Q / 16, false);     // du[row, h] = sum_n dqkv[row,n] Wqkv[n,h]
      umma_commit(b_mma);                                           // the LayerNorm stage only needs du
      mg_issue(tmem + LB_WQ, DQ_mn, U_mn, H + 16, 8, false);        // dWqkv[n, h] = sum_rows dqkv[row,n] u[row,h] ; column H: dbqkv[n]
      umma_commit(b_wg);
    }
    float gam1[HC], bet1[HC];
    if (!is_side) {
      mbar_wait(b_mma, ph_mma); ph_mma ^= 1;
      tc_fence_after();
      float du[HC], xh[HC], g[HC];
      tmem_ld_32x8(my_tmem + LB_DU + hc0, du);
      {
        const float4 x0 = *mg_f32(sZ, r, hc0), x1 = *mg_f32(sZ, r, hc0 + 4);
        xh[0] = x0.x; xh[1] = x0.y; xh[2] = x0.z; xh[3] = x0.w; xh[4] = x1.x; xh[5] = x1.y; xh[6] = x1.z; xh[7] = x1.w;
      }
#pragma unroll
      for (int j = 0; j < HC; ++j) {
        xh[j] = (xh[j] - mu1) * rs1;
        du[j] = valid ? bf16_round(du[j]) : 0.f;
        gam1[j] = du[j] * xh[j];
        bet1[j] = du[j];
      }
      float c1, c2;
      mb_ln_bwd_sums(du, xh, g1, hc0, s_ex, r, cg, g, c1, c2);
#pragma unroll
      for (int j = 0; j < HC; ++j) dz[j] = valid ? dz[j] + rs1 * (g[j] - c1 - xh[j] * c2) : 0.f;   // gradient of the layer input
    } else if (has_side) {
      mbar_wait(b_wq, par);
      // du[h] = sum_n dqkv[n] Wqkv[n][h]  (h = lane); the dqkv row of token 128 sits in sv[SV_DQKV ..] (both heads)
      // warps 0..2 contract 32 rows of Wqkv each
      float part = 0.f;
      if (sw < 3) {
#pragma unroll 8
        for (int nn = 0; nn < 32; ++nn) part = fmaf(sv[SV_DQKV + 32 * sw + nn], __bfloat162float(*mg_elem(sWq, 32 * sw + nn, lane)), part);
      }
      sp[sw * 32 + lane] = part;
      mb_bar_side();
      const float du = bf16_round((sp[lane] + sp[32 + lane]) + sp[64 + lane]);
      const float xh = (zs_ - mu1s) * rs1s;
      const float gg = du * g1[lane];
      const float c1 = mg_wsum(gg) * (1.f / H), c2 = mg_wsum(gg * xh) * (1.f / H);
      dzs = (ct ? 0.f : dhs) + rs1s * (gg - c1 - xh * c2);   // (CLS-only top layer: token 128 has no gradient above this point)
      if (s0) { sv[SV_LN1 + lane] = du * xh; sv[SV_LN1 + 32 + lane] = du; }
    }
    {
      // parameter gradients of the QKV projection and LN1 -> gpart (staging in the dead dS region would collide with u / z:
      // the P~ half, R1 @32K, is free only when no pre-GELU rows are arriving, so the side-vector area's neighbours are used)
      if (!is_side) mbar_wait(b_wg, par);
      tc_fence_after();
      float4* t_wq = reinterpret_cast<float4*>(base + MB_R3);               // [96 rows][8 pieces] dWqkv (12 KB; q|k|v tiles are dead)
      float* t_bq = reinterpret_cast<float*>(base + MB_R3 + 12288);         // [96] dbqkv
      float* red = reinterpret_cast<float*>(base + MB_R3 + 16384);          // [2 * H][128] LN1 gamma / beta terms (32 KB)
      if (!is_side) {
        if ((warp & 3) < 3) {  // dWqkv rows n = lanes 0..95
          float v[HC];
          tmem_ld_32x8(my_tmem + LB_WQ + hc0, v);
          t_wq[r * 8 + ((2 * cg) ^ (r & 7))] = make_float4(v[0], v[1], v[2], v[3]);
          t_wq[r * 8 + ((2 * cg + 1) ^ (r & 7))] = make_float4(v[4], v[5], v[6], v[7]);
          if (cg == 0) {
            float b8[8];
            tmem_ld_32x8(my_tmem + LB_WQ + H, b8);
            t_bq[r] = b8[0];
          }
        }
#pragma unroll
        for (int j = 0; j < HC; ++j) { red[(hc0 + j) * 128 + r] = gam1[j]; red[(H + hc0 + j) * 128 + r] = bet1[j]; }
      }
      tc_fence_before();
      __syncthreads();
      if (l == tl_l) VB_TL(tl_mega_bwd, 11);
      if (tid == 0 && l > 0) { load_upper(l - 1); load_wq(l - 1); }   // u / z / dqkv are consumed: m, u2, hmid of the layer below
      if (!is_side && (csz == 1 || !lead)) {   // (CTA pair: rank 1 writes this half, rank 0 wrote the MLP half)
        const size_t lo = (size_t)P.off_layer0 + (size_t)l * P.layer_stride;
        float4* o_wq = reinterpret_cast<float4*>(gp + lo + P.o_wqkv);
        const float* sdq = sv + SV_DQKV; const float* su = sv + SV_U;
        for (int e = tid; e < MG_Q * 8; e += MG_MAIN) {
          const int rr = e >> 3, c = e & 7;
          float4 t = t_wq[rr * 8 + (c ^ (rr & 7))];
          if (has_side) { const float d = sdq[rr]; t.x = fmaf(d, su[4 * c], t.x); t.y = fmaf(d, su[4 * c + 1], t.y); t.z = fmaf(d, su[4 * c + 2], t.z); t.w = fmaf(d, su[4 * c + 3], t.w); }
          o_wq[e] = t;
        }
        if (tid < MG_Q) gp[lo + P.o_bqkv + tid] = t_bq[tid] + (has_side ? sdq[tid] : 0.f);
        mb_reduce_rows(red, 2 * H, [&](int e, float s) {
          if (e < H) gp[lo + P.o_ln1g + e] = s + (has_side ? sv[SV_LN1 + e] : 0.f);
          else gp[lo + P.o_ln1b + e - H] = s + (has_side ? sv[SV_LN1 + 32 + e - H] : 0.f);
        });
      }
      __syncthreads();   // staging (R3) is read: the next layer's ddelta / da tiles go there
      // (signalled by a side thread: the release is a MEMBAR.ALL.GPU, ~2 k cycles that thread 0 -- the TMA / MMA issuer --
      //  must not spend here)
      if (tid == MG_MAIN + 96 && PA.done) {   // group 1 + l = layer l; group 1 + L = final LayerNorm + head (written first of all)
        if (l == L - 1) mb_signal(PA.done, 1 + L);
        mb_signal(PA.done, 1 + l);
      }
    }
  }

  // =============================== embedding (embedding.py:79-100, tokenization.py:43-50) ===============================
  VB_TL(tl_mega_bwd, 12);
  {
    const DropCtx dc = make_drop(P.p_hidden, seed, step, VITB200_SITE_EMB);
    float g[HC];
    float gs = 0.f;
    if (!is_side) {
      float kp[8];
      drop8(dc, (grow * H + hc0) >> 3, kp);
#pragma unroll
      for (int j = 0; j < HC; ++j) g[j] = valid ? dz[j] * kp[j] : 0.f;
      if (PA.dz0 && valid && lead) {   // d loss / d z0 for a trainable preprocessor in front of the model (engine.pixel_grad)
        float4* o = reinterpret_cast<float4*>(PA.dz0 + grow * H + hc0);
        o[0] = make_float4(dz[0], dz[1], dz[2], dz[3]); o[1] = make_float4(dz[4], dz[5], dz[6], dz[7]);
      }
      if (lead) {
        if (P.off_pos >= 0 && valid) {   // learned positions: one sample per CTA, so the row IS the partial gradient
          float4* o = reinterpret_cast<float4*>(gp + P.off_pos + (size_t)r * H + hc0);
          o[0] = make_float4(g[0], g[1], g[2], g[3]); o[1] = make_float4(g[4], g[5], g[6], g[7]);
        }
        if (r == 0) {                    // CLS row: gradient of cls_token, no patch
          float4* o = reinterpret_cast<float4*>(gp + P.off_cls + hc0);
          o[0] = make_float4(g[0], g[1], g[2], g[3]); o[1] = make_float4(g[4], g[5], g[6], g[7]);
        }
      }
      if (r == 0) {
#pragma unroll
        for (int j = 0; j < HC; ++j) g[j] = 0.f;
      }
      *reinterpret_cast<uint4*>(mg_chunk(sD, r, cg)) = mg_pack8(g);
      *reinterpret_cast<uint4*>(mg_chunk(sD, r, 4 + cg)) = make_uint4(0u, 0u, 0u, 0u);
      // patch windows of the tokens (bf16), as in the forward kernel
      const int nchunk = P.P / 8;
      const bool vec = (P.S % 4 == 0) && (P.L % 4 == 0);
      const bool has = valid && r >= 1 && (r - 1) < P.n_valid;
      const float* xp = P.x + bsrc * P.L + (size_t)(r >= 1 ? r - 1 : 0) * P.S;
      for (int c = cg; c < 8; c += MG_CG) {
        float v[8];
        if (c < nchunk && has) {
          if (vec) {
            const float4 a0 = *reinterpret_cast<const float4*>(xp + c * 8), a1 = *reinterpret_cast<const float4*>(xp + c * 8 + 4);
            v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = xp[c * 8 + q];
          }
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = 0.f;
        }
        *reinterpret_cast<uint4*>(mg_chunk(sX, r, c)) = mg_pack8(v);
      }
    } else if (has_side) {
      gs = dzs * drop1(dc, grow * H + lane);
      if (s0) {
        if (lead) {
          if (PA.dz0) PA.dz0[grow * H + lane] = dzs;
          if (P.off_pos >= 0) gp[P.off_pos + (size_t)128 * H + lane] = gs;
        }
        sv[SV_DD2 + lane] = bf16_round(gs);
        const bool has = 127 < P.n_valid;
        const float* xp = P.x + bsrc * P.L + (size_t)127 * P.S;
        sv[SV_M + lane] = (has && lane < P.P) ? bf16_round(xp[lane]) : 0.f;
        sv[SV_M + 32 + lane] = (has && lane + 32 < P.P) ? bf16_round(xp[lane + 32]) : 0.f;
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    const MgOp X_mn{smem_u32(sX), 16384, 0, 1};
    if (tid == 0) {
      tc_fence_after();
      mg_issue(tmem + EB_WP, D_mn, X_mn, P.P, 8, false);   // dWp[h, j] = sum_rows dtok[row, h] x[row, j]
      umma_commit(b_mma);
    }
    float acc_bp = 0.f;
    if (!is_side) {
      acc_bp = mb_colsum8(sD, tid & 31, (tid >> 5) * 8);
      bsum[tid] = acc_bp;
      mbar_wait(b_mma, ph_mma); ph_mma ^= 1;
      tc_fence_after();
    }
    __syncthreads();
    if (!is_side && lead) {
      if ((warp & 3) == 0) {  // dWp rows h = lanes 0..31
#pragma unroll 1
        for (int c0 = hc0; c0 < P.P; c0 += 8 * MG_CG) {
          float v[8];
          tmem_ld_32x8(my_tmem + EB_WP + c0, v);
          if (has_side) {
            const float d = sv[SV_DD2 + r];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = fmaf(d, sv[SV_M + c0 + q], v[q]);
          }
          float4* op = reinterpret_cast<float4*>(gp + P.off_wp + (size_t)r * P.P + c0);
          op[0] = make_float4(v[0], v[1], v[2], v[3]); op[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
      if (tid >= 128 && tid < 128 + H) {
        const int c = tid - 128;
        float s = 0.f;
        for (int k = 0; k < 16; ++k) s += bsum[k * 32 + c];
        gp[P.off_bp + c] = s + (has_side ? sv[SV_DD2 + c] : 0.f);
      }
    }
  }
  VB_TL(tl_mega_bwd, 13);
  tc_fence_before();
  __syncthreads();
  if (tid == MG_MAIN + 96 && PA.done) mb_signal(PA.done, 0);   // group 0 = embeddings (cls, positions, patch projection)
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace vb

using namespace vb;

VB_TL_EXPORT(vitb200_tl_mega_bwd, vb::tl_mega_bwd)

extern "C" int vitb200_mega_bwd_supported(int H, int heads, int T, int P, int C, int layers, int B, int cluster) {
  if (!vitb200_mega_supported(H, heads, T, P, C, layers, 0)) return 0;
  if (P % 16 != 0) return 0;                       // the dWp MMA needs N = P to be a multiple of 16
  if (cluster != 1 && cluster != 2) return 0;
  if (B < 1 || B * cluster > 148) return 0;        // one sample per CTA (pair), one wave
  return 1;
}
extern "C" size_t vitb200_mega_bwd_smem_bytes(int layers) { return 1024 + MB_PRM + (size_t)(layers * 64 + 32 + 32) * sizeof(float); }

extern "C" int vitb200_mega_bwd(const vitb200_mega_bwd_args* pa, void* stream) {
  if (!pa) return VITB200_ERR_ARG;
  const vitb200_mega_fwd_args* a = &pa->f;
  if (!a->x || !a->params || !a->shadow || !a->z || !a->hmid || !a->u || !a->u2 || !a->qkv || !a->ctx || !a->a || !a->m ||
      !a->stats || !a->lse || !a->s_cls || !a->logits || !pa->labels || !pa->gpart ||
      (a->rows && (!a->rng || !a->rows_base || pa->loss_kind == VITB200_LOSS_GIVEN)) ||
      (a->defer_loss && (!a->ws || !a->loss || !a->labels)))
    return VITB200_ERR_ARG;
  const int T = a->Np + 1;
  if (!vitb200_mega_bwd_supported(MG_H, MG_NH, T, a->P, a->C, a->layers, a->B, a->cluster)) return VITB200_ERR_SHAPE;
  if (pa->loss_kind < 0 || pa->loss_kind > 3) return VITB200_ERR_ARG;
  const int L = a->layers, B = a->B;
  const bf16* sh = reinterpret_cast<const bf16*>(a->shadow);
  MegaBwdMaps tm;
  int rc;
  const bf16* l0 = sh + a->off_layer0;
  if ((rc = get_tmap_weight(l0 + a->o_wqkv, MG_H, MG_Q, L, a->layer_stride, MG_Q, &tm.wq))) return rc;
  if ((rc = get_tmap_weight(l0 + a->o_wo, MG_H, MG_H, L, a->layer_stride, MG_H, &tm.wo))) return rc;
  if ((rc = get_tmap_weight(l0 + a->o_w1, MG_H, MG_I, L, a->layer_stride, MG_I, &tm.w1))) return rc;
  if ((rc = get_tmap_weight(l0 + a->o_w2, MG_I, MG_H, L, a->layer_stride, MG_H, &tm.w2))) return rc;
  if ((rc = get_tmap_act(a->z, 2 * MG_H, T, B, L + 1, &tm.z))) return rc;
  if ((rc = get_tmap_act(a->hmid, 2 * MG_H, T, B, L, &tm.hmid))) return rc;
  if ((rc = get_tmap_act(a->u, MG_H, T, B, L, &tm.u))) return rc;
  if ((rc = get_tmap_act(a->u2, MG_H, T, B, L, &tm.u2))) return rc;
  if ((rc = get_tmap_act(a->qkv, MG_Q, T, B, L, &tm.qkv))) return rc;
  if ((rc = get_tmap_act(a->ctx, MG_H, T, B, L, &tm.ctx))) return rc;
  if ((rc = get_tmap_act(a->a, MG_I, T, B, L, &tm.a))) return rc;
  if ((rc = get_tmap_act(a->m, MG_I, T, B, L, &tm.m))) return rc;
  const int smem = (int)vitb200_mega_bwd_smem_bytes(L);
  if (smem > 227 * 1024) return VITB200_ERR_SHAPE;
  static int max_set = 0;
  if (smem > max_set) {
    cudaError_t e = cudaFuncSetAttribute(mega_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return vb_cuda_error(e);
    max_set = smem;
  }
  vb_launch_pdl_cluster(mega_bwd_kernel, dim3(B * a->cluster), dim3(MB_THREADS), smem, (cudaStream_t)stream, a->cluster, tm, *pa);
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}
