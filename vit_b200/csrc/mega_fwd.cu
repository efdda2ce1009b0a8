// Whole-network forward kernel (bf16, H = 32, 2 heads of 16, T <= 129).  See include/vit_b200.h and mega_common.cuh.
//
// A CTA loops over samples.  Per sample the phases are (every phase ends in one CTA barrier):
//   embed   : patch windows -> A tile, GEMM P->H, +bias / CLS / pos, dropout -> z0 (registers), LN1 -> u
//   qkv     : u Wqkv^T + b -> q|k|v tile in shared memory (144 rows: row 128 = the side row, rows 129.. zero)
//   per layer
//     attn  : S_h = Q_h K_h^T for both heads (UMMA descriptors start at the head's columns inside the tile: nothing is
//             copied or transposed), softmax split over 4 threads per query row, P~ -> smem, O_h = P~ V_h, ctx tile
//     hop 1 : ctx Wo^T, dropout, +residual, LN2 -> u2          hop 2 : u2 W1^T + b, GELU -> a, m
//     hop 3 : m W2^T, dropout, +residual -> z, next LN1 -> u    hop 4 : QKV of the next layer (or: final LN of the CLS row)
//   head    : logits, loss term of the sample
// The residual row lives in registers for the whole network (8 columns per thread); tensors backward needs leave
// through TMA stores from the same swizzled images the next GEMM reads.  Weights of ONE layer are resident; each
// weight tile is re-loaded for the next layer right after its GEMM (the load has the rest of the layer to land).
// Warp 16 carries the 129th token with FMAs (weights read from the same shared-memory tiles) and shares every barrier.
#include "mega_common.cuh"

namespace vb {

// shared-memory plan (byte offsets from the 1 KB-aligned base)
constexpr uint32_t MF_A = 0;                        // 16 KB  A operand (patches -> u -> u2 -> u ...), also a store image
constexpr uint32_t MF_QKV = MF_A + 16384;           // 2 x 18 KB  q|k block, v block; 144 rows
constexpr uint32_t MF_QKV_BLK = 18432;
constexpr uint32_t MF_H = MF_QKV + 2 * MF_QKV_BLK;  // 16 KB  fp32 residual rows (store image of z0 / hmid / z)
constexpr uint32_t MF_U = MF_H + 16384;             // 64 KB  attention: P~ (3 x 16 KB);  MLP: m (32 KB) + a (32 KB)
constexpr uint32_t MF_CTX = MF_U + 65536;           // 16 KB  attention output tile (in a CTA pair each CTA writes its head's
                                                    //        columns into BOTH CTAs' tiles)
constexpr uint32_t MF_WO = MF_CTX + 16384;          // 4 KB
constexpr uint32_t MF_W1 = MF_WO + 4096;            // 16 KB
constexpr uint32_t MF_W2 = MF_W1 + 16384;           // 8 KB
constexpr uint32_t MF_WQ = MF_W2 + 8192;            // 12 KB
constexpr uint32_t MF_WP = MF_WQ + 12288;           // 4 KB
constexpr uint32_t MF_BAR = MF_WP + 4096;           // barriers, TMEM slot
constexpr uint32_t MF_EX = MF_BAR + 256;            // float2 [4][128] LayerNorm statistics exchange
constexpr uint32_t MF_MX = MF_EX + 4096;            // float [4][128] row-max exchange
constexpr uint32_t MF_SM = MF_MX + 2048;            // float [2][4][128] row-sum exchange (both heads)
constexpr uint32_t MF_SC = MF_SM + 4096;            // float [32] final-LayerNorm'd CLS row, [32] side-row attention output,
                                                    // [32] CLS row of the residual stream
constexpr uint32_t MF_SP = MF_SC + 512;             // 1 KB  side group: partial sums [4][32], attention partials [4][20]
constexpr uint32_t MF_PRM = MF_SP + 1024;           // staged fp32 vectors
// The side row is carried by FOUR warps (see mega_bwd.cu): 32-wide vectors are replicated per warp (lane = column),
// keys / MLP columns are spread over the 128 lanes, contractions over 128 are split by warp and combined in warp order.
constexpr int MF_SIDE_WARPS = 4;
constexpr int MF_THREADS = MG_MAIN + 32 * MF_SIDE_WARPS;   // 640
__device__ __forceinline__ void mf_bar_side() { asm volatile("bar.sync 2, 128;" ::: "memory"); }
// TMEM columns
constexpr uint32_t MC_S0 = 0, MC_S1 = 160, MC_O = 320, MC_LIN = 352, MC_COLS = 512;

VB_TL_DECL(tl_mega_fwd)

// ---- CTA pair (cluster of 2: one attention head per CTA, everything else computed by both) ----
__device__ __forceinline__ uint32_t mf_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mf_peer_addr(const void* p, uint32_t peer) {   // shared::cluster address of the peer's copy
  uint32_t a;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(smem_u32(p)), "r"(peer));
  return a;
}
__device__ __forceinline__ void mf_st_peer16(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void mf_st_peer4(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// DSMEM push that signals the RECEIVER's mbarrier itself (st.async ... mbarrier::complete_tx::bytes): the receiver posts the
// byte count it expects per layer and waits on its own barrier -- no cluster-wide release / acquire barrier (whose
// arrive.release is a MEMBAR.ALL.GPU) between the push and the consumer.
__device__ __forceinline__ void mf_st_async16(uint32_t addr, uint4 v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void mf_st_async4(uint32_t addr, float v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(addr), "r"(__float_as_uint(v)),
               "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void mf_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
// "my exchange buffers may be overwritten" (write-after-read hand-off, nothing is published): no memory ordering needed.
// The .release form costs a MEMBAR.ALL.GPU + ERRBAR (~2 k cycles) in front of the arrive.
__device__ __forceinline__ void mf_cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void mf_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void mf_cluster_sync() { mf_cluster_arrive(); mf_cluster_wait(); }

struct MegaFwdMaps { CUtensorMap wp, wq, wo, w1, w2, z, hmid, u, u2, qkv, ctx, a, m; };

// LayerNorm statistics of a row spread over 4 threads (8 columns each): per-group (mean, M2) merged with Chan's formula
__device__ __forceinline__ void mf_row_stats(const float (&x)[MG_HC], float2* s_ln, int r, int cg, float eps, float& mu, float& rs) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < MG_HC; ++j) s += x[j];
  const float mc = s * (1.f / MG_HC);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < MG_HC; ++j) { const float d = x[j] - mc; q = fmaf(d, d, q); }
  s_ln[cg * 128 + r] = make_float2(mc, q);
  mg_bar_main();
  float2 p[MG_CG];
#pragma unroll
  for (int g = 0; g < MG_CG; ++g) p[g] = s_ln[g * 128 + r];
  float m = 0.f;
#pragma unroll
  for (int g = 0; g < MG_CG; ++g) m += p[g].x;
  mu = m * (1.f / MG_CG);
  float m2 = 0.f;
#pragma unroll
  for (int g = 0; g < MG_CG; ++g) { const float d = p[g].x - mu; m2 += fmaf(d * d, (float)MG_HC, p[g].y); }
  rs = rsqrtf(m2 * (1.f / MG_H) + eps);   // (s_ln is rewritten only after the CTA barrier that ends the phase)
}

__device__ __forceinline__ float mf_loss_term(const float (&lg)[4], const void* labels, int b, int C, int kind) {
  if (kind == VITB200_LOSS_CE) {
    const long long y = reinterpret_cast<const long long*>(labels)[b];
    float mx = -INFINITY, ly = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) if (c < C) { mx = fmaxf(mx, lg[c]); if (c == (int)y) ly = lg[c]; }
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) if (c < C) se += expf(lg[c] - mx);
    return mx + logf(se) - ly;
  }
  const float* yl = reinterpret_cast<const float*>(labels) + (size_t)b * C;
  float t = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (c < C) {
      const float d = lg[c] - yl[c];
      t += kind == VITB200_LOSS_L1 ? fabsf(d) : d * d;
    }
  }
  return t;
}

__global__ void __launch_bounds__(MF_THREADS, 1)
mega_fwd_kernel(const __grid_constant__ MegaFwdMaps TM, const vitb200_mega_fwd_args P) {
  constexpr int H = MG_H, I = MG_I, HC = MG_HC;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t *sA = base + MF_A, *sQ0 = base + MF_QKV, *sQ1 = base + MF_QKV + MF_QKV_BLK, *sH = base + MF_H;
  uint8_t *sU = base + MF_U, *sM = base + MF_U, *sAct = base + MF_U + 32768, *sCtx = base + MF_CTX;
  uint8_t *sWo = base + MF_WO, *sW1 = base + MF_W1, *sW2 = base + MF_W2, *sWq = base + MF_WQ, *sWp = base + MF_WP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + MF_BAR);
  uint64_t *b_wp = bars, *b_wo = bars + 1, *b_w1 = bars + 2, *b_w2 = bars + 3, *b_wq = bars + 4, *b_mma = bars + 5, *b_x = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float2* s_ln = reinterpret_cast<float2*>(base + MF_EX);
  float* s_mx = reinterpret_cast<float*>(base + MF_MX);
  float* s_sm = reinterpret_cast<float*>(base + MF_SM);
  float* s_sc = reinterpret_cast<float*>(base + MF_SC);
  float* s_prm = reinterpret_cast<float*>(base + MF_PRM);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  VB_TL(tl_mega_fwd, 0);
  // CTA pair: rank h runs attention head h; every other phase is computed by both CTAs (identical bits), rank 0 stores
  const int csz = P.cluster == 2 ? 2 : 1;
  const uint32_t crank = csz == 2 ? mf_cluster_rank() : 0u;
  const bool lead = crank == 0;
  const int hd_lo = csz == 2 ? (int)crank : 0, hd_hi = csz == 2 ? (int)crank + 1 : MG_NH;
  const bool is_side = warp >= 16;
  const int sw = warp - 16, sid = tid - MG_MAIN;   // side group: warp / lane index inside the group
  const bool s0 = sw == 0;                         // the side warp that publishes replicated vectors
  float* sp = reinterpret_cast<float*>(base + MF_SP);        // [4][32]
  float* spa = reinterpret_cast<float*>(base + MF_SP) + 128; // [4][20]: max, sum, o[16] partials of the side query
  const int r = ((warp & 3) << 5) | lane;    // main path: token row of the sample = TMEM lane
  const int cg = warp >> 2;                  // main path: column group
  const int hc0 = cg * HC;
  const int T = P.Np + 1, L = P.layers, B = P.B, C = P.C;
  const int Tm = T < 128 ? T : 128;          // rows on the tensor-core path
  const bool has_side = T > 128;             // token 128 runs on warp 16
  const int KP = (T + 15) & ~15, nch = KP >> 4;
  const size_t M = (size_t)B * T;
  const bool valid = !is_side && r < Tm;
  float* s_fin = s_prm + L * MG_PRM_LAYER;   // lnf_g[32] lnf_b[32] b_p[32] cls[32] w_h[C*32] b_h[C]

  if (tid == 0) {
    tma_prefetch_desc(&TM.wp); tma_prefetch_desc(&TM.wq); tma_prefetch_desc(&TM.wo); tma_prefetch_desc(&TM.w1);
    tma_prefetch_desc(&TM.w2); tma_prefetch_desc(&TM.z); tma_prefetch_desc(&TM.hmid); tma_prefetch_desc(&TM.u);
    tma_prefetch_desc(&TM.u2); tma_prefetch_desc(&TM.qkv); tma_prefetch_desc(&TM.ctx); tma_prefetch_desc(&TM.a);
    tma_prefetch_desc(&TM.m);
    mbar_init(b_wp, 1); mbar_init(b_wo, 1); mbar_init(b_w1, 1); mbar_init(b_w2, 1); mbar_init(b_wq, 1); mbar_init(b_mma, 1); mbar_init(b_x, 1);
    fence_barrier_init();
  }
  // key rows 128 .. 143 of the q|k and v blocks: zero once (row 128 is rewritten per layer when the sample has 129 tokens;
  // the others only have to be finite: their probabilities are exact zeros)
  if (tid < 256) *reinterpret_cast<uint4*>(base + MF_QKV + (tid >> 7) * MF_QKV_BLK + 16384 + (tid & 127) * 16) = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  // The previous kernel of the stream is the optimizer of the previous step: it writes every parameter, so nothing is
  // read before the dependency wait.
  pdl_wait();
  pdl_trigger();
  const bool reload = L > 1;   // one layer: the staged weights serve every sample
  auto load_wq = [&](int l) { mbar_expect_tx(b_wq, 12288); tma_load_3d(sWq, &TM.wq, b_wq, 0, 0, l); };
  auto load_wo = [&](int l) { mbar_expect_tx(b_wo, 4096); tma_load_3d(sWo, &TM.wo, b_wo, 0, 0, l); };
  auto load_w1 = [&](int l) { mbar_expect_tx(b_w1, 16384); tma_load_3d(sW1, &TM.w1, b_w1, 0, 0, l); };
  auto load_w2 = [&](int l) {
    mbar_expect_tx(b_w2, 8192);
    tma_load_3d(sW2, &TM.w2, b_w2, 0, 0, l);
    tma_load_3d(sW2 + 4096, &TM.w2, b_w2, 64, 0, l);
  };
  if (tid == 0) {
    mbar_expect_tx(b_wp, 4096);
    tma_load_2d(sWp, &TM.wp, b_wp, 0, 0);
    load_wq(0); load_wo(0); load_w1(0); load_w2(0);
  }
  if (warp == 0) tmem_alloc(tmem_slot, MC_COLS);
  for (int j = tid; j < L * MG_PRM_LAYER; j += MF_THREADS) {
    const int l = j / MG_PRM_LAYER, e = j - l * MG_PRM_LAYER;
    const float* lp = P.params + P.off_layer0 + (size_t)l * P.layer_stride;
    const float* src = e < MP_LN1B ? lp + P.o_ln1g + e : e < MP_BQ ? lp + P.o_ln1b + (e - MP_LN1B)
                     : e < MP_BO ? lp + P.o_bqkv + (e - MP_BQ) : e < MP_G2 ? lp + P.o_bo + (e - MP_BO)
                     : e < MP_B2LN ? lp + P.o_ln2g + (e - MP_G2) : e < MP_B1 ? lp + P.o_ln2b + (e - MP_B2LN)
                     : e < MP_B2 ? lp + P.o_b1 + (e - MP_B1) : lp + P.o_b2 + (e - MP_B2);
    s_prm[j] = *src;
  }
  if (tid < 32) {
    s_fin[tid] = P.params[P.off_lnfg + tid]; s_fin[32 + tid] = P.params[P.off_lnfb + tid];
    s_fin[64 + tid] = P.params[P.off_bp + tid]; s_fin[96 + tid] = P.params[P.off_cls + tid];
  }
  for (int j = tid; j < C * H + C; j += MF_THREADS)
    s_fin[128 + j] = j < C * H ? __bfloat162float(reinterpret_cast<const bf16*>(P.shadow)[P.off_wh + j]) : P.params[P.off_bh + (j - C * H)];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t my_tmem = tmem + ((uint32_t)((warp & 3) * 32) << 16);

  const uint32_t aA = smem_u32(sA), aQ0 = smem_u32(sQ0), aQ1 = smem_u32(sQ1), aU = smem_u32(sU);
  const MgOp A_k{aA, 16, 16384, 0};                    // H-wide A tile
  const MgOp CTX_k{smem_u32(sCtx), 16, 16384, 0};      // attention output tile
  const MgOp M_k{aU, 16, 16384, 0};                    // gelu tile: 2 k-blocks
  const MgOp P_k{aU, 16, 16384, 0};                    // P~ tile: up to 3 k-blocks
  const MgOp WP_k{smem_u32(sWp), 16, 0, 0}, WQ_k{smem_u32(sWq), 16, 0, 0}, WO_k{smem_u32(sWo), 16, 0, 0};
  const MgOp W1_k{smem_u32(sW1), 16, 0, 0}, W2_k{smem_u32(sW2), 16, 4096, 0};

  const uint64_t seed = P.rng ? P.rng[0] : 0ull;
  const uint32_t step = P.rng ? (uint32_t)P.rng[1] : 0u;
  const size_t pos = P.rows ? (size_t)(P.rng[1] - P.rows_base[0]) : 0;   // steps into the epoch's permutation
  const float scale = rsqrtf((float)MG_D), sl2 = scale * MG_LOG2E;
  const int Tpad = attn_drop_tpad(T);
  uint32_t ph_mma = 0, ph_x = 0;
  int it = 0;                 // layer iterations done (over samples): parity of the weight barriers
  float loss_acc = 0.f;       // lane 0 of warp 16: loss terms of this CTA's samples, in sample order

  VB_TL(tl_mega_fwd, 1);
  // CTA pair protocol, per layer: [wait B] push my head's attention output into both ctx tiles [sync A] ... hop 1 reads
  // the tile ... [arrive B] = "my tile may be overwritten".  B is split (arrive here, wait a whole layer later), so it
  // costs nothing; the first wait is matched by this arrive.
  if (csz == 2) mf_cluster_arrive();   // (release: also publishes the mbarrier initialisation to the peer)
  const int nsamp_par = (int)gridDim.x / csz;   // samples in flight
  for (int b = (int)blockIdx.x / csz; b < B; b += nsamp_par) {
    const bool more = b + nsamp_par < B;
    const size_t grow = (size_t)b * T + (is_side ? 128 : (valid ? r : 0));   // global token row (clamped for idle rows)
    float h[HC];            // main: residual row, columns hc0 .. hc0 + 7
    float zs = 0.f;         // side: residual row, column = lane
    float us = 0.f;         // side: LayerNorm output feeding the next GEMV (bf16-rounded)
    float qs = 0.f;         // side: q row (column = lane) of the current layer
    const size_t bsrc = P.rows ? (size_t)P.rows[pos * (size_t)B + b] : (size_t)b;   // dataset row of this sample

    // =============================== embedding ===============================
    if (!is_side) {
      // A rows = patch windows of tokens 1 .. (tokenization.py:45-48); the CLS row and padded windows are zero.
      const int nchunk = (P.P + 15) / 16 * 2;   // whole 16-element k-steps are read by the MMA
      const bool vec = (P.S % 4 == 0) && (P.L % 4 == 0) && (P.P % 8 == 0);
      for (int itx = tid; itx < 128 * nchunk; itx += MG_MAIN) {
        const int rr = itx / nchunk, c = itx - rr * nchunk;
        const bool has = rr >= 1 && rr < Tm && (rr - 1) < P.n_valid;
        const float* xp = P.x + bsrc * P.L + (size_t)(rr >= 1 ? rr - 1 : 0) * P.S;
        float v[8];
        if (vec && has && c * 8 + 8 <= P.P) {
          const float4 a0 = *reinterpret_cast<const float4*>(xp + c * 8), a1 = *reinterpret_cast<const float4*>(xp + c * 8 + 4);
          v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) { const int j = c * 8 + q; v[q] = (has && j < P.P) ? xp[j] : 0.f; }
        }
        *reinterpret_cast<uint4*>(mg_chunk(sA, rr, c)) = mg_pack8(v);
      }
    }
    float es = 0.f;   // side: patch projection of token 128 (before bias)
    if (is_side && has_side) {
      const bool has = 127 < P.n_valid;
      const float* xp = P.x + bsrc * P.L + (size_t)127 * P.S;
      const float x0 = (has && lane < P.P) ? bf16_round(xp[lane]) : 0.f;
      const float x1 = (has && lane + 32 < P.P) ? bf16_round(xp[lane + 32]) : 0.f;
      mbar_wait(b_wp, 0);
      for (int c = 0; c < P.P / 8; ++c) {
        float w[8];
        mg_unpack8(*reinterpret_cast<const uint4*>(mg_chunk(sWp, lane, c)), w);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int k = c * 8 + q;
          es = fmaf(__shfl_sync(0xffffffffu, k < 32 ? x0 : x1, k & 31), w[q], es);
        }
      }
    }
    fence_proxy_async();
    tc_fence_before();
    if (tid == 0) tma_store_wait_read<0>();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(b_wp, 0);
      mg_issue(tmem + MC_LIN, A_k, WP_k, H, (P.P + 15) / 16, false);
      umma_commit(b_mma);
    }
    {
      const DropCtx dc = make_drop(P.p_hidden, seed, step, VITB200_SITE_EMB);
      const float* lp0 = s_prm;   // layer 0
      if (!is_side) {
        mbar_wait(b_mma, ph_mma); ph_mma ^= 1;
        tc_fence_after();
        float v[HC], kp[8];
        tmem_ld_32x8(my_tmem + MC_LIN + hc0, v);
        drop8(dc, (grow * H + hc0) >> 3, kp);
#pragma unroll
        for (int q = 0; q < HC; ++q) {
          const int c = hc0 + q;
          float e = r == 0 ? s_fin[96 + c] : bf16_round(v[q] + s_fin[64 + c]);
          if (P.off_pos >= 0 && valid) e += P.params[P.off_pos + (size_t)r * H + c];
          h[q] = e * kp[q];
        }
        float mu, rs;
        mf_row_stats(h, s_ln, r, cg, P.eps, mu, rs);
        float u[HC];
#pragma unroll
        for (int j = 0; j < HC; ++j) u[j] = bf16_round((h[j] - mu) * rs * lp0[MP_LN1G + hc0 + j] + lp0[MP_LN1B + hc0 + j]);
        *mg_f32(sH, r, hc0) = make_float4(h[0], h[1], h[2], h[3]);
        *mg_f32(sH, r, hc0 + 4) = make_float4(h[4], h[5], h[6], h[7]);
        if (lead && valid && cg == 0) { P.stats[grow] = mu; P.stats[M + grow] = rs; }
        *reinterpret_cast<uint4*>(mg_chunk(sA, r, cg)) = mg_pack8(u);
        if (r == 0) {   // CLS row of the residual stream: the side warp takes it over in a CLS-only last layer
#pragma unroll
          for (int j = 0; j < HC; ++j) s_sc[64 + hc0 + j] = h[j];
        }
      } else if (has_side) {
        float e = bf16_round(es + s_fin[64 + lane]);
        if (P.off_pos >= 0) e += P.params[P.off_pos + (size_t)128 * H + lane];
        zs = e * drop1(dc, grow * H + lane);
        float mu, rs;
        mg_side_ln(zs, P.eps, mu, rs);
        us = bf16_round((zs - mu) * rs * lp0[MP_LN1G + lane] + lp0[MP_LN1B + lane]);
        if (lead && s0) {
          P.z[grow * H + lane] = zs;
          reinterpret_cast<bf16*>(P.u)[grow * H + lane] = __float2bfloat16_rn(us);
          if (lane == 0) { P.stats[grow] = mu; P.stats[M + grow] = rs; }
        }
      }
    }

    // =============================== layers ===============================
    for (int l = 0; l <= L; ++l) {
      // ---- QKV projection of layer l (l == L: head instead) from the LayerNorm output sitting in sA ----
      fence_proxy_async();
      tc_fence_before();
      if (tid == 0) tma_store_wait_read<0>();
      __syncthreads();   // closes the embedding phase / hop 3 of layer l - 1
      // W2 of the previous layer is dead now (its GEMM completed, the side warp has read it): stage the next one
      if (tid == 0 && reload && l > 0 && (l < L || more)) load_w2(l < L ? l : 0);
      const bool last = l == L;
      if (last) {
        // residual stream after the last layer; the final LayerNorm of the CLS row was written to s_sc in hop 3
        if (tid == 0 && lead && !P.cls_only) {
          tma_store_4d(&TM.z, sH, 0, 0, b, L);
          tma_store_commit();
        }
        if (is_side && s0 && lead) {
          // head (specvit.py:81-89): logits = s . Wh^T + bh; loss term of the sample
          float mine[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c < C) {
              const float lg = mg_wsum(s_sc[lane] * s_fin[128 + c * H + lane]);
              mine[c] = bf16_round(lg + s_fin[128 + C * H + c]);
            }
          }
          if (lane == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) if (c < C) P.logits[(size_t)b * C + c] = mine[c];
            if (P.labels) loss_acc += mf_loss_term(mine, P.labels, (int)bsrc, C, P.loss_kind);
          }
        }
        break;
      }
      const float* lp = s_prm + l * MG_PRM_LAYER;
      const uint32_t wpar = reload ? (uint32_t)(it & 1) : 0u;
      // CLS-only last layer (training steps, logits-only evaluation): the head reads nothing but last_hidden_state[:, 0]
      // (specvit.py:78), so in the last layer only the CLS row needs attention output, MLP and final LayerNorm.  Its
      // keys / values are still every token's (the QKV projection above runs for all rows); the row itself (tile row 0)
      // is carried by the side warp with FMAs while the tensor-core rows idle.
      const bool ct = P.cls_only != 0 && l == L - 1;
      const bool s_on = has_side || ct;                 // the side warp carries a row in this layer
      const int stok = ct ? 0 : 128;                    // ... this token
      const size_t sgrow = (size_t)b * T + stok;
      if (tid == 0) {
        tc_fence_after();
        mbar_wait(b_wq, wpar);
        mg_issue(tmem + MC_LIN, A_k, WQ_k, MG_Q, H / 16, false);
        umma_commit(b_mma);
        if (lead) {
          tma_store_4d(&TM.z, sH, 0, 0, b, l);      // residual stream entering layer l
          tma_store_4d(&TM.u, sA, 0, 0, b, l);
          tma_store_commit();
        }
      }
      float ks = 0.f, vs = 0.f;   // side: k / v rows (column = lane)
      if (!is_side) {
        mbar_wait(b_mma, ph_mma); ph_mma ^= 1;
        tc_fence_after();
        // 96 columns = 3 groups of 32: group g goes to cg = g, cg 3 idles (the epilogue is a few hundred cycles)
        if (cg < 3) {
          float v[32];
          tmem_ld_32x32(my_tmem + MC_LIN + cg * 32, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += lp[MP_BQ + cg * 32 + j];
          uint8_t* blk = cg < 2 ? sQ0 : sQ1;
#pragma unroll
          for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(mg_chunk(blk, r, (cg & 1) * 4 + j)) = mg_pack8(&v[8 * j]);
        }
      } else if (has_side) {
        mbar_wait(b_wq, wpar);
        float y[3];
        mg_side_gemv32<3>(sWq, lane, us, y);
        qs = bf16_round(y[0] + lp[MP_BQ + lane]);
        ks = bf16_round(y[1] + lp[MP_BQ + 32 + lane]);
        vs = bf16_round(y[2] + lp[MP_BQ + 64 + lane]);
        if (lead && s0) {
          bf16* qrow = reinterpret_cast<bf16*>(P.qkv) + ((size_t)l * M + grow) * MG_Q;
          qrow[lane] = __float2bfloat16_rn(qs); qrow[32 + lane] = __float2bfloat16_rn(ks); qrow[64 + lane] = __float2bfloat16_rn(vs);
        }
        if (P.rope_cos) {   // rotate q and k of the side row (rope.py:60-98): partner column = lane ^ 8 inside a head
          const int c = lane & 7;
          const float cs = P.rope_cos[(size_t)128 * 8 + c], sn = P.rope_sin[(size_t)128 * 8 + c];
          const float qo = __shfl_xor_sync(0xffffffffu, qs, 8), ko = __shfl_xor_sync(0xffffffffu, ks, 8);
          qs = bf16_round((lane & 8) ? qs * cs + qo * sn : qs * cs - qo * sn);
          ks = bf16_round((lane & 8) ? ks * cs + ko * sn : ks * cs - ko * sn);
        }
        if (s0) {
          *mg_elem(sQ0, 128, 32 + lane) = __float2bfloat16_rn(ks);
          *mg_elem(sQ1, 128, lane) = __float2bfloat16_rn(vs);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      if (tid == 0) tma_store_wait_read<0>();
      __syncthreads();
      if (l == 0) VB_TL(tl_mega_fwd, 2);
      if (l == 0) VB_TL_T(tl_mega_fwd, 18, MG_MAIN);
      auto issue_scores = [&]() {   // thread 0: S[i, j] = q_i . k_j for this CTA's head(s)  (HF:232-249 / vit_with_rope.py:43-84)
        tc_fence_after();
        for (int hd = hd_lo; hd < hd_hi; ++hd) {
          const MgOp Qh{aQ0 + hd * 32, 16, 0, 0}, Kh{aQ0 + 64 + hd * 32, 16, 0, 0};
          mg_issue(tmem + (hd ? MC_S1 : MC_S0), Qh, Kh, KP, 1, false);
        }
        umma_commit(b_mma);
      };
      auto store_qkv = [&]() {      // thread 0: the q|k|v rows leave for HBM; the next layer's Wqkv is staged
        if (lead) {
          tma_store_4d(&TM.qkv, sQ0, 0, 0, b, l);
          tma_store_4d(&TM.qkv, sQ1, 64, 0, b, l);
          tma_store_commit();
        }
        const int ln = l + 1 < L ? l + 1 : 0;   // the next QKV event: layer l + 1, or layer 0 of the next sample
        if (reload && (l + 1 < L || more)) load_wq(ln);
      };
      if (!P.rope_cos) {
        if (tid == 0) { if (!ct) issue_scores(); store_qkv(); }
      } else {
        // q / k of the tensor-core rows are rotated in place AFTER the un-rotated rows left for HBM (backward rotates again)
        if (tid == 0) { store_qkv(); tma_store_wait_read<0>(); }
        __syncthreads();
        if (!is_side) {
          // cg 0: q head 0, cg 1: q head 1, cg 2: k head 0, cg 3: k head 1  (16 columns = 2 chunks each)
          float x[16];
          mg_unpack8(*reinterpret_cast<const uint4*>(mg_chunk(sQ0, r, 2 * cg)), &x[0]);
          mg_unpack8(*reinterpret_cast<const uint4*>(mg_chunk(sQ0, r, 2 * cg + 1)), &x[8]);
          const int t = valid ? r : 0;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float cs = P.rope_cos[(size_t)t * 8 + c], sn = P.rope_sin[(size_t)t * 8 + c];
            const float lo = x[c], hi = x[c + 8];
            x[c] = lo * cs - hi * sn;
            x[c + 8] = hi * cs + lo * sn;
          }
          *reinterpret_cast<uint4*>(mg_chunk(sQ0, r, 2 * cg)) = mg_pack8(&x[0]);
          *reinterpret_cast<uint4*>(mg_chunk(sQ0, r, 2 * cg + 1)) = mg_pack8(&x[8]);
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0 && !ct) issue_scores();
      }

      // ---- attention ----
      const DropCtx dca = make_drop(P.p_attn, seed, step, VITB200_SITE_ATTN(l));
      // bytes the peer pushes into this CTA per layer: its head's ctx columns of 128 rows (16 B each, two column groups)
      // and the 16 ctx values of its head for the side row
      if (csz == 2 && tid == 0) mbar_expect_tx(b_x, (ct ? 0u : 4096u) + (s_on ? 64u : 0u));
      const uint32_t x_peer = csz == 2 ? mf_peer_addr(b_x, crank ^ 1u) : 0u;
      float cs_ = 0.f;   // side: attention output row (column = lane)
      if (is_side) {
        if (csz == 2) mf_cluster_wait();   // the peer has finished reading last layer's exchange buffers
        if (ct) {   // take over the CLS row: its residual row and its (rotated) query row, tile row 0
          zs = s_sc[64 + lane];
          qs = __bfloat162float(*mg_elem(sQ0, 0, lane));
        }
        if (s_on) {
          for (int hd = hd_lo; hd < hd_hi; ++hd) {
            float qf[MG_D];
#pragma unroll
            for (int c = 0; c < MG_D; ++c) qf[c] = __shfl_sync(0xffffffffu, qs, hd * MG_D + c);
            // key j = sid (one key per lane of the group); key 128 -- the side row -- is lane 0's second key
            float sc[2];
            float mx = -INFINITY;
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              const int j = sid + 128 * jj;
              float acc = -INFINITY;
              if (j < T && (jj == 0 || sid == 0)) {
                float kr[MG_D];
                mg_unpack8(*reinterpret_cast<const uint4*>(mg_chunk(sQ0, j, 4 + 2 * hd)), &kr[0]);
                mg_unpack8(*reinterpret_cast<const uint4*>(mg_chunk(sQ0, j, 5 + 2 * hd)), &kr[8]);
                acc = 0.f;
#pragma unroll
                for (int c = 0; c < MG_D; ++c) acc = fmaf(qf[c], kr[c], acc);
              }
              sc[jj] = acc;
              mx = fmaxf(mx, acc);
            }
            mx = mg_wmax(mx);
            if (lane == 0) spa[sw * 20] = mx;
            mf_bar_side();
            mx = fmaxf(fmaxf(spa[0], spa[20]), fmaxf(spa[40], spa[60]));
            const uint64_t drow = ((uint64_t)(b * MG_NH + hd) * T + stok) * (uint64_t)Tpad;
            float sum = 0.f, o[MG_D];
#pragma unroll
            for (int c = 0; c < MG_D; ++c) o[c] = 0.f;
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              const int j = sid + 128 * jj;
              if (j < T && (jj == 0 || sid == 0)) {
                float p = exp2f((sc[jj] - mx) * sl2);
                sum += p;
                p = bf16_round(p * drop1(dca, drow + (uint64_t)j));
                float vr[MG_D];
                mg_unpack8(*reinterpret_cast<const uint4*>(mg_chunk(sQ1, j, 2 * hd)), &vr[0]);
                mg_unpack8(*reinterpret_cast<const uint4*>(mg_chunk(sQ1, j, 2 * hd + 1)), &vr[8]);
#pragma unroll
                for (int c = 0; c < MG_D; ++c) o[c] = fmaf(p, vr[c], o[c]);
              }
            }
            // warp sums, then the four warps' partials in warp order
            sum = mg_wsum(sum);
            if (lane == 0) spa[sw * 20 + 1] = sum;
#pragma unroll
            for (int c = 0; c < MG_D; ++c) {
              const float t = mg_wsum(o[c]);
              if (lane == c) spa[sw * 20 + 2 + c] = t;
            }
            mf_bar_side();
            sum = (spa[1] + spa[21]) + (spa[41] + spa[61]);
            const float inv = 1.f / sum;
            if ((lane >> 4) == hd) {
              const int c = lane & 15;
              cs_ = bf16_round(((spa[2 + c] + spa[22 + c]) + (spa[42 + c] + spa[62 + c])) * inv);
            }
            if (lane == 0 && s0) P.lse[(((size_t)l * B + b) * MG_NH + hd) * T + stok] = mx * scale + logf(sum);
            mf_bar_side();   // spa is rewritten by the next head
          }
          if (s0 && (lane >> 4) >= hd_lo && (lane >> 4) < hd_hi) {   // this CTA's head(s)
            reinterpret_cast<bf16*>(P.ctx)[((size_t)l * M + sgrow) * H + lane] = __float2bfloat16_rn(cs_);
            s_sc[32 + lane] = cs_;
            if (csz == 2) mf_st_async4(mf_peer_addr(&s_sc[32 + lane], crank ^ 1u), cs_, x_peer);
          }
        }
      } else if (ct) {
        if (csz == 2) mf_cluster_wait();
      } else {
        float mxh[MG_NH];
        for (int hd = hd_lo; hd < hd_hi; ++hd) {
          const uint32_t cS = hd ? MC_S1 : MC_S0;
          if (hd == hd_lo) { mbar_wait(b_mma, ph_mma); ph_mma ^= 1; tc_fence_after(); if (l == 0) VB_TL(tl_mega_fwd, 3); }
          // -- row maximum: this thread reads the 16-key chunks cg, cg + 4, cg + 8 of its query row --
          float mx = -INFINITY;
          for (int ch = cg; ch < nch; ch += MG_CG) {
            const int c0 = ch * 16;
            if (T - c0 == 1) {   // one live key in the last chunk (T = 16 n + 1: the patches + CLS): one element, not a masked chunk
              float v8[8];
              tmem_ld_32x8(my_tmem + cS + c0, v8);
              mx = fmaxf(mx, v8[0]);
              continue;
            }
            float v[16];
            tmem_ld_32x16(my_tmem + cS + c0, v);
            if (c0 + 16 <= T) {
#pragma unroll
              for (int j = 0; j < 16; ++j) mx = fmaxf(mx, v[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) if (c0 + j < T) mx = fmaxf(mx, v[j]);
            }
          }
          s_mx[cg * 128 + r] = mx;
          mg_bar_main();
          mx = fmaxf(fmaxf(s_mx[r], s_mx[128 + r]), fmaxf(s_mx[256 + r], s_mx[384 + r]));
          mxh[hd] = mx;
          const float nmxs = -mx * sl2;
          if (l == 0 && hd == hd_lo) VB_TL(tl_mega_fwd, 4);
          if (hd > hd_lo) { mbar_wait(b_mma, ph_mma); ph_mma ^= 1; }   // O_0 = P~_0 V_0 has finished reading the P~ tile
          // -- probabilities (the S chunks are read from TMEM a second time instead of being kept in registers) --
          float sum = 0.f;
          const uint64_t drow = ((uint64_t)(b * MG_NH + hd) * T + (valid ? r : 0)) * (uint64_t)Tpad;
          for (int ch = cg; ch < nch; ch += MG_CG) {
            const int c0 = ch * 16;
            if (T - c0 == 1) {   // one live key in the last chunk: a single exponential and mask, zeros for the padding keys
              float v8[8], kp[8];
              tmem_ld_32x8(my_tmem + cS + c0, v8);
              drop8(dca, (drow + (uint64_t)c0) >> 3, kp);
              const float p = ex2_approx(fmaf(v8[0], sl2, nmxs));
              sum += p;
              v8[0] = p * kp[0];
#pragma unroll
              for (int q = 1; q < 8; ++q) v8[q] = 0.f;
              *reinterpret_cast<uint4*>(mg_swz(sU, r, c0 >> 3)) = mg_pack8(v8);
              *reinterpret_cast<uint4*>(mg_swz(sU, r, (c0 >> 3) + 1)) = make_uint4(0u, 0u, 0u, 0u);
              continue;
            }
            float v[16];
            tmem_ld_32x16(my_tmem + cS + c0, v);
            if (c0 + 16 <= T) {
#pragma unroll
              for (int j = 0; j < 16; j += 8) {
                float kp[8];
                drop8(dca, (drow + (uint64_t)(c0 + j)) >> 3, kp);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const float p = ex2_approx(fmaf(v[j + q], sl2, nmxs));
                  sum += p;
                  v[j + q] = p * kp[q];
                }
                *reinterpret_cast<uint4*>(mg_swz(sU, r, (c0 + j) >> 3)) = mg_pack8(&v[j]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; j += 8) {
                float kp[8];
                if (c0 + j < T) drop8(dca, (drow + (uint64_t)(c0 + j)) >> 3, kp);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const float p = (c0 + j + q < T) ? ex2_approx(fmaf(v[j + q], sl2, nmxs)) : 0.f;
                  sum += p;
                  v[j + q] = (c0 + j < T) ? p * kp[q] : 0.f;
                }
                *reinterpret_cast<uint4*>(mg_swz(sU, r, (c0 + j) >> 3)) = mg_pack8(&v[j]);
              }
            }
          }
          s_sm[(hd * MG_CG + cg) * 128 + r] = sum;
          fence_proxy_async();
          tc_fence_before();
          mg_bar_main();
          if (l == 0 && hd == hd_lo) VB_TL(tl_mega_fwd, 5);
          if (tid == 0) {
            tc_fence_after();
            const MgOp Vh{aQ1 + hd * 32, 16384, 0, 1};
            mg_issue(tmem + MC_O + hd * MG_D, P_k, Vh, MG_D, KP / 16, false);   // O[i, c] = sum_j P~[i, j] v[j, c]
            umma_commit(b_mma);
          }
        }
        mbar_wait(b_mma, ph_mma); ph_mma ^= 1;
        tc_fence_after();
        if (csz == 2) mf_cluster_wait();   // the peer has finished reading last layer's ctx tile
        if (l == 0) VB_TL(tl_mega_fwd, 6);
        // ctx columns 8 c .. 8 c + 7 (chunk c) belong to head c >> 1.  One CTA: thread group cg takes chunk cg; CTA pair:
        // groups 0, 1 take the two chunks of this CTA's head and write them into both CTAs' tiles.
        const int chunk = csz == 2 ? 2 * (int)crank + cg : cg;
        if (csz == 1 || cg < 2) {
          const int hd = chunk >> 1;
          const float* ss = s_sm + hd * MG_CG * 128;
          const float sum = (ss[r] + ss[128 + r]) + (ss[256 + r] + ss[384 + r]);
          float o[8];
          tmem_ld_32x8(my_tmem + MC_O + chunk * 8, o);
          const float inv = 1.f / sum;
#pragma unroll
          for (int c = 0; c < 8; ++c) o[c] *= inv;
          const uint4 pk = mg_pack8(o);
          uint8_t* dst = mg_chunk(sCtx, r, chunk);
          *reinterpret_cast<uint4*>(dst) = pk;
          if (csz == 2) mf_st_async16(mf_peer_addr(dst, crank ^ 1u), pk, x_peer);
          if (valid && (chunk & 1) == 0) P.lse[(((size_t)l * B + b) * MG_NH + hd) * T + r] = mxh[hd] * scale + logf(sum);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      if (tid == 0) tma_store_wait_read<0>();
      if (l == 0) { VB_TL(tl_mega_fwd, 27); VB_TL_T(tl_mega_fwd, 28, MG_MAIN); }   // (debug: own work of the attention phase done)
      if (csz == 2) { mbar_wait(b_x, ph_x); ph_x ^= 1; }   // the peer's pushes have landed
      __syncthreads();                                     // (and this CTA's own writes)
      if (csz == 2) fence_proxy_async();
      if (is_side && s_on) cs_ = s_sc[32 + lane];   // the full attention output row (both heads)
      if (l == 0) VB_TL(tl_mega_fwd, 7);
      if (l == 0) VB_TL_T(tl_mega_fwd, 23, MG_MAIN);

      // ---- hop 1: attention output projection + dropout + residual, LayerNorm-after (HF:262-268,337-340) ----
      if (tid == 0) {
        tc_fence_after();
        mbar_wait(b_wo, wpar);
        if (!ct) {
          mg_issue(tmem + MC_LIN, CTX_k, WO_k, H, H / 16, false);
          umma_commit(b_mma);
          if (lead) {
            tma_store_4d(&TM.ctx, sCtx, 0, 0, b, l);
            tma_store_commit();
          }
        }
      }
      {
        const DropCtx dc = make_drop(P.p_hidden, seed, step, VITB200_SITE_PROJ(l));
        if (!is_side && !ct) {
          mbar_wait(b_mma, ph_mma); ph_mma ^= 1;
          tc_fence_after();
          float v[HC], kp[8];
          tmem_ld_32x8(my_tmem + MC_LIN + hc0, v);
          drop8(dc, (grow * H + hc0) >> 3, kp);
#pragma unroll
          for (int e = 0; e < HC; ++e) h[e] += bf16_round(bf16_round(v[e] + lp[MP_BO + hc0 + e]) * kp[e]);
          float mu, rs;
          mf_row_stats(h, s_ln, r, cg, P.eps, mu, rs);
          float u2[HC];
#pragma unroll
          for (int j = 0; j < HC; ++j) u2[j] = bf16_round((h[j] - mu) * rs * lp[MP_G2 + hc0 + j] + lp[MP_B2LN + hc0 + j]);
          *mg_f32(sH, r, hc0) = make_float4(h[0], h[1], h[2], h[3]);
          *mg_f32(sH, r, hc0 + 4) = make_float4(h[4], h[5], h[6], h[7]);
          if (lead && valid && cg == 0) { P.stats[(size_t)(4 * l + 2) * M + grow] = mu; P.stats[(size_t)(4 * l + 3) * M + grow] = rs; }
          *reinterpret_cast<uint4*>(mg_chunk(sA, r, cg)) = mg_pack8(u2);
        } else if (is_side && s_on) {
          mbar_wait(b_wo, wpar);
          float y[1];
          mg_side_gemv32<1>(sWo, lane, cs_, y);
          zs += bf16_round(bf16_round(y[0] + lp[MP_BO + lane]) * drop1(dc, sgrow * H + lane));
          float mu, rs;
          mg_side_ln(zs, P.eps, mu, rs);
          us = bf16_round((zs - mu) * rs * lp[MP_G2 + lane] + lp[MP_B2LN + lane]);
          if (lead && s0) {
            P.hmid[((size_t)l * M + sgrow) * H + lane] = zs;
            reinterpret_cast<bf16*>(P.u2)[((size_t)l * M + sgrow) * H + lane] = __float2bfloat16_rn(us);
            if (lane == 0) { P.stats[(size_t)(4 * l + 2) * M + sgrow] = mu; P.stats[(size_t)(4 * l + 3) * M + sgrow] = rs; }
          }
        }
      }
      fence_proxy_async();
      tc_fence_before();
      if (tid == 0) tma_store_wait_read<0>();
      __syncthreads();
      if (csz == 2) mf_cluster_arrive_relaxed();   // hop 1 (and the ctx store) have consumed the exchange buffers
      if (l == 0) VB_TL(tl_mega_fwd, 8);
      if (l == 0) VB_TL_T(tl_mega_fwd, 24, MG_MAIN);

      // ---- hop 2: MLP up + GELU (HF:296-299) ----
      if (tid == 0) {
        tc_fence_after();
        if (reload && (l + 1 < L || more)) load_wo(l + 1 < L ? l + 1 : 0);
        mbar_wait(b_w1, wpar);
        if (!ct) {
          mg_issue(tmem + MC_LIN, A_k, W1_k, I, H / 16, false);
          umma_commit(b_mma);
          if (lead) {
            tma_store_4d(&TM.hmid, sH, 0, 0, b, l);
            tma_store_4d(&TM.u2, sA, 0, 0, b, l);
            tma_store_commit();
          }
        }
      }
      float ms = 0.f;   // side: gelu output, column sid
      if (!is_side && !ct) {
        mbar_wait(b_mma, ph_mma); ph_mma ^= 1;
        tc_fence_after();
        const int c0 = cg * 32;
        float v[32];
        tmem_ld_32x32(my_tmem + MC_LIN + c0, v);
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          // pre-activation rounded to bf16 once (packed), GELU evaluated on the rounded value, packed again
          uint32_t wa[4], wm[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 pa = __floats2bfloat162_rn(v[j + 2 * e] + lp[MP_B1 + c0 + j + 2 * e], v[j + 2 * e + 1] + lp[MP_B1 + c0 + j + 2 * e + 1]);
            const float2 f = __bfloat1622float2(pa);
            const __nv_bfloat162 pm = __floats2bfloat162_rn(gelu_f(f.x), gelu_f(f.y));
            wa[e] = *reinterpret_cast<const uint32_t*>(&pa);
            wm[e] = *reinterpret_cast<const uint32_t*>(&pm);
          }
          *reinterpret_cast<uint4*>(mg_swz(sAct, r, (c0 + j) >> 3)) = make_uint4(wa[0], wa[1], wa[2], wa[3]);
          *reinterpret_cast<uint4*>(mg_swz(sM, r, (c0 + j) >> 3)) = make_uint4(wm[0], wm[1], wm[2], wm[3]);
        }
      } else if (is_side && s_on) {
        mbar_wait(b_w1, wpar);
        float y[1];
        mg_side_gemv32<1>(sW1, sid, us, y);   // MLP column sid (us is replicated in every warp)
        const float a = bf16_round(y[0] + lp[MP_B1 + sid]);
        ms = bf16_round(gelu_f(a));
        if (lead) {
          reinterpret_cast<bf16*>(P.a)[((size_t)l * M + sgrow) * I + sid] = __float2bfloat16_rn(a);
          reinterpret_cast<bf16*>(P.m)[((size_t)l * M + sgrow) * I + sid] = __float2bfloat16_rn(ms);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      if (tid == 0) tma_store_wait_read<0>();
      __syncthreads();
      if (l == 0) VB_TL(tl_mega_fwd, 9);
      if (l == 0) VB_TL_T(tl_mega_fwd, 25, MG_MAIN);

      // ---- hop 3: MLP down + dropout + residual, then LN1 of the next layer / the final LN of the CLS row ----
      if (tid == 0) {
        tc_fence_after();
        if (reload && (l + 1 < L || more)) load_w1(l + 1 < L ? l + 1 : 0);
        mbar_wait(b_w2, wpar);
        if (!ct) {
          mg_issue(tmem + MC_LIN, M_k, W2_k, H, I / 16, false);
          umma_commit(b_mma);
          if (lead) {
            tma_store_4d(&TM.a, sAct, 0, 0, b, l);
            tma_store_4d(&TM.a, sAct + 16384, 64, 0, b, l);
            tma_store_4d(&TM.m, sM, 0, 0, b, l);
            tma_store_4d(&TM.m, sM + 16384, 64, 0, b, l);
            tma_store_commit();
          }
        }
      }
      {
        const DropCtx dc = make_drop(P.p_hidden, seed, step, VITB200_SITE_MLP(l));
        const bool top = l + 1 == L;
        const float* gn = top ? s_fin : s_prm + (l + 1) * MG_PRM_LAYER + MP_LN1G;
        const float* bn = top ? s_fin + 32 : s_prm + (l + 1) * MG_PRM_LAYER + MP_LN1B;
        if (!is_side && !ct) {
          mbar_wait(b_mma, ph_mma); ph_mma ^= 1;
          tc_fence_after();
          float v[HC], kp[8];
          tmem_ld_32x8(my_tmem + MC_LIN + hc0, v);
          drop8(dc, (grow * H + hc0) >> 3, kp);
#pragma unroll
          for (int e = 0; e < HC; ++e) h[e] += bf16_round(bf16_round(v[e] + lp[MP_B2 + hc0 + e]) * kp[e]);
          *mg_f32(sH, r, hc0) = make_float4(h[0], h[1], h[2], h[3]);
          *mg_f32(sH, r, hc0 + 4) = make_float4(h[4], h[5], h[6], h[7]);
          if (r == 0) {
#pragma unroll
            for (int j = 0; j < HC; ++j) s_sc[64 + hc0 + j] = h[j];
          }
          float mu, rs;
          mf_row_stats(h, s_ln, r, cg, P.eps, mu, rs);
          float un[HC];
#pragma unroll
          for (int j = 0; j < HC; ++j) un[j] = bf16_round((h[j] - mu) * rs * gn[hc0 + j] + bn[hc0 + j]);
          if (!top) {
            if (lead && valid && cg == 0) { P.stats[(size_t)(4 * l + 4) * M + grow] = mu; P.stats[(size_t)(4 * l + 5) * M + grow] = rs; }
            *reinterpret_cast<uint4*>(mg_chunk(sA, r, cg)) = mg_pack8(un);
          } else if (r == 0) {   // CLS row (specvit.py:78): final LayerNorm, handed to the head
            if (lead && cg == 0) { P.stats[(size_t)(4 * L) * M + b] = mu; P.stats[(size_t)(4 * L + 1) * M + b] = rs; }
            if (lead) *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(P.s_cls) + (size_t)b * H + hc0) = mg_pack8(un);
#pragma unroll
            for (int j = 0; j < HC; ++j) s_sc[hc0 + j] = un[j];
          }
        } else if (is_side && s_on) {
          mbar_wait(b_w2, wpar);
          // warp w contracts columns 32 w .. 32 w + 31 (k-block w >> 1, chunks 4 (w & 1) ..) of output row n = lane
          float part = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float w[8];
            mg_unpack8(*reinterpret_cast<const uint4*>(mg_chunk(sW2 + (sw >> 1) * 4096, lane, 4 * (sw & 1) + c)), w);
#pragma unroll
            for (int q = 0; q < 8; ++q) part = fmaf(__shfl_sync(0xffffffffu, ms, 8 * c + q), w[q], part);
          }
          sp[sw * 32 + lane] = part;
          mf_bar_side();
          const float y = (sp[lane] + sp[32 + lane]) + (sp[64 + lane] + sp[96 + lane]);
          zs += bf16_round(bf16_round(y + lp[MP_B2 + lane]) * drop1(dc, sgrow * H + lane));
          if (lead && s0) P.z[((size_t)(l + 1) * M + sgrow) * H + lane] = zs;
          if (ct) {   // final LayerNorm of the CLS row (HF:455), handed to the head
            float mu, rs;
            mg_side_ln(zs, P.eps, mu, rs);
            const float un = bf16_round((zs - mu) * rs * gn[lane] + bn[lane]);
            if (s0) s_sc[lane] = un;
            if (lead && s0) {
              reinterpret_cast<bf16*>(P.s_cls)[(size_t)b * H + lane] = __float2bfloat16_rn(un);
              if (lane == 0) { P.stats[(size_t)(4 * L) * M + b] = mu; P.stats[(size_t)(4 * L + 1) * M + b] = rs; }
            }
          } else if (!top) {
            float mu, rs;
            mg_side_ln(zs, P.eps, mu, rs);
            us = bf16_round((zs - mu) * rs * gn[lane] + bn[lane]);
            if (lead && s0) {
              reinterpret_cast<bf16*>(P.u)[((size_t)(l + 1) * M + grow) * H + lane] = __float2bfloat16_rn(us);
              if (lane == 0) { P.stats[(size_t)(4 * l + 4) * M + grow] = mu; P.stats[(size_t)(4 * l + 5) * M + grow] = rs; }
            }
          }
        }
      }
      if (l == 0) VB_TL(tl_mega_fwd, 10);
      if (l == 0) VB_TL_T(tl_mega_fwd, 26, MG_MAIN);
      ++it;   // (the barrier that closes hop 3 is the one at the top of the loop)
    }
  }

  // ---- loss: CTA partials summed by the last CTA, in CTA order ----
  VB_TL(tl_mega_fwd, 11);
  // no CTA of a pair exits while its peer may still touch its shared memory (nothing is published: relaxed arrive)
  if (csz == 2) { mf_cluster_wait(); mf_cluster_arrive_relaxed(); mf_cluster_wait(); }
  float* loss_part = reinterpret_cast<float*>(reinterpret_cast<char*>(P.ws) + 256);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(P.ws);
  if (tid == MG_MAIN) loss_part[blockIdx.x] = loss_acc;   // lane 0 of side warp 0
  if (tid == 0) tma_store_wait_read<0>();                 // (shared memory may be released; the stores complete with the grid)
  tc_fence_before();
  // defer_loss: the backward kernel of the same step (vitb200_mega_bwd) sums the partials -- no fence / ticket here
  if (!P.defer_loss && last_block_ticket(ticket, gridDim.x) && P.labels && warp == 0) {
    // fixed-order sum of the CTA partials: lane l adds partials l, l + 32, ... (all loads in flight at once), then a
    // shuffle tree -- the serial loop over 128 dependent L2 reads used to hold the next kernel back by ~4 us
    const float t = mg_loss_sum(loss_part, gridDim.x, lane);
    if (lane == 0) {
      const float v = t / (P.loss_kind == VITB200_LOSS_CE ? (float)B : (float)B * (float)C);
      P.loss[0] = v;
      if (P.loss_log) P.loss_log[pos] = v;
    }
  }
  if (P.defer_loss) __syncthreads();   // (the ticket's barrier otherwise: TMEM is released after every warp is done with it)
  if (warp == 0) tmem_dealloc(tmem, MC_COLS);
  VB_TL(tl_mega_fwd, 12);
}

}  // namespace vb

using namespace vb;

VB_TL_EXPORT(vitb200_tl_mega_fwd, vb::tl_mega_fwd)

extern "C" int vitb200_mega_supported(int H, int heads, int T, int P, int C, int layers, int rope) {
  (void)rope;
  if (H != MG_H || heads != MG_NH) return 0;
  if (T < 2 || T > 129) return 0;
  if (P % 8 != 0 || P < 8 || P > 64) return 0;
  if (C < 1 || C > 4) return 0;
  if (layers < 1 || layers > 16) return 0;
  return 1;
}
extern "C" size_t vitb200_mega_ws_bytes(void) { return 256 + 160 * sizeof(float); }
extern "C" size_t vitb200_mega_fwd_smem_bytes(int layers) {
  return 1024 + MF_PRM + (size_t)(layers * MG_PRM_LAYER + 128 + 4 * MG_H + 4 + 32) * sizeof(float);
}
extern "C" int vitb200_mega_grid(int B, int cluster) {
  const int per = cluster == 2 ? 2 : 1;
  int g = B * per;
  if (g > 148) g = 148 / per * per;
  return g < per ? per : g;
}

extern "C" int vitb200_mega_fwd(const vitb200_mega_fwd_args* a, void* stream) {
  if (!a || !a->x || !a->params || !a->shadow || !a->z || !a->hmid || !a->u || !a->u2 || !a->qkv || !a->ctx || !a->a ||
      !a->m || !a->stats || !a->lse || !a->s_cls || !a->logits || !a->ws || (a->labels && !a->loss) ||
      (a->rows && (!a->rng || !a->rows_base)))
    return VITB200_ERR_ARG;
  const int T = a->Np + 1;
  if (a->B <= 0 || !vitb200_mega_supported(MG_H, MG_NH, T, a->P, a->C, a->layers, a->rope_cos != nullptr)) return VITB200_ERR_SHAPE;
  if (a->cluster != 1 && a->cluster != 2) return VITB200_ERR_ARG;
  const int L = a->layers, B = a->B;
  const bf16* sh = reinterpret_cast<const bf16*>(a->shadow);
  MegaFwdMaps tm;
  int rc;
  if ((rc = get_tmap(sh + a->off_wp, a->P, MG_H, 64, MG_H, &tm.wp))) return rc;
  const bf16* l0 = sh + a->off_layer0;
  if ((rc = get_tmap_weight(l0 + a->o_wqkv, MG_H, MG_Q, L, a->layer_stride, MG_Q, &tm.wq))) return rc;
  if ((rc = get_tmap_weight(l0 + a->o_wo, MG_H, MG_H, L, a->layer_stride, MG_H, &tm.wo))) return rc;
  if ((rc = get_tmap_weight(l0 + a->o_w1, MG_H, MG_I, L, a->layer_stride, MG_I, &tm.w1))) return rc;
  if ((rc = get_tmap_weight(l0 + a->o_w2, MG_I, MG_H, L, a->layer_stride, MG_H, &tm.w2))) return rc;
  if ((rc = get_tmap_act(a->z, 2 * MG_H, T, B, L + 1, &tm.z))) return rc;      // fp32 rows as 2H bf16 columns
  if ((rc = get_tmap_act(a->hmid, 2 * MG_H, T, B, L, &tm.hmid))) return rc;
  if ((rc = get_tmap_act(a->u, MG_H, T, B, L, &tm.u))) return rc;
  if ((rc = get_tmap_act(a->u2, MG_H, T, B, L, &tm.u2))) return rc;
  if ((rc = get_tmap_act(a->qkv, MG_Q, T, B, L, &tm.qkv))) return rc;
  if ((rc = get_tmap_act(a->ctx, MG_H, T, B, L, &tm.ctx))) return rc;
  if ((rc = get_tmap_act(a->a, MG_I, T, B, L, &tm.a))) return rc;
  if ((rc = get_tmap_act(a->m, MG_I, T, B, L, &tm.m))) return rc;
  const int smem = (int)vitb200_mega_fwd_smem_bytes(L);
  if (smem > 227 * 1024) return VITB200_ERR_SHAPE;
  static int max_set = 0;
  if (smem > max_set) {
    cudaError_t e = cudaFuncSetAttribute(mega_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return vb_cuda_error(e);
    max_set = smem;
  }
  vb_launch_pdl_cluster(mega_fwd_kernel, dim3(vitb200_mega_grid(B, a->cluster)), dim3(MF_THREADS), smem, (cudaStream_t)stream,
                        a->cluster, tm, *a);
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}
