// Fused forward row-chain kernels (bf16 mode, H in {32, 64}).  See include/vit_b200.h.
//
// One CTA = 128 token rows x 4 column groups = 512 threads (16 warps).  Warp w works on TMEM lane quarter (w & 3)
// -- rows 32 (w & 3) .. +31 of the tile, the only lanes a warp may read -- and on column group cg = w >> 2: every
// epilogue splits its columns four ways, so a row's work (bias, Philox dropout, GELU, bf16 packing, stores) is shared
// by four threads and each SM scheduler has four warps to interleave.  (The first version had one thread per row:
// 4 warps per SM, 4.7 cycles per issued instruction -- a pure dependent-latency chain.)  LayerNorm statistics of a row
// are combined across its four threads through shared memory (per-group mean / M2, merged with Chan's formula).
// Thread 0 issues TMA and tcgen05.mma; every GEMM's accumulator sits in TMEM columns [0, N); the epilogues store what
// backward needs to HBM and write the next GEMM's A operand into shared memory in the UMMA K-major / 128B-swizzle
// layout (16-byte chunk c of row r goes to chunk c ^ (r & 7)).
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_host.cuh"

namespace vb {
using namespace vb::tc;

constexpr int FF_ROWS = 128;
constexpr int FF_CG = 4;                       // column groups = threads per row
constexpr int FF_THREADS = FF_ROWS * FF_CG;    // 512
VB_TL_DECL(tl_layer_fwd)

__device__ __forceinline__ uint4 pack8_bf16(const float* v) {
  __nv_bfloat162 t0 = __floats2bfloat162_rn(v[0], v[1]), t1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 t2 = __floats2bfloat162_rn(v[4], v[5]), t3 = __floats2bfloat162_rn(v[6], v[7]);
  uint4 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
  pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
  return pk;
}
// A-operand tile: k-blocks of 64 columns, 16 KB each; row r = 128 bytes; chunk swizzle within the row
__device__ __forceinline__ void swz_store(uint8_t* tile, int r, int chunk, uint4 v) {
  uint8_t* p = tile + (chunk >> 3) * 16384 + r * 128 + (((chunk & 7) ^ (r & 7)) << 4);
  *reinterpret_cast<uint4*>(p) = v;
}
// values v[0..N) (already rounded to bf16 precision) -> global row (bf16) and the swizzled smem tile
template <int N>
__device__ __forceinline__ void emit_row_bf16(const float (&v)[N], bf16* gdst, bool store_g, uint8_t* tile, int r,
                                              int chunk0) {
#pragma unroll
  for (int j = 0; j < N; j += 8) {
    uint4 pk = pack8_bf16(&v[j]);
    if (store_g) *reinterpret_cast<uint4*>(gdst + j) = pk;
    if (tile) swz_store(tile, r, chunk0 + (j >> 3), pk);
  }
}
// LayerNorm statistics of a row whose H columns are spread over FF_CG threads (HC columns each).
// Each thread publishes (mean, M2) of its columns; after the barrier every thread merges the four pairs.
template <int HC>
__device__ __forceinline__ void row_stats(const float (&x)[HC], float2* s_ln, int r, int cg, float eps, float& mu,
                                          float& rs) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < HC; ++j) s += x[j];
  const float mc = s * (1.f / HC);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < HC; ++j) { const float d = x[j] - mc; q = fmaf(d, d, q); }
  s_ln[cg * FF_ROWS + r] = make_float2(mc, q);
  __syncthreads();
  float2 p[FF_CG];
#pragma unroll
  for (int g = 0; g < FF_CG; ++g) p[g] = s_ln[g * FF_ROWS + r];
  float m = 0.f;
#pragma unroll
  for (int g = 0; g < FF_CG; ++g) m += p[g].x;
  mu = m * (1.f / FF_CG);
  float m2 = 0.f;
#pragma unroll
  for (int g = 0; g < FF_CG; ++g) { const float d = p[g].x - mu; m2 += fmaf(d * d, (float)HC, p[g].y); }
  rs = rsqrtf(m2 * (1.f / (HC * FF_CG)) + eps);
}
// one thread issues a whole K-major x K-major GEMM: D[128, N] = A[128, K] * B[N, K]^T
__device__ __forceinline__ void issue_gemm_kk(uint32_t tmem_d, uint32_t sA, uint32_t sB, uint32_t b_kblock_bytes, int N,
                                              int K, uint64_t* done_bar) {
  const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
  const int ksteps = (K + 15) / 16;
  for (int k = 0; k < ksteps; ++k) {
    const int kb = k >> 2, kk = k & 3;
    const uint64_t da = make_sdesc_sw128(sA + kb * 16384 + kk * 32, 16, 1024);
    const uint64_t db = make_sdesc_sw128(sB + kb * b_kblock_bytes + kk * 32, 16, 1024);
    umma_bf16(tmem_d, da, db, idesc, k > 0 ? 1u : 0u);
  }
  umma_commit(done_bar);
}

// ================================================================================================
// layer kernel
// ================================================================================================
// 16-byte piece of an fp32 [128, H] tile staged in the TMA image (rows of 128 bytes per 32-float block, 128B swizzle):
// floats [c, c + 4) of row r, c a multiple of 4
__device__ __forceinline__ float4* f32_tile_ptr(uint8_t* tile, int r, int c) {
  return reinterpret_cast<float4*>(tile + (c >> 5) * 16384 + r * 128 + ((((c & 31) >> 2) ^ (r & 7)) << 4));
}
struct LayerFwdMaps {   // one kernel parameter: the 13 tensor maps of the layer kernel
  CUtensorMap ctx, wo, w1, w2, wq;          // loads
  CUtensorMap zin;                          // load (fp32 rows as 2H bf16 columns)
  CUtensorMap hmid, u2, a, m, zout, unext, qkv;   // stores
};

template <int H>
__global__ void __launch_bounds__(FF_THREADS, 1)
fused_layer_fwd_kernel(const __grid_constant__ LayerFwdMaps TM, const vitb200_layer_fwd_args P) {
  constexpr int I = 4 * H;
  constexpr int HC = H / FF_CG;                    // residual-stream columns per thread (8 / 16)
  constexpr int IC = I / FF_CG;                    // MLP columns per thread (32 / 64)
  constexpr int KB_I = I / 64;                     // k-blocks of the MLP-down GEMM
  constexpr int KB_H2 = 2 * H / 64;                // 16 KB blocks of an fp32 [128, H] tile
  constexpr int KB_Q = (3 * H + 63) / 64;          // 16 KB blocks of the qkv tile (staged over sM)
  constexpr bool STAGE_ACT = (H <= 32);            // H = 64: no room for a pre-activation image, `a` is stored directly
  constexpr uint32_t SZ_A = 16384, SZ_M = KB_I * 16384, SZ_WO = H * 128, SZ_W1 = I * 128, SZ_W2 = KB_I * H * 128,
                     SZ_WQ = 3 * H * 128, SZ_H = KB_H2 * 16384;
  static_assert(KB_Q <= KB_I, "qkv staging aliases the gelu tile");
  constexpr uint32_t TMEM_COLS = I < 32 ? 32 : I;  // 128 / 256
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = base;                 // ctx -> u2 -> u_next (A operand of GEMMs 1, 2, 4; also the u2 / u_next store image)
  uint8_t* sM = sA + SZ_A;            // gelu output (A operand of GEMM 3, store image of m); later the qkv store image
  uint8_t* sAct = sM + SZ_M;          // pre-activation store image (STAGE_ACT only)
  uint8_t* sH = sAct + (STAGE_ACT ? SZ_M : 0u);   // fp32 residual rows: z_in (TMA load) -> hmid (store) -> z_out (store)
  uint8_t* sWo = sH + SZ_H;
  uint8_t* sW1 = sWo + SZ_WO;
  uint8_t* sW2 = sW1 + SZ_W1;
  uint8_t* sWq = sW2 + SZ_W2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sWq + SZ_WQ);
  uint64_t *b_in = bars, *b_w1 = bars + 1, *b_w2 = bars + 2, *b_wq = bars + 3, *b_mma = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  // bias / LayerNorm vectors: staged once in shared memory (a dependent global load per use stalls every epilogue)
  float* prm = reinterpret_cast<float*>(bars + 8);
  float *s_bo = prm, *s_g2 = prm + H, *s_b2ln = prm + 2 * H, *s_b2 = prm + 3 * H, *s_gn = prm + 4 * H, *s_bn = prm + 5 * H,
        *s_b1 = prm + 6 * H, *s_bq = prm + 6 * H + I;
  float2* s_ln = reinterpret_cast<float2*>(prm + 13 * H);  // [FF_CG][128] row-statistics exchange

  const int tid = threadIdx.x, warp = tid >> 5;
  VB_TL(tl_layer_fwd, 0);
  const int r = ((warp & 3) << 5) | (tid & 31);    // row of the tile = TMEM lane
  const int cg = warp >> 2;                        // column group
  const int M = P.B * P.T;
  const int r0 = blockIdx.x * FF_ROWS;
  const int row = r0 + r;
  const bool valid = row < M;

  if (tid == 0) {
    tma_prefetch_desc(&TM.ctx); tma_prefetch_desc(&TM.wo); tma_prefetch_desc(&TM.w1); tma_prefetch_desc(&TM.w2);
    tma_prefetch_desc(&TM.zin);
    if (!P.last) tma_prefetch_desc(&TM.wq);
    mbar_init(b_in, 1); mbar_init(b_w1, 1); mbar_init(b_w2, 1); mbar_init(b_wq, 1); mbar_init(b_mma, 1);
    fence_barrier_init();
  }
  for (int j = threadIdx.x; j < H; j += FF_THREADS) {
    s_bo[j] = P.b_o[j]; s_g2[j] = P.ln2_g[j]; s_b2ln[j] = P.ln2_b[j]; s_b2[j] = P.b_2[j];
    s_gn[j] = P.lnn_g[j]; s_bn[j] = P.lnn_b[j];
  }
  for (int j = threadIdx.x; j < I; j += FF_THREADS) s_b1[j] = P.b_1[j];
  if (!P.last) for (int j = threadIdx.x; j < 3 * H; j += FF_THREADS) s_bq[j] = P.b_qkv[j];
  __syncthreads();
  if (tid == 0) {  // weights: not produced inside a step, so they are staged while the previous kernel still runs
    mbar_expect_tx(b_w1, SZ_WO + SZ_W1);
    tma_load_2d(sWo, &TM.wo, b_w1, 0, 0);
    tma_load_2d(sW1, &TM.w1, b_w1, 0, 0);
    mbar_expect_tx(b_w2, SZ_W2);
#pragma unroll
    for (int kb = 0; kb < KB_I; ++kb) tma_load_2d(sW2 + kb * H * 128, &TM.w2, b_w2, kb * 64, 0);
    if (!P.last) {
      mbar_expect_tx(b_wq, SZ_WQ);
      tma_load_2d(sWq, &TM.wq, b_wq, 0, 0);
    }
  }
  VB_TL(tl_layer_fwd, 1);
  pdl_wait();     // everything below reads what earlier kernels of this step produced
  pdl_trigger();  // the next kernel may start its prologue now (never earlier: see common.cuh)
  VB_TL(tl_layer_fwd, 2);
  if (tid == 0) {  // attention output (A of GEMM 1) and the fp32 residual rows
    mbar_expect_tx(b_in, SZ_A + SZ_H);
    tma_load_2d(sA, &TM.ctx, b_in, 0, r0);
#pragma unroll
    for (int kb = 0; kb < KB_H2; ++kb) tma_load_2d(sH + kb * 16384, &TM.zin, b_in, kb * 64, r0);
  }
  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t my_tmem = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  VB_TL(tl_layer_fwd, 3);
  const int hc0 = cg * HC;
  const uint64_t seed = P.rng ? P.rng[0] : 0ull;
  const uint32_t step = P.rng ? (uint32_t)P.rng[1] : 0u;
  const size_t erow = (size_t)(valid ? row : M - 1) * H + hc0;   // dropout element index of this thread's first column
  uint32_t mma_phase = 0;

  // ---- 1. attention output projection + dropout + residual, LayerNorm-after ----
  if (tid == 0) {
    mbar_wait(b_in, 0);
    mbar_wait(b_w1, 0);
    tc_fence_after();
    issue_gemm_kk(tmem, smem_u32(sA), smem_u32(sWo), 0, H, H, b_mma);
  }
  // this thread's HC columns of the residual row
  float h[HC];
  mbar_wait(b_in, 0);
#pragma unroll
  for (int j = 0; j < HC / 4; ++j) {
    const float4 t = *f32_tile_ptr(sH, r, hc0 + 4 * j);
    h[4 * j] = t.x; h[4 * j + 1] = t.y; h[4 * j + 2] = t.z; h[4 * j + 3] = t.w;
  }
  mbar_wait(b_mma, mma_phase); mma_phase ^= 1;
  VB_TL(tl_layer_fwd, 4);
  tc_fence_after();
  {
    const DropCtx dc = make_drop(P.p_drop, seed, step, P.site_proj);
    float v[HC];
    tmem_ld_cols<HC>(my_tmem + hc0, v);
#pragma unroll
    for (int j = 0; j < HC; j += 8) {
      float kp[8];
      drop8(dc, (erow + j) >> 3, kp);
#pragma unroll
      for (int e = 0; e < 8; ++e) h[j + e] += bf16_round(bf16_round(v[j + e] + s_bo[hc0 + j + e]) * kp[e]);
    }
    float mu, rs;
    row_stats<HC>(h, s_ln, r, cg, P.eps, mu, rs);
    float u2[HC];
#pragma unroll
    for (int j = 0; j < HC; ++j) u2[j] = bf16_round((h[j] - mu) * rs * s_g2[hc0 + j] + s_b2ln[hc0 + j]);
#pragma unroll
    for (int j = 0; j < HC / 4; ++j) *f32_tile_ptr(sH, r, hc0 + 4 * j) = make_float4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
    if (valid && cg == 0) { P.mean2[row] = mu; P.rstd2[row] = rs; }
    emit_row_bf16<HC>(u2, nullptr, false, sA, r, hc0 >> 3);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  VB_TL(tl_layer_fwd, 5);

  // ---- 2. MLP up + GELU ----
  if (tid == 0) {
    tc_fence_after();
    issue_gemm_kk(tmem, smem_u32(sA), smem_u32(sW1), 0, I, H, b_mma);
    // hmid and u2 leave through TMA (row-strided st.global from 512 threads costs a tag lookup per row)
#pragma unroll
    for (int kb = 0; kb < KB_H2; ++kb) tma_store_2d(&TM.hmid, sH + kb * 16384, kb * 64, r0);
    tma_store_2d(&TM.u2, sA, 0, r0);
    tma_store_commit();
  }
  mbar_wait(b_mma, mma_phase); mma_phase ^= 1;
  VB_TL(tl_layer_fwd, 6);
  tc_fence_after();
#pragma unroll 1
  for (int c0 = cg * IC; c0 < (cg + 1) * IC; c0 += 32) {
    float v[32];
    tmem_ld_32x32(my_tmem + c0, v);
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      // pre-activation rounded to bf16 once (packed), GELU evaluated on the rounded value, packed again
      uint32_t wa[4], wm[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 pa = __floats2bfloat162_rn(v[j + 2 * e] + s_b1[c0 + j + 2 * e], v[j + 2 * e + 1] + s_b1[c0 + j + 2 * e + 1]);
        const float2 f = __bfloat1622float2(pa);
        const __nv_bfloat162 pm = __floats2bfloat162_rn(gelu_f(f.x), gelu_f(f.y));
        wa[e] = *reinterpret_cast<const uint32_t*>(&pa);
        wm[e] = *reinterpret_cast<const uint32_t*>(&pm);
      }
      if (STAGE_ACT) swz_store(sAct, r, (c0 + j) >> 3, make_uint4(wa[0], wa[1], wa[2], wa[3]));
      else if (valid) *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(P.a) + (size_t)row * I + c0 + j) = make_uint4(wa[0], wa[1], wa[2], wa[3]);
      swz_store(sM, r, (c0 + j) >> 3, make_uint4(wm[0], wm[1], wm[2], wm[3]));
    }
  }
  fence_proxy_async();
  tc_fence_before();
  if (tid == 0) tma_store_wait_read<0>();  // hmid / u2 images are rewritten in stage 3
  __syncthreads();
  VB_TL(tl_layer_fwd, 7);

  // ---- 3. MLP down + dropout + residual, then LN1 of the next layer (or the final LN of CLS rows) ----
  if (tid == 0) {
    tc_fence_after();
    mbar_wait(b_w2, 0);
    issue_gemm_kk(tmem, smem_u32(sM), smem_u32(sW2), H * 128, H, I, b_mma);
#pragma unroll
    for (int kb = 0; kb < KB_I; ++kb) {
      if (STAGE_ACT) tma_store_2d(&TM.a, sAct + kb * 16384, kb * 64, r0);
      tma_store_2d(&TM.m, sM + kb * 16384, kb * 64, r0);
    }
    tma_store_commit();
  }
  mbar_wait(b_mma, mma_phase); mma_phase ^= 1;
  VB_TL(tl_layer_fwd, 8);
  tc_fence_after();
  {
    const DropCtx dc = make_drop(P.p_drop, seed, step, P.site_mlp);
    float v[HC];
    tmem_ld_cols<HC>(my_tmem + hc0, v);
#pragma unroll
    for (int j = 0; j < HC; j += 8) {
      float kp[8];
      drop8(dc, (erow + j) >> 3, kp);
#pragma unroll
      for (int e = 0; e < 8; ++e) h[j + e] += bf16_round(bf16_round(v[j + e] + s_b2[hc0 + j + e]) * kp[e]);
    }
#pragma unroll
    for (int j = 0; j < HC / 4; ++j) *f32_tile_ptr(sH, r, hc0 + 4 * j) = make_float4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
    float mu, rs;
    row_stats<HC>(h, s_ln, r, cg, P.eps, mu, rs);
    float un[HC];
#pragma unroll
    for (int j = 0; j < HC; ++j) un[j] = bf16_round((h[j] - mu) * rs * s_gn[hc0 + j] + s_bn[hc0 + j]);
    if (!P.last) {
      if (valid && cg == 0) { P.mean_n[row] = mu; P.rstd_n[row] = rs; }
      emit_row_bf16<HC>(un, nullptr, false, sA, r, hc0 >> 3);
    } else if (valid && (row % P.T) == 0) {
      const int b = row / P.T;
      if (cg == 0) { P.mean_n[b] = mu; P.rstd_n[b] = rs; }
      emit_row_bf16<HC>(un, reinterpret_cast<bf16*>(P.u_next) + (size_t)b * H + hc0, true, nullptr, 0, 0);
    }
  }
  VB_TL(tl_layer_fwd, 9);
  fence_proxy_async();
  tc_fence_before();
  if (tid == 0) tma_store_wait_read<0>();  // the gelu image is reused for qkv below
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int kb = 0; kb < KB_H2; ++kb) tma_store_2d(&TM.zout, sH + kb * 16384, kb * 64, r0);
    if (!P.last) tma_store_2d(&TM.unext, sA, 0, r0);
    tma_store_commit();
  }
  if (!P.last) {
    // ---- 4. fused Q/K/V projection of the next layer ----
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(b_wq, 0);
      issue_gemm_kk(tmem, smem_u32(sA), smem_u32(sWq), 0, 3 * H, H, b_mma);
    }
    mbar_wait(b_mma, mma_phase); mma_phase ^= 1;
  VB_TL(tl_layer_fwd, 10);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = cg * 32; c0 < 3 * H; c0 += 32 * FF_CG) {
      float v[32];
      tmem_ld_32x32(my_tmem + c0, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += s_bq[c0 + j];
      emit_row_bf16<32>(v, nullptr, false, sM, r, c0 >> 3);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int kb = 0; kb < KB_Q; ++kb) tma_store_2d(&TM.qkv, sM + kb * 16384, kb * 64, r0);
      tma_store_commit();
    }
  } else {
    tc_fence_before();
    __syncthreads();
  }
  VB_TL(tl_layer_fwd, 11);
  if (tid == 0) tma_store_wait_all();   // outputs are globally visible before the CTA (and, last, the grid) retires
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

// ================================================================================================
// embedding kernel
// ================================================================================================
struct EmbedFwdMaps { CUtensorMap wp, wq, z0, u, qkv; };

template <int H>
__global__ void __launch_bounds__(FF_THREADS, 1)
fused_embed_fwd_kernel(const __grid_constant__ EmbedFwdMaps TM, const vitb200_embed_fwd_args P) {
  constexpr int HC = H / FF_CG;
  constexpr int KB_H2 = 2 * H / 64, KB_Q = (3 * H + 63) / 64;
  constexpr uint32_t SZ_A = 16384, SZ_WP = H * 128, SZ_WQ = 3 * H * 128, SZ_H = KB_H2 * 16384, SZ_Q = KB_Q * 16384;
  constexpr uint32_t TMEM_COLS = 3 * H <= 128 ? 128 : 256;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = base;                 // patch windows (A of GEMM 1) -> u (A of GEMM 2, store image)
  uint8_t* sH = sA + SZ_A;            // z0 store image (fp32 rows)
  uint8_t* sQ = sH + SZ_H;            // qkv store image
  uint8_t* sWp = sQ + SZ_Q;
  uint8_t* sWq = sWp + SZ_WP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sWq + SZ_WQ);
  uint64_t *b_wp = bars, *b_wq = bars + 1, *b_mma = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  float* prm = reinterpret_cast<float*>(bars + 8);
  float *s_bp = prm, *s_cls = prm + H, *s_g = prm + 2 * H, *s_b = prm + 3 * H, *s_bq = prm + 4 * H;
  float2* s_ln = reinterpret_cast<float2*>(prm + 7 * H);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int r = ((warp & 3) << 5) | (tid & 31);
  const int cg = warp >> 2;
  const int T = P.Np + 1, M = P.B * T;
  const int r0 = blockIdx.x * FF_ROWS;
  const int row = r0 + r;
  const bool valid = row < M;
  const int rowc = valid ? row : M - 1;
  const int t = rowc % T;

  if (tid == 0) {
    tma_prefetch_desc(&TM.wp); tma_prefetch_desc(&TM.wq);
    mbar_init(b_wp, 1); mbar_init(b_wq, 1); mbar_init(b_mma, 1);
    fence_barrier_init();
  }
  __syncthreads();
  // The embedding kernel is the FIRST kernel of a step: its predecessor is the previous step's optimizer kernel,
  // which writes the weights.  So here even the weight staging has to come after pdl_wait().
  pdl_wait();
  pdl_trigger();
  if (tid == 0) {
    mbar_expect_tx(b_wp, SZ_WP);
    tma_load_2d(sWp, &TM.wp, b_wp, 0, 0);
    mbar_expect_tx(b_wq, SZ_WQ);
    tma_load_2d(sWq, &TM.wq, b_wq, 0, 0);
  }
  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
  for (int j = threadIdx.x; j < H; j += FF_THREADS) { s_bp[j] = P.b_p[j]; s_cls[j] = P.cls[j]; s_g[j] = P.ln_g[j]; s_b[j] = P.ln_b[j]; }
  for (int j = threadIdx.x; j < 3 * H; j += FF_THREADS) s_bq[j] = P.b_qkv[j];
  // A rows = the patch windows of the tokens (tokenization.py:45-48); CLS rows and padded windows are zero.
  // Work items (row, 8-float chunk) are dealt so that consecutive lanes read consecutive chunks of one window:
  // a warp-wide load touches 32 / nchunk windows instead of 32.
  {
    const int nchunk = (P.P + 15) / 16 * 2;  // whole 16-element k-steps are read by the MMA
    const bool vec = (P.S % 4 == 0) && (P.L % 4 == 0) && (P.P % 8 == 0);
    for (int it = tid; it < FF_ROWS * nchunk; it += FF_THREADS) {
      const int rr = it / nchunk, c = it - rr * nchunk;
      const int grow = r0 + rr;
      const int gb = grow / T, gt = grow - gb * T;
      const bool has = grow < M && gt >= 1 && (gt - 1) < P.n_valid;
      const float* xp = P.x + (size_t)gb * P.L + (size_t)(gt >= 1 ? gt - 1 : 0) * P.S;
      float v[8];
      if (vec && has && c * 8 + 8 <= P.P) {
        const float4 a0 = *reinterpret_cast<const float4*>(xp + c * 8), a1 = *reinterpret_cast<const float4*>(xp + c * 8 + 4);
        v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int j = c * 8 + q;
          v[q] = (has && j < P.P) ? xp[j] : 0.f;
        }
      }
      swz_store(sA, rr, c, pack8_bf16(v));
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t my_tmem = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t mma_phase = 0;
  if (tid == 0) {
    mbar_wait(b_wp, 0);
    issue_gemm_kk(tmem, smem_u32(sA), smem_u32(sWp), 0, H, P.P, b_mma);
  }
  mbar_wait(b_mma, mma_phase); mma_phase ^= 1;
  tc_fence_after();
  {
    const uint64_t seed = P.rng ? P.rng[0] : 0ull;
    const uint32_t step = P.rng ? (uint32_t)P.rng[1] : 0u;
    const DropCtx dc = make_drop(P.p_drop, seed, step, VITB200_SITE_EMB);
    const int hc0 = cg * HC;
    float z[HC], v[HC];
    tmem_ld_cols<HC>(my_tmem + hc0, v);
#pragma unroll
    for (int j = 0; j < HC; j += 8) {
      float kp[8];
      drop8(dc, ((size_t)rowc * H + hc0 + j) >> 3, kp);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int c = hc0 + j + q;
        float e = t == 0 ? s_cls[c] : bf16_round(v[j + q] + s_bp[c]);
        if (P.pos) e += P.pos[(size_t)t * H + c];
        z[j + q] = e * kp[q];
      }
    }
    float mu, rs;
    row_stats<HC>(z, s_ln, r, cg, P.eps, mu, rs);
    float u[HC];
#pragma unroll
    for (int j = 0; j < HC; ++j) u[j] = bf16_round((z[j] - mu) * rs * s_g[hc0 + j] + s_b[hc0 + j]);
#pragma unroll
    for (int j = 0; j < HC / 4; ++j) *f32_tile_ptr(sH, r, hc0 + 4 * j) = make_float4(z[4 * j], z[4 * j + 1], z[4 * j + 2], z[4 * j + 3]);
    if (valid && cg == 0) { P.mean[row] = mu; P.rstd[row] = rs; }
    emit_row_bf16<HC>(u, nullptr, false, sA, r, hc0 >> 3);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    mbar_wait(b_wq, 0);
    issue_gemm_kk(tmem, smem_u32(sA), smem_u32(sWq), 0, 3 * H, H, b_mma);
#pragma unroll
    for (int kb = 0; kb < KB_H2; ++kb) tma_store_2d(&TM.z0, sH + kb * 16384, kb * 64, r0);
    tma_store_2d(&TM.u, sA, 0, r0);
    tma_store_commit();
  }
  mbar_wait(b_mma, mma_phase); mma_phase ^= 1;
  tc_fence_after();
#pragma unroll 1
  for (int c0 = cg * 32; c0 < 3 * H; c0 += 32 * FF_CG) {
    float v[32];
    tmem_ld_32x32(my_tmem + c0, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += s_bq[c0 + j];
    emit_row_bf16<32>(v, nullptr, false, sQ, r, c0 >> 3);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int kb = 0; kb < KB_Q; ++kb) tma_store_2d(&TM.qkv, sQ + kb * 16384, kb * 64, r0);
    tma_store_commit();
    tma_store_wait_all();
  }
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

template <int H> static constexpr int layer_smem() {
  return 16384 + (H <= 32 ? 2 : 1) * (4 * H / 64) * 16384 + (2 * H / 64) * 16384 + H * 128 + 4 * H * 128 + (4 * H / 64) * H * 128 +
         3 * H * 128 + 1024 + 64 + (13 * H) * 4 + 4096 + 64;
}
template <int H> static constexpr int embed_smem() {
  return 16384 + (2 * H / 64) * 16384 + ((3 * H + 63) / 64) * 16384 + H * 128 + 3 * H * 128 + 1024 + 64 + (7 * H) * 4 + 4096 + 64;
}

template <int H>
static int launch_layer(const vitb200_layer_fwd_args* a, cudaStream_t st) {
  const int M = a->B * a->T, I = 4 * H;
  LayerFwdMaps tm;
  int rc;
  if ((rc = get_tmap(a->ctx, H, M, 64, 128, &tm.ctx))) return rc;
  if ((rc = get_tmap(a->w_o, H, H, 64, H, &tm.wo))) return rc;
  if ((rc = get_tmap(a->w_1, H, I, 64, I, &tm.w1))) return rc;
  if ((rc = get_tmap(a->w_2, I, H, 64, H, &tm.w2))) return rc;
  if (a->last) tm.wq = tm.wo;
  else if ((rc = get_tmap(a->w_qkv, H, 3 * H, 64, 3 * H, &tm.wq))) return rc;
  // fp32 [M, H] rows are moved as 2H bf16 columns
  if ((rc = get_tmap(a->z_in, 2 * H, M, 64, 128, &tm.zin))) return rc;
  if ((rc = get_tmap(a->hmid, 2 * H, M, 64, 128, &tm.hmid))) return rc;
  if ((rc = get_tmap(a->z_out, 2 * H, M, 64, 128, &tm.zout))) return rc;
  if ((rc = get_tmap(a->u2, H, M, 64, 128, &tm.u2))) return rc;
  if ((rc = get_tmap(a->a, I, M, 64, 128, &tm.a))) return rc;
  if ((rc = get_tmap(a->m, I, M, 64, 128, &tm.m))) return rc;
  if (a->last) { tm.unext = tm.u2; tm.qkv = tm.u2; }
  else {
    if ((rc = get_tmap(a->u_next, H, M, 64, 128, &tm.unext))) return rc;
    if ((rc = get_tmap(a->qkv_next, 3 * H, M, 64, 128, &tm.qkv))) return rc;
  }
  auto kern = fused_layer_fwd_kernel<H>;
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, layer_smem<H>());
    if (e != cudaSuccess) return vb_cuda_error(e);
    done = true;
  }
  vb_launch_pdl(kern, dim3((M + FF_ROWS - 1) / FF_ROWS), dim3(FF_THREADS), layer_smem<H>(), st, tm, *a);
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

template <int H>
static int launch_embed(const vitb200_embed_fwd_args* a, cudaStream_t st) {
  const int M = a->B * (a->Np + 1);
  EmbedFwdMaps tm;
  int rc;
  if ((rc = get_tmap(a->w_p, a->P, H, 64, H, &tm.wp))) return rc;
  if ((rc = get_tmap(a->w_qkv, H, 3 * H, 64, 3 * H, &tm.wq))) return rc;
  if ((rc = get_tmap(a->z0, 2 * H, M, 64, 128, &tm.z0))) return rc;   // fp32 rows as 2H bf16 columns
  if ((rc = get_tmap(a->u, H, M, 64, 128, &tm.u))) return rc;
  if ((rc = get_tmap(a->qkv, 3 * H, M, 64, 128, &tm.qkv))) return rc;
  auto kern = fused_embed_fwd_kernel<H>;
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, embed_smem<H>());
    if (e != cudaSuccess) return vb_cuda_error(e);
    done = true;
  }
  vb_launch_pdl(kern, dim3((M + FF_ROWS - 1) / FF_ROWS), dim3(FF_THREADS), embed_smem<H>(), st, tm, *a);
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

}  // namespace vb

using namespace vb;

VB_TL_EXPORT(vitb200_tl_layer_fwd, vb::tl_layer_fwd)

extern "C" int vitb200_fused_supported(int H, int P) {
  return ((H == 32 || H == 64) && P % 8 == 0 && P >= 8 && P <= 64) ? 1 : 0;
}

extern "C" int vitb200_fused_layer_fwd(const vitb200_layer_fwd_args* a, void* stream) {
  if (!a || !a->ctx || !a->z_in || !a->w_o || !a->w_1 || !a->w_2 || !a->hmid || !a->u2 || !a->a || !a->m || !a->z_out ||
      !a->u_next || !a->mean2 || !a->rstd2 || !a->mean_n || !a->rstd_n)
    return VITB200_ERR_ARG;
  if (!a->last && (!a->w_qkv || !a->qkv_next || !a->b_qkv)) return VITB200_ERR_ARG;
  if (a->B <= 0 || a->T <= 0) return VITB200_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (a->H == 32) return launch_layer<32>(a, st);
  if (a->H == 64) return launch_layer<64>(a, st);
  return VITB200_ERR_SHAPE;
}

extern "C" int vitb200_fused_embed_fwd(const vitb200_embed_fwd_args* a, void* stream) {
  if (!a || !a->x || !a->w_p || !a->b_p || !a->cls || !a->ln_g || !a->ln_b || !a->w_qkv || !a->b_qkv || !a->z0 || !a->u ||
      !a->mean || !a->rstd || !a->qkv)
    return VITB200_ERR_ARG;
  if (a->B <= 0 || !vitb200_fused_supported(a->H, a->P)) return VITB200_ERR_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  if (a->H == 32) return launch_embed<32>(a, st);
  return launch_embed<64>(a, st);
}
