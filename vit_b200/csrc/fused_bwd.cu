// Fused backward row-chain kernels (bf16 mode, H = 32).  See include/vit_b200.h.
//
// One CTA = 128 rows x 4 column groups = 512 threads (16 warps): warp w reads TMEM lane quarter (w & 3) = rows
// 32 (w & 3) .. +31 of the current tile and handles column group cg = w >> 2 of every epilogue (see fused_fwd.cu for
// why: one thread per row left each SM scheduler with a single, latency-bound warp).
// Every activation-gradient tile (ddelta2, da, ddelta1, dqkv) is written/loaded ONCE into shared memory in
// the 128B-swizzled row image and then consumed twice by tcgen05.mma:
//   dgrad  dX = dY . W      : the tile is operand A, K-major view   (contraction over its columns)
//   wgrad  dW = dY^T . X    : the tile is operand A, MN-major view  (contraction over its 128 rows)
// wgrad accumulators stay in TMEM across the tiles of a persistent CTA and are written out once.
// All parameter-gradient reductions run in a fixed order (deterministic=True, basemodule.py:250).
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_host.cuh"

namespace vb {
using namespace vb::tc;

constexpr int FB_CG = 4;
constexpr int FB_THREADS = 128 * FB_CG;
constexpr int FB_H = 32;
constexpr int FB_I = 128;
constexpr int FB_HC = FB_H / FB_CG;   // 8 residual-stream columns per thread
VB_TL_DECL(tl_bwd_upper)

__device__ __forceinline__ uint4 fb_pack8(const float* v) {
  __nv_bfloat162 t0 = __floats2bfloat162_rn(v[0], v[1]), t1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 t2 = __floats2bfloat162_rn(v[4], v[5]), t3 = __floats2bfloat162_rn(v[6], v[7]);
  uint4 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
  pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
  return pk;
}
__device__ __forceinline__ void fb_swz_store(uint8_t* tile, int r, int chunk, uint4 v) {
  *reinterpret_cast<uint4*>(tile + (chunk >> 3) * 16384 + r * 128 + (((chunk & 7) ^ (r & 7)) << 4)) = v;
}
// sum over rows [rb, rb + NR) of column `col` of a swizzled bf16 tile (rows in order => deterministic)
template <int NR>
__device__ __forceinline__ float fb_colsum(const uint8_t* tile, int col, int rb) {
  const uint8_t* blk = tile + (col >> 6) * 16384;
  const int chunk = (col & 63) >> 3, within = (col & 7) * 2;
  float s = 0.f;
#pragma unroll 8
  for (int r = rb; r < rb + NR; ++r) {
    const bf16 v = *reinterpret_cast<const bf16*>(blk + r * 128 + ((chunk ^ (r & 7)) << 4) + within);
    s += __bfloat162float(v);
  }
  return s;
}

// a 128-byte row of a TMA-loaded (128B-swizzled) tile: 16-byte chunk c of row r
__device__ __forceinline__ const uint8_t* fb_swz_ptr(const uint8_t* tile, int r, int chunk) {
  return tile + (chunk >> 3) * 16384 + r * 128 + (((chunk & 7) ^ (r & 7)) << 4);
}
// FB_HC = 8 consecutive fp32 values (columns 8 cg .. 8 cg + 7) of row r of a TMA-staged fp32 [128, 32] tile
__device__ __forceinline__ void fb_ld_f8(const uint8_t* tile, int r, int cg, float (&x)[FB_HC]) {
  const float4 a = *reinterpret_cast<const float4*>(fb_swz_ptr(tile, r, 2 * cg));
  const float4 b = *reinterpret_cast<const float4*>(fb_swz_ptr(tile, r, 2 * cg + 1));
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}

// operand view for one tcgen05 GEMM
struct Opnd {
  uint32_t addr;   // shared address of the tile
  uint32_t lbo;    // MN-major: byte stride between 64-wide MN groups; K-major: ignored (16)
  uint32_t kblk;   // K-major: byte stride between 64-wide k-blocks
  int mn;          // 1 = MN-major view (k = tile rows), 0 = K-major view (k = tile columns)
};
__device__ __forceinline__ uint64_t opnd_desc(const Opnd& o, int k) {
  if (o.mn) return make_sdesc_sw128(o.addr + k * 2048, o.lbo, 1024);
  return make_sdesc_sw128(o.addr + (k >> 2) * o.kblk + (k & 3) * 32, 16, 1024);
}
// D[128, N] (+)= A * B over `ksteps` k-steps of 16
__device__ __forceinline__ void fb_issue(uint32_t tmem_d, const Opnd& A, const Opnd& B, int N, int ksteps, bool accumulate) {
  const uint32_t idesc = make_idesc_bf16(128, N, A.mn, B.mn);
  for (int k = 0; k < ksteps; ++k) umma_bf16(tmem_d, opnd_desc(A, k), opnd_desc(B, k), idesc, (accumulate || k > 0) ? 1u : 0u);
}

// LayerNorm backward of one row spread over FB_CG threads: the two row sums are exchanged through shared memory.
// du: upstream gradient (bf16-rounded), xh: normalised input; on return g = du * gamma and (c1, c2) the row means.
__device__ __forceinline__ void fb_ln_bwd_sums(const float (&du)[FB_HC], const float (&xh)[FB_HC], const float* s_gamma, int hc0,
                                               float2* s_ex, int r, int cg, float (&g)[FB_HC], float& c1, float& c2) {
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < FB_HC; ++j) {
    g[j] = du[j] * s_gamma[hc0 + j];
    s1 += g[j];
    s2 = fmaf(g[j], xh[j], s2);
  }
  s_ex[cg * 128 + r] = make_float2(s1, s2);
  __syncthreads();
  float t1 = 0.f, t2 = 0.f;
#pragma unroll
  for (int q = 0; q < FB_CG; ++q) { const float2 p = s_ex[q * 128 + r]; t1 += p.x; t2 += p.y; }
  c1 = t1 * (1.f / FB_H);
  c2 = t2 * (1.f / FB_H);
}

// red[e * 128 + row] (e < NE) -> out(e) = sum over the 128 rows, in a fixed order: 8 threads x 16 rows, then a shuffle tree.
// Call with all FB_THREADS threads after a __syncthreads(); NE * 8 <= FB_THREADS.
template <typename F>
__device__ __forceinline__ void fb_reduce_rows(const float* red, int NE, F&& out) {
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

// ================================================================================================
// upper: dz -> [W2 dgrad/wgrad] -> gelu' -> [W1 dgrad/wgrad] -> LN2 bwd -> dh ; [Wo dgrad/wgrad] -> dctx
// ================================================================================================
__global__ void __launch_bounds__(FB_THREADS, 1)
fused_bwd_upper_kernel(const __grid_constant__ CUtensorMap tmM, const __grid_constant__ CUtensorMap tmU2,
                       const __grid_constant__ CUtensorMap tmCtx, const __grid_constant__ CUtensorMap tmW2,
                       const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmWo,
                       const __grid_constant__ CUtensorMap tmAct, const __grid_constant__ CUtensorMap tmHm,
                       const __grid_constant__ CUtensorMap tmDz, const __grid_constant__ CUtensorMap tmDh,
                       const __grid_constant__ CUtensorMap tmDctx, const vitb200_layer_bwd_upper_args P) {
  constexpr int H = FB_H, I = FB_I, HC = FB_HC;
  // shared memory map (all tile bases 1024-aligned); sD must be directly followed by sDA (see wgrad A views)
  constexpr uint32_t O_D = 0, O_DA = 16384, O_M = O_DA + 32768, O_U2 = O_M + 32768, O_CTX = O_U2 + 16384,
                     O_W2 = O_CTX + 16384, O_W1 = O_W2 + 8192, O_WO = O_W1 + 16384, O_ACT = O_WO + 4096,
                     O_HM = O_ACT + 32768, O_DZ = O_HM + 16384, O_BAR = O_DZ + 16384, O_EX = O_BAR + 1024;
  // TMEM columns
  // (dW1 is 48 columns wide: column H is the bias gradient db1, produced by a column of ones in the u2 tile)
  constexpr uint32_t C_W2 = 0, C_WO = 160, C_DM = 192, C_DU2 = 320, C_DCTX = 352, C_W1 = 384, TMEM_COLS = 512;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t *sD = base + O_D, *sDA = base + O_DA, *sM = base + O_M, *sU2 = base + O_U2, *sCtx = base + O_CTX;
  uint8_t *sW2 = base + O_W2, *sW1 = base + O_W1, *sWo = base + O_WO;
  const uint8_t *sAct = base + O_ACT, *sHm = base + O_HM;  // pre-GELU activations (bf16) and hmid rows (fp32), TMA-staged
  uint8_t* sDz = base + O_DZ;   // fp32 rows: dz (TMA load) -> dh (TMA store image)
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + O_BAR);
  uint64_t *b_w = bars, *b_tile = bars + 1, *b_mma = bars + 2, *b_dz = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  float* s_g2 = reinterpret_cast<float*>(bars + 8);
  float2* s_ex = reinterpret_cast<float2*>(base + O_EX);   // [FB_CG][128] row-sum exchange (4 KB)

  const int tid = threadIdx.x, warp = tid >> 5;
  VB_TL(tl_bwd_upper, 0);
  const int r = ((warp & 3) << 5) | (tid & 31);  // row of the tile = TMEM lane
  const int cg = warp >> 2, hc0 = cg * HC;
  const int M = P.B * P.T;
  const int ntiles = (M + 127) / 128;

  if (tid < H) s_g2[tid] = P.ln2_g[tid];
  if (tid == 0) {
    tma_prefetch_desc(&tmM); tma_prefetch_desc(&tmU2); tma_prefetch_desc(&tmCtx);
    tma_prefetch_desc(&tmW2); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmWo);
    tma_prefetch_desc(&tmAct); tma_prefetch_desc(&tmHm); tma_prefetch_desc(&tmDz);
    mbar_init(b_w, 1); mbar_init(b_tile, 1); mbar_init(b_mma, 1); mbar_init(b_dz, 1);
    fence_barrier_init();
  }
  // columns 32..63 of the sD tile are never written by the epilogues: zero them once (MN-major wgrad view reads them)
  fb_swz_store(sD, r, 4 + cg, make_uint4(0u, 0u, 0u, 0u));
  __syncthreads();
  if (tid == 0) {  // weights are not written inside a step: staged before the dependency wait
    mbar_expect_tx(b_w, 8192 + 16384 + 4096);
    tma_load_2d(sW2, &tmW2, b_w, 0, 0);          // W2 [H rows, I cols]: two {64 cols, H rows} boxes
    tma_load_2d(sW2 + 4096, &tmW2, b_w, 64, 0);
    tma_load_2d(sW1, &tmW1, b_w, 0, 0);          // W1 [I rows, H cols]: one {64 cols (32 valid), 128 rows} box
    tma_load_2d(sWo, &tmWo, b_w, 0, 0);          // Wo [H rows, H cols]
  }
  VB_TL(tl_bwd_upper, 1);
  pdl_wait();
  pdl_trigger();
  VB_TL(tl_bwd_upper, 2);
  const bool from_cls = P.dz_cls != nullptr;
  auto load_tile = [&](int r0) {   // thread 0: all TMA loads of one 128-row tile
    if (!from_cls) {   // dz gates the first stage: it gets its own barrier and goes first
      mbar_expect_tx(b_dz, 16384);
      tma_load_2d(sDz, &tmDz, b_dz, 0, r0);
    }
    mbar_expect_tx(b_tile, 32768 + 16384 + 16384 + 32768 + 16384);
    tma_load_2d(sM, &tmM, b_tile, 0, r0);
    tma_load_2d(sM + 16384, &tmM, b_tile, 64, r0);
    tma_load_2d(sU2, &tmU2, b_tile, 0, r0);
    tma_load_2d(sCtx, &tmCtx, b_tile, 0, r0);
    tma_load_2d(base + O_ACT, &tmAct, b_tile, 0, r0);
    tma_load_2d(base + O_ACT + 16384, &tmAct, b_tile, 64, r0);
    tma_load_2d(base + O_HM, &tmHm, b_tile, 0, r0);
  };
  if (tid == 0 && (int)blockIdx.x < ntiles) load_tile(blockIdx.x * 128);   // first tile: in flight while TMEM is allocated
  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t my_tmem = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  VB_TL(tl_bwd_upper, 3);
  const uint32_t aD = smem_u32(sD), aDA = smem_u32(sDA), aM = smem_u32(sM), aU2 = smem_u32(sU2), aCtx = smem_u32(sCtx);
  const uint32_t aW2 = smem_u32(sW2), aW1 = smem_u32(sW1), aWo = smem_u32(sWo);
  // operand views
  const Opnd D_k{aD, 16, 16384, 0};            // ddelta tile, K-major (K = H)
  const Opnd D_mn{aD, 16384, 0, 1};            // ddelta tile, MN-major; MN group 1 aliases sDA: rows 64..127 of the
                                               // product are never read (only H = 32 weight rows exist)
  const Opnd DA_k{aDA, 16, 16384, 0};          // da tile, K-major (K = I, two k-blocks)
  const Opnd DA_mn{aDA, 16384, 0, 1};          // da tile, MN-major (MN = I = 2 groups)
  const Opnd M_mn{aM, 16384, 0, 1}, U2_mn{aU2, 16384, 0, 1}, CTX_mn{aCtx, 16384, 0, 1};
  const Opnd W2_mn{aW2, 4096, 0, 1};           // B(n = i, k = h): N = I (2 groups of 64), K = H rows
  const Opnd W1_mn{aW1, 16384, 0, 1};          // B(n = h, k = i): N = H, K = I rows
  const Opnd WO_mn{aWo, 4096, 0, 1};           // B(n = k_out, k = n_in)

  const uint64_t seed = P.rng ? P.rng[0] : 0ull;
  const uint32_t step = P.rng ? (uint32_t)P.rng[1] : 0u;
  const DropCtx dc_mlp = make_drop(P.p_drop, seed, step, P.site_mlp);
  const DropCtx dc_proj = make_drop(P.p_drop, seed, step, P.site_proj);

  float acc_g[HC], acc_b[HC];  // LN2 gamma / beta gradients of this thread's row, columns hc0 .. hc0 + 7
#pragma unroll
  for (int j = 0; j < HC; ++j) { acc_g[j] = 0.f; acc_b[j] = 0.f; }
  // bias gradients = column sums of the gradient tiles; partial per thread:
  //   H-wide tiles: column tid & 31, rows 8 (tid >> 5) .. +7 ;  I-wide tile: column tid & 127, rows 32 (tid >> 7) .. +31
  float acc_b2 = 0.f, acc_bo = 0.f;
  uint32_t ph_mma = 0, ph_tile = 0;
  int iter = 0;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++iter) {
    const int r0 = tile * 128, row = r0 + r;
    const bool valid = row < M;
    const int rowc = valid ? row : M - 1;
    if (tid == 0 && iter > 0) {
      tma_store_wait_read<0>();   // the previous tile's dh / dctx images are about to be overwritten
      load_tile(r0);
    }
    const float mu = P.mean2[rowc], rs = P.rstd2[rowc];  // issued early, consumed by the LayerNorm stage
    // ---- ddelta2 = dropout'(dz) (bf16) -> sD ----
    float dz[HC];
    {
      if (from_cls) {
        // top layer: only the CLS rows carry a gradient (the head reads last_hidden_state[:, 0], specvit.py:78)
        const bool nz = (rowc % P.T) == 0;
        const float4* p = reinterpret_cast<const float4*>(P.dz_cls + (size_t)(rowc / P.T) * H + hc0);
#pragma unroll
        for (int j = 0; j < HC / 4; ++j) {
          float4 t = nz ? p[j] : make_float4(0.f, 0.f, 0.f, 0.f);
          dz[4 * j] = t.x; dz[4 * j + 1] = t.y; dz[4 * j + 2] = t.z; dz[4 * j + 3] = t.w;
        }
      }
      float kp[8];
      drop8(dc_mlp, ((size_t)rowc * H + hc0) >> 3, kp);   // overlaps the tile loads
      if (!from_cls) {
        mbar_wait(b_dz, ph_tile);
        fb_ld_f8(sDz, r, cg, dz);
      }
      float d2[HC];
#pragma unroll
      for (int j = 0; j < HC; ++j) d2[j] = valid ? bf16_round(dz[j]) * kp[j] : 0.f;
      fb_swz_store(sD, r, cg, fb_pack8(d2));
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
  VB_TL(tl_bwd_upper, 4);
    if (tid == 0) {
      tc_fence_after();
      if (iter == 0) mbar_wait(b_w, 0);
      fb_issue(tmem + C_DM, D_k, W2_mn, I, H / 16, false);        // dm[row, i]  = sum_h ddelta2[row,h] W2[h,i]
      umma_commit(b_mma);                                         // the gelu' stage only needs dm
      mbar_wait(b_tile, ph_tile);
      tc_fence_after();
      fb_issue(tmem + C_W2, D_mn, M_mn, I, 8, iter > 0);          // dW2[h, i]  += sum_rows ddelta2[row,h] m[row,i]
    }                                                             // (completion is covered by the next commit)
    acc_b2 += fb_colsum<8>(sD, tid & 31, (tid >> 5) * 8);  // overlaps the MMAs
    mbar_wait(b_tile, ph_tile);  // every thread reads the TMA-staged a / hmid rows below
    ph_tile ^= 1;
    // a column of ones next to u2 (columns H..63 of the tile are TMA zero fill): the dW1 MMA then also emits db1
    if (cg == 0) *reinterpret_cast<uint32_t*>(const_cast<uint8_t*>(fb_swz_ptr(sU2, r, H / 8))) = 0x00003F80u;   // bf16 {1, 0}
    mbar_wait(b_mma, ph_mma); ph_mma ^= 1;
  VB_TL(tl_bwd_upper, 5);
    tc_fence_after();
    // ---- da = dm * gelu'(a) -> sDA ----
    {
      const int c0 = cg * 32;
      float v[32];
      tmem_ld_32x32(my_tmem + C_DM + c0, v);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 av = *reinterpret_cast<const uint4*>(fb_swz_ptr(sAct, r, (c0 >> 3) + q));
        const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&av);
        float o[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(a2[e]);
          const float2 dm = __bfloat1622float2(__floats2bfloat162_rn(v[q * 8 + 2 * e], v[q * 8 + 2 * e + 1]));
          o[2 * e] = valid ? dm.x * gelu_grad_f(f.x) : 0.f;
          o[2 * e + 1] = valid ? dm.y * gelu_grad_f(f.y) : 0.f;
        }
        fb_swz_store(sDA, r, (c0 >> 3) + q, fb_pack8(o));
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
  VB_TL(tl_bwd_upper, 6);
      fb_issue(tmem + C_DU2, DA_k, W1_mn, H, I / 16, false);      // du2[row, h] = sum_i da[row,i] W1[i,h]
      fb_issue(tmem + C_W1, DA_mn, U2_mn, H + 16, 8, iter > 0);   // dW1[i, h]  += sum_rows da[row,i] u2[row,h] ; column H: db1[i]
      umma_commit(b_mma);
    }
    mbar_wait(b_mma, ph_mma); ph_mma ^= 1;
  VB_TL(tl_bwd_upper, 7);
    tc_fence_after();
    // ---- LayerNorm-after backward + residual -> dh ; ddelta1 = dropout'(dh) -> sD ----
    {
      float du[HC], xh[HC], g[HC];
      tmem_ld_32x8(my_tmem + C_DU2 + hc0, du);
      fb_ld_f8(sHm, r, cg, xh);
#pragma unroll
      for (int j = 0; j < HC; ++j) {
        xh[j] = (xh[j] - mu) * rs;
        du[j] = valid ? bf16_round(du[j]) : 0.f;
        acc_g[j] = fmaf(du[j], xh[j], acc_g[j]);
        acc_b[j] += du[j];
      }
      float c1, c2;
      fb_ln_bwd_sums(du, xh, s_g2, hc0, s_ex, r, cg, g, c1, c2);
#pragma unroll
      for (int j = 0; j < HC; ++j) dz[j] += rs * (g[j] - c1 - xh[j] * c2);  // dz now holds dh
      *reinterpret_cast<float4*>(const_cast<uint8_t*>(fb_swz_ptr(sDz, r, 2 * cg))) = make_float4(dz[0], dz[1], dz[2], dz[3]);
      *reinterpret_cast<float4*>(const_cast<uint8_t*>(fb_swz_ptr(sDz, r, 2 * cg + 1))) = make_float4(dz[4], dz[5], dz[6], dz[7]);
      float kp[8], d1[HC];
      drop8(dc_proj, ((size_t)rowc * H + hc0) >> 3, kp);
#pragma unroll
      for (int j = 0; j < HC; ++j) d1[j] = valid ? bf16_round(dz[j]) * kp[j] : 0.f;
      fb_swz_store(sD, r, cg, fb_pack8(d1));
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
  VB_TL(tl_bwd_upper, 8);
      fb_issue(tmem + C_DCTX, D_k, WO_mn, H, H / 16, false);      // dctx[row, k] = sum_n ddelta1[row,n] Wo[n,k]
      fb_issue(tmem + C_WO, D_mn, CTX_mn, H, 8, iter > 0);        // dWo[n, k]  += sum_rows ddelta1[row,n] ctx[row,k]
      umma_commit(b_mma);
      tma_store_2d(&tmDh, sDz, 0, r0);   // dh leaves through TMA (row-strided st.global costs a tag lookup per row)
      tma_store_commit();
    }
    acc_bo += fb_colsum<8>(sD, tid & 31, (tid >> 5) * 8);
    mbar_wait(b_mma, ph_mma); ph_mma ^= 1;
  VB_TL(tl_bwd_upper, 9);
    tc_fence_after();
    {
      float v[HC];
      tmem_ld_32x8(my_tmem + C_DCTX + hc0, v);
      fb_swz_store(sCtx, r, cg, fb_pack8(v));   // the ctx tile is dead (MMA 3 done): it becomes the dctx store image
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();  // all reads of sD / TMEM done before the next tile overwrites them
    if (tid == 0) {
      tma_store_2d(&tmDctx, sCtx, 0, r0);
      tma_store_commit();
    }
  }

  VB_TL(tl_bwd_upper, 10);
  // ---- write this CTA's partial parameter gradients ----
  // TMEM rows are parameter rows: a direct st.global would touch one 128-byte line per lane.  The tiles go through
  // shared memory (16-byte pieces, XOR-swizzled so that neither side has bank conflicts) and leave as coalesced float4s.
  float* gp = P.gpart + (size_t)blockIdx.x * P.n_opt;
  tc_fence_after();
  float4* t_w1 = reinterpret_cast<float4*>(sM);            // [128 rows][8 pieces]   dW1 (16 KB)
  float4* t_w2 = reinterpret_cast<float4*>(sM + 16384);    // [32 rows][32 pieces]   dW2 (16 KB)
  float4* t_wo = reinterpret_cast<float4*>(sDA);           // [32 rows][8 pieces]    dWo (4 KB)
  float* t_b1 = reinterpret_cast<float*>(sDA + 4096);      // [128]                  db1
  float* red = reinterpret_cast<float*>(base + O_ACT);     // [2 * H][128]           LN gamma / beta partials (32 KB)
  float* bsum = reinterpret_cast<float*>(sDA + 8192);      // [2][512]               db2 / dbo partials
  if (iter > 0) {
    if ((warp & 3) == 0) {  // dW2 / dWo rows h = lanes 0..31: column group cg of each
      float v[32];
      tmem_ld_32x32(my_tmem + C_W2 + cg * 32, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) t_w2[r * 32 + ((cg * 8 + j) ^ (r & 31))] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      float w[HC];
      tmem_ld_32x8(my_tmem + C_WO + hc0, w);
      t_wo[r * 8 + ((2 * cg) ^ (r & 7))] = make_float4(w[0], w[1], w[2], w[3]);
      t_wo[r * 8 + ((2 * cg + 1) ^ (r & 7))] = make_float4(w[4], w[5], w[6], w[7]);
    }
    {  // dW1 rows i = lanes 0..127, H columns (+ db1 in column H)
      float v[HC];
      tmem_ld_32x8(my_tmem + C_W1 + hc0, v);
      t_w1[r * 8 + ((2 * cg) ^ (r & 7))] = make_float4(v[0], v[1], v[2], v[3]);
      t_w1[r * 8 + ((2 * cg + 1) ^ (r & 7))] = make_float4(v[4], v[5], v[6], v[7]);
      if (cg == 0) {
        float b8[8];
        tmem_ld_32x8(my_tmem + C_W1 + H, b8);
        t_b1[r] = b8[0];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < HC; ++j) { red[(hc0 + j) * 128 + r] = acc_g[j]; red[(H + hc0 + j) * 128 + r] = acc_b[j]; }
  bsum[tid] = acc_b2; bsum[512 + tid] = acc_bo;
  tc_fence_before();
  __syncthreads();
  {
    float4* o_w1 = reinterpret_cast<float4*>(gp + P.off_w1);
    float4* o_w2 = reinterpret_cast<float4*>(gp + P.off_w2);
    float4* o_wo = reinterpret_cast<float4*>(gp + P.off_wo);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);   // a CTA without tiles contributes zeros
    for (int e = tid; e < 128 * 8; e += FB_THREADS) { const int rr = e >> 3, c = e & 7; o_w1[e] = iter > 0 ? t_w1[rr * 8 + (c ^ (rr & 7))] : z4; }
    for (int e = tid; e < 32 * 32; e += FB_THREADS) { const int rr = e >> 5, c = e & 31; o_w2[e] = iter > 0 ? t_w2[rr * 32 + (c ^ (rr & 31))] : z4; }
    for (int e = tid; e < 32 * 8; e += FB_THREADS) { const int rr = e >> 3, c = e & 7; o_wo[e] = iter > 0 ? t_wo[rr * 8 + (c ^ (rr & 7))] : z4; }
    if (tid < I) gp[P.off_b1 + tid] = iter > 0 ? t_b1[tid] : 0.f;
  }
  fb_reduce_rows(red, 2 * H, [&](int e, float s) { if (e < H) gp[P.off_ln2g + e] = s; else gp[P.off_ln2b + e - H] = s; });
  if (tid >= 128 && tid < 128 + H) {
    const int c = tid - 128;
    float s2 = 0.f, so = 0.f;
    for (int k = 0; k < 16; ++k) { s2 += bsum[k * 32 + c]; so += bsum[512 + k * 32 + c]; }
    gp[P.off_b2 + c] = s2; gp[P.off_bo + c] = so;
  }
  __syncthreads();
  VB_TL(tl_bwd_upper, 11);
  if (tid == 0) tma_store_wait_all();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

// ================================================================================================
// lower: dqkv -> [Wqkv dgrad/wgrad] -> LN1 bwd (+ dh) -> dz_in
// ================================================================================================
__global__ void __launch_bounds__(FB_THREADS, 1)
fused_bwd_lower_kernel(const __grid_constant__ CUtensorMap tmDQ, const __grid_constant__ CUtensorMap tmU,
                       const __grid_constant__ CUtensorMap tmWq, const __grid_constant__ CUtensorMap tmZ,
                       const __grid_constant__ CUtensorMap tmDh, const __grid_constant__ CUtensorMap tmDzOut,
                       const vitb200_layer_bwd_lower_args P) {
  constexpr int H = FB_H, Q = 3 * FB_H, HC = FB_HC;
  constexpr uint32_t O_DQ = 0, O_U = 32768, O_WQ = O_U + 16384, O_Z = O_WQ + 12288, O_DH = O_Z + 16384,
                     O_BAR = O_DH + 16384, O_RED = O_BAR + 1024, O_EX = O_RED + 32768;
  constexpr uint32_t C_WQ = 0, C_DU = 64, TMEM_COLS = 128;   // dWqkv is 48 columns: column H = dbqkv (ones column in the u tile)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t *sDQ = base + O_DQ, *sU = base + O_U, *sWq = base + O_WQ;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + O_BAR);
  uint64_t *b_w = bars, *b_tile = bars + 1, *b_mma = bars + 2, *b_wg = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  float* red = reinterpret_cast<float*>(base + O_RED);  // [2][H][128] floats = 32 KB
  float2* s_ex = reinterpret_cast<float2*>(base + O_EX);
  const uint8_t *sZ = base + O_Z, *sDh = base + O_DH;   // fp32 rows of z and dh, TMA-staged
  float* s_g1 = reinterpret_cast<float*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int r = ((warp & 3) << 5) | (tid & 31);
  const int cg = warp >> 2, hc0 = cg * HC;
  const int M = P.B * P.T;
  const int ntiles = (M + 127) / 128;
  if (tid < H) s_g1[tid] = P.ln1_g[tid];
  if (tid == 0) {
    tma_prefetch_desc(&tmDQ); tma_prefetch_desc(&tmU); tma_prefetch_desc(&tmWq);
    tma_prefetch_desc(&tmZ); tma_prefetch_desc(&tmDh);
    mbar_init(b_w, 1); mbar_init(b_tile, 1); mbar_init(b_mma, 1); mbar_init(b_wg, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(b_w, 12288);
    tma_load_2d(sWq, &tmWq, b_w, 0, 0);  // Wqkv [3H rows, H cols]: {64 cols (32 valid), 96 rows}
  }
  pdl_wait();
  pdl_trigger();
  auto load_tile = [&](int r0) {   // thread 0: all TMA loads of one 128-row tile
    mbar_expect_tx(b_tile, 32768 + 16384 + 16384 + 16384);
    tma_load_2d(sDQ, &tmDQ, b_tile, 0, r0);
    tma_load_2d(sDQ + 16384, &tmDQ, b_tile, 64, r0);
    tma_load_2d(sU, &tmU, b_tile, 0, r0);
    tma_load_2d(base + O_Z, &tmZ, b_tile, 0, r0);
    tma_load_2d(base + O_DH, &tmDh, b_tile, 0, r0);
  };
  if (tid == 0 && (int)blockIdx.x < ntiles) load_tile(blockIdx.x * 128);   // first tile: in flight while TMEM is allocated
  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t my_tmem = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const Opnd DQ_k{smem_u32(sDQ), 16, 16384, 0};        // dqkv tile, K-major (K = 3H = 96: 6 k-steps over 2 k-blocks)
  const Opnd DQ_mn{smem_u32(sDQ), 16384, 0, 1};        // MN-major (MN = 96 of 128)
  const Opnd U_mn{smem_u32(sU), 16384, 0, 1};
  const Opnd WQ_mn{smem_u32(sWq), 16384, 0, 1};        // B(n = h, k = qkv row): N = H, K = 96 rows

  float acc_g[HC], acc_b[HC];
#pragma unroll
  for (int j = 0; j < HC; ++j) { acc_g[j] = 0.f; acc_b[j] = 0.f; }
  uint32_t ph_mma = 0, ph_tile = 0;
  int iter = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++iter) {
    const int r0 = tile * 128, row = r0 + r;
    const bool valid = row < M;
    const int rowc = valid ? row : M - 1;
    const float mu = P.mean1[rowc], rs = P.rstd1[rowc];  // issued early
    if (tid == 0) {
      if (iter > 0) {
        tma_store_wait_read<0>();   // the previous tile's dz image (sDh) is about to be overwritten
        load_tile(r0);
      }
      if (iter == 0) mbar_wait(b_w, 0);
      mbar_wait(b_tile, ph_tile);
      tc_fence_after();
      fb_issue(tmem + C_DU, DQ_k, WQ_mn, H, Q / 16, false);      // du[row, h]   = sum_n dqkv[row,n] Wqkv[n,h]
      umma_commit(b_mma);                                         // the LayerNorm stage only needs du
    }
    mbar_wait(b_tile, ph_tile);  // every thread reads the TMA-written z / dh rows below
    ph_tile ^= 1;
    // a column of ones next to u (columns H..63 are TMA zero fill): the dWqkv MMA then also emits dbqkv in column H
    if (cg == 0) *reinterpret_cast<uint32_t*>(const_cast<uint8_t*>(fb_swz_ptr(sU, r, H / 8))) = 0x00003F80u;   // bf16 {1, 0}
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      fb_issue(tmem + C_WQ, DQ_mn, U_mn, H + 16, 8, iter > 0);   // dWqkv[n, h] += sum_rows dqkv[row,n] u[row,h] ; column H: dbqkv[n]
      umma_commit(b_wg);
    }
    mbar_wait(b_mma, ph_mma);
    tc_fence_after();
    {
      float du[HC], xh[HC], g[HC], dz[HC];
      tmem_ld_32x8(my_tmem + C_DU + hc0, du);
      fb_ld_f8(sZ, r, cg, xh);
      fb_ld_f8(sDh, r, cg, dz);
#pragma unroll
      for (int j = 0; j < HC; ++j) {
        xh[j] = (xh[j] - mu) * rs;
        du[j] = valid ? bf16_round(du[j]) : 0.f;
        acc_g[j] = fmaf(du[j], xh[j], acc_g[j]);
        acc_b[j] += du[j];
      }
      float c1, c2;
      fb_ln_bwd_sums(du, xh, s_g1, hc0, s_ex, r, cg, g, c1, c2);
#pragma unroll
      for (int j = 0; j < HC / 4; ++j) {   // dz overwrites this thread's piece of the dh tile: the tile is the store image
        float4 o;
        o.x = dz[4 * j] + rs * (g[4 * j] - c1 - xh[4 * j] * c2);
        o.y = dz[4 * j + 1] + rs * (g[4 * j + 1] - c1 - xh[4 * j + 1] * c2);
        o.z = dz[4 * j + 2] + rs * (g[4 * j + 2] - c1 - xh[4 * j + 2] * c2);
        o.w = dz[4 * j + 3] + rs * (g[4 * j + 3] - c1 - xh[4 * j + 3] * c2);
        *reinterpret_cast<float4*>(const_cast<uint8_t*>(fb_swz_ptr(sDh, r, 2 * cg + j))) = o;
      }
    }
    fence_proxy_async();
    tc_fence_before();
    mbar_wait(b_wg, ph_mma); ph_mma ^= 1;   // the wgrad MMA has finished reading the dqkv / u tiles (long done by now)
    __syncthreads();
    if (tid == 0) {
      tma_store_2d(&tmDzOut, sDh, 0, r0);
      tma_store_commit();
    }
  }
  float* gp = P.gpart + (size_t)blockIdx.x * P.n_opt;
  tc_fence_after();
  float4* t_wq = reinterpret_cast<float4*>(sDQ);            // [96 rows][8 pieces] dWqkv (12 KB; the dqkv tile is dead)
  float* t_bq = reinterpret_cast<float*>(sDQ + 16384);      // [96] dbqkv
  if (iter > 0 && (warp & 3) < 3) {  // dWqkv rows n = lanes 0..95
    float v[HC];
    tmem_ld_32x8(my_tmem + C_WQ + hc0, v);
    t_wq[r * 8 + ((2 * cg) ^ (r & 7))] = make_float4(v[0], v[1], v[2], v[3]);
    t_wq[r * 8 + ((2 * cg + 1) ^ (r & 7))] = make_float4(v[4], v[5], v[6], v[7]);
    if (cg == 0) {
      float b8[8];
      tmem_ld_32x8(my_tmem + C_WQ + H, b8);
      t_bq[r] = b8[0];
    }
  }
#pragma unroll
  for (int j = 0; j < HC; ++j) { red[(hc0 + j) * 128 + r] = acc_g[j]; red[(H + hc0 + j) * 128 + r] = acc_b[j]; }
  tc_fence_before();
  __syncthreads();
  {
    float4* o_wq = reinterpret_cast<float4*>(gp + P.off_wqkv);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = tid; e < Q * 8; e += FB_THREADS) { const int rr = e >> 3, c = e & 7; o_wq[e] = iter > 0 ? t_wq[rr * 8 + (c ^ (rr & 7))] : z4; }
    if (tid < Q) gp[P.off_bqkv + tid] = iter > 0 ? t_bq[tid] : 0.f;
  }
  fb_reduce_rows(red, 2 * H, [&](int e, float s) { if (e < H) gp[P.off_ln1g + e] = s; else gp[P.off_ln1b + e - H] = s; });
  __syncthreads();
  if (tid == 0) tma_store_wait_all();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

// ================================================================================================
// embedding backward: dz0 -> dropout' -> dW_p = dtok^T . patches, db_p, dcls
// ================================================================================================
__global__ void __launch_bounds__(FB_THREADS, 1)
fused_embed_bwd_kernel(const vitb200_embed_bwd_args P) {
  constexpr int H = FB_H, HC = FB_HC;
  constexpr uint32_t O_D = 0, O_X = 16384, O_BAR = 32768, O_RED = O_BAR + 1024;
  constexpr uint32_t TMEM_COLS = 64;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t *sD = base + O_D, *sX = base + O_X;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + O_BAR);
  uint64_t* b_mma = bars;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1);
  float* red = reinterpret_cast<float*>(base + O_RED);  // [H][128]

  const int tid = threadIdx.x, warp = tid >> 5;
  const int r = ((warp & 3) << 5) | (tid & 31);
  const int cg = warp >> 2, hc0 = cg * HC;
  const int T = P.Np + 1, M = P.B * T;
  const int ntiles = (M + 127) / 128;
  if (tid == 0) { mbar_init(b_mma, 1); fence_barrier_init(); }
  // zero both tiles once: unused columns must be finite for the MN-major views
  fb_swz_store(sD, r, cg, make_uint4(0u, 0u, 0u, 0u));     fb_swz_store(sD, r, 4 + cg, make_uint4(0u, 0u, 0u, 0u));
  fb_swz_store(sX, r, cg, make_uint4(0u, 0u, 0u, 0u));     fb_swz_store(sX, r, 4 + cg, make_uint4(0u, 0u, 0u, 0u));
  pdl_wait();
  pdl_trigger();
  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t my_tmem = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const Opnd D_mn{smem_u32(sD), 16384, 0, 1};   // MN group 1 aliases sX: product rows 64..127 are never read
  const Opnd X_mn{smem_u32(sX), 16384, 0, 1};
  const DropCtx dc = make_drop(P.p_drop, P.rng ? P.rng[0] : 0ull, P.rng ? (uint32_t)P.rng[1] : 0u, VITB200_SITE_EMB);
  float acc_cls[HC];
#pragma unroll
  for (int j = 0; j < HC; ++j) acc_cls[j] = 0.f;
  float acc_bp = 0.f;  // column tid & 31, rows 8 (tid >> 5) .. +7
  uint32_t ph = 0;
  int iter = 0;
  const int nchunk = P.P / 8;
  const bool vec = (P.S % 4 == 0) && (P.L % 4 == 0);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++iter) {
    const int row = tile * 128 + r;
    const bool valid = row < M;
    const int rowc = valid ? row : M - 1;
    const int b = rowc / T, t = rowc - b * T;
    float g[HC];
    {
      const float4* p = reinterpret_cast<const float4*>(P.dz0 + (size_t)rowc * H + hc0);
      const float4 v0 = p[0], v1 = p[1];
      float kp[8];
      drop8(dc, ((size_t)rowc * H + hc0) >> 3, kp);
      g[0] = v0.x * kp[0]; g[1] = v0.y * kp[1]; g[2] = v0.z * kp[2]; g[3] = v0.w * kp[3];
      g[4] = v1.x * kp[4]; g[5] = v1.y * kp[5]; g[6] = v1.z * kp[6]; g[7] = v1.w * kp[7];
#pragma unroll
      for (int j = 0; j < HC; ++j) g[j] = valid ? g[j] : 0.f;
    }
    if (t == 0) {  // CLS row: gradient of cls_token, no patch
#pragma unroll
      for (int j = 0; j < HC; ++j) { acc_cls[j] += g[j]; g[j] = 0.f; }
    }
    fb_swz_store(sD, r, cg, fb_pack8(g));
    {
      const bool has = valid && t >= 1 && (t - 1) < P.n_valid;
      const float* xp = P.x + (size_t)b * P.L + (size_t)(t >= 1 ? t - 1 : 0) * P.S;
      for (int c = cg; c < nchunk; c += FB_CG) {
        float v[8];
        if (vec && has) {
          const float4 a0 = *reinterpret_cast<const float4*>(xp + c * 8), a1 = *reinterpret_cast<const float4*>(xp + c * 8 + 4);
          v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = has ? xp[c * 8 + q] : 0.f;
        }
        fb_swz_store(sX, r, c, fb_pack8(v));
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      fb_issue(tmem, D_mn, X_mn, P.P, 8, iter > 0);   // dWp[h, j] += sum_rows dtok[row, h] x[row, j]
      umma_commit(b_mma);
    }
    acc_bp += fb_colsum<8>(sD, tid & 31, (tid >> 5) * 8);
    mbar_wait(b_mma, ph); ph ^= 1;
    tc_fence_after();
    tc_fence_before();
    __syncthreads();
  }
  float* gp = P.gpart + (size_t)blockIdx.x * P.n_opt;
  tc_fence_after();
  if ((warp & 3) == 0) {  // dWp rows h = lanes 0..31
    if (iter > 0) {
#pragma unroll 1
      for (int c0 = hc0; c0 < P.P; c0 += 8 * FB_CG) {
        float v[8];
        tmem_ld_32x8(my_tmem + c0, v);
        float4* op = reinterpret_cast<float4*>(gp + P.off_wp + (size_t)r * P.P + c0);
        op[0] = make_float4(v[0], v[1], v[2], v[3]); op[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
    } else {
      for (int e = hc0; e < P.P; e += 8 * FB_CG)
        for (int q = 0; q < 8; ++q) gp[P.off_wp + (size_t)r * P.P + e + q] = 0.f;
    }
  }
#pragma unroll
  for (int j = 0; j < HC; ++j) red[(hc0 + j) * 128 + r] = acc_cls[j];
  float* bsum = reinterpret_cast<float*>(sX);  // [16][32]
  tc_fence_before();
  __syncthreads();
  bsum[tid] = acc_bp;
  __syncthreads();
  fb_reduce_rows(red, H, [&](int e, float s) { gp[P.off_cls + e] = s; });
  if (tid < H) {
    float s = 0.f;
    for (int k = 0; k < 16; ++k) s += bsum[k * 32 + tid];
    gp[P.off_bp + tid] = s;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}
constexpr int EMBED_BWD_SMEM = 16384 + 16384 + 1024 + 16384 + 1024;

// grad[i] = sum over slots (in slot order) of gpart[s*stride + i]
__global__ void __launch_bounds__(256)
grad_reduce_kernel(const float* __restrict__ gpart, int slots, size_t stride, size_t start, size_t n4,
                   float* __restrict__ grad) {
  pdl_wait();
  pdl_trigger();
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
    const float4* src = reinterpret_cast<const float4*>(gpart + start) + i;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    int z = 0;
    for (; z + 4 <= slots; z += 4) {
      float4 t[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) t[q] = __ldcg(src + (size_t)(z + q) * (stride / 4));
#pragma unroll
      for (int q = 0; q < 4; ++q) { s.x += t[q].x; s.y += t[q].y; s.z += t[q].z; s.w += t[q].w; }
    }
    for (; z < slots; ++z) {
      const float4 t = __ldcg(src + (size_t)z * (stride / 4));
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    reinterpret_cast<float4*>(grad + start)[i] = s;
  }
}

constexpr int UPPER_SMEM = 16384 + 32768 + 32768 + 16384 + 16384 + 8192 + 16384 + 4096 + 32768 + 16384 + 16384 + 1024 + 4096 + 1024;
constexpr int LOWER_SMEM = 32768 + 16384 + 12288 + 16384 + 16384 + 1024 + 32768 + 4096 + 1024;

}  // namespace vb

using namespace vb;

VB_TL_EXPORT(vitb200_tl_bwd_upper, vb::tl_bwd_upper)

extern "C" int vitb200_fused_bwd_supported(int H) { return H == FB_H ? 1 : 0; }
extern "C" int vitb200_fused_bwd_grid(int M) {
  int tiles = (M + 127) / 128;
  return tiles < 148 ? (tiles < 1 ? 1 : tiles) : 148;
}

extern "C" int vitb200_fused_layer_bwd_upper(const vitb200_layer_bwd_upper_args* a, void* stream) {
  if (!a || (!a->dz && !a->dz_cls) || !a->m || !a->a || !a->u2 || !a->ctx || !a->hmid || !a->mean2 || !a->rstd2 || !a->ln2_g ||
      !a->w_2 || !a->w_1 || !a->w_o || !a->dh || !a->dctx || !a->gpart)
    return VITB200_ERR_ARG;
  if (a->H != FB_H) return VITB200_ERR_SHAPE;
  if (a->B <= 0 || a->T <= 0) return VITB200_ERR_ARG;
  const int M = a->B * a->T, H = FB_H, I = FB_I;
  CUtensorMap tM, tU2, tCtx, tW2, tW1, tWo, tAct, tHm;
  int rc;
  if ((rc = get_tmap(a->a, I, M, 64, 128, &tAct))) return rc;
  if ((rc = get_tmap(a->hmid, 2 * H, M, 64, 128, &tHm))) return rc;  // fp32 [M, H] rows viewed as 2H bf16 (128 B)
  if ((rc = get_tmap(a->m, I, M, 64, 128, &tM))) return rc;
  if ((rc = get_tmap(a->u2, H, M, 64, 128, &tU2))) return rc;
  if ((rc = get_tmap(a->ctx, H, M, 64, 128, &tCtx))) return rc;
  if ((rc = get_tmap(a->w_2, I, H, 64, H, &tW2))) return rc;
  if ((rc = get_tmap(a->w_1, H, I, 64, I, &tW1))) return rc;
  if ((rc = get_tmap(a->w_o, H, H, 64, H, &tWo))) return rc;
  CUtensorMap tDz, tDh, tDctx;
  if ((rc = get_tmap(a->dh, 2 * H, M, 64, 128, &tDh))) return rc;
  if (a->dz) { if ((rc = get_tmap(a->dz, 2 * H, M, 64, 128, &tDz))) return rc; }
  else tDz = tDh;
  if ((rc = get_tmap(a->dctx, H, M, 64, 128, &tDctx))) return rc;
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(fused_bwd_upper_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, UPPER_SMEM);
    if (e != cudaSuccess) return vb_cuda_error(e);
    done = true;
  }
  vb_launch_pdl(fused_bwd_upper_kernel, dim3(vitb200_fused_bwd_grid(M)), dim3(FB_THREADS), UPPER_SMEM, (cudaStream_t)stream,
                tM, tU2, tCtx, tW2, tW1, tWo, tAct, tHm, tDz, tDh, tDctx, *a);
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

extern "C" int vitb200_fused_layer_bwd_lower(const vitb200_layer_bwd_lower_args* a, void* stream) {
  if (!a || !a->dqkv || !a->u || !a->z || !a->mean1 || !a->rstd1 || !a->ln1_g || !a->dh || !a->w_qkv || !a->dz || !a->gpart)
    return VITB200_ERR_ARG;
  if (a->H != FB_H) return VITB200_ERR_SHAPE;
  if (a->B <= 0 || a->T <= 0) return VITB200_ERR_ARG;
  const int M = a->B * a->T, H = FB_H;
  CUtensorMap tDQ, tU, tWq, tZ, tDh;
  int rc;
  if ((rc = get_tmap(a->z, 2 * H, M, 64, 128, &tZ))) return rc;    // fp32 rows viewed as 2H bf16
  if ((rc = get_tmap(a->dh, 2 * H, M, 64, 128, &tDh))) return rc;
  if ((rc = get_tmap(a->dqkv, 3 * H, M, 64, 128, &tDQ))) return rc;
  if ((rc = get_tmap(a->u, H, M, 64, 128, &tU))) return rc;
  if ((rc = get_tmap(a->w_qkv, H, 3 * H, 64, 3 * H, &tWq))) return rc;
  CUtensorMap tDzOut;
  if ((rc = get_tmap(a->dz, 2 * H, M, 64, 128, &tDzOut))) return rc;
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(fused_bwd_lower_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LOWER_SMEM);
    if (e != cudaSuccess) return vb_cuda_error(e);
    done = true;
  }
  vb_launch_pdl(fused_bwd_lower_kernel, dim3(vitb200_fused_bwd_grid(M)), dim3(FB_THREADS), LOWER_SMEM, (cudaStream_t)stream,
                tDQ, tU, tWq, tZ, tDh, tDzOut, *a);
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

extern "C" int vitb200_fused_embed_bwd_supported(int H, int P, int learned_pos) {
  return (H == FB_H && !learned_pos && P % 16 == 0 && P >= 16 && P <= 64) ? 1 : 0;
}

extern "C" int vitb200_fused_embed_bwd(const vitb200_embed_bwd_args* a, void* stream) {
  if (!a || !a->dz0 || !a->x || !a->gpart) return VITB200_ERR_ARG;
  if (!vitb200_fused_embed_bwd_supported(a->H, a->P, 0)) return VITB200_ERR_SHAPE;
  if (a->B <= 0 || a->Np <= 0) return VITB200_ERR_ARG;
  const int M = a->B * (a->Np + 1);
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(fused_embed_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, EMBED_BWD_SMEM);
    if (e != cudaSuccess) return vb_cuda_error(e);
    done = true;
  }
  vb_launch_pdl(fused_embed_bwd_kernel, dim3(vitb200_fused_bwd_grid(M)), dim3(FB_THREADS), EMBED_BWD_SMEM, (cudaStream_t)stream, *a);
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

extern "C" int vitb200_grad_reduce(const float* gpart, int slots, size_t stride, size_t start, size_t end, float* grad,
                                   void* stream) {
  if (!gpart || !grad || slots <= 0 || end < start) return VITB200_ERR_ARG;
  if ((stride | start | end) % 4 != 0) return VITB200_ERR_SHAPE;
  if (end == start) return VITB200_OK;
  const size_t n4 = (end - start) / 4;
  size_t g = (n4 + 255) / 256;
  if (g > 592) g = 592;
  vb_launch_pdl(grad_reduce_kernel, dim3((int)g), dim3(256), 0, (cudaStream_t)stream, gpart, slots, stride, start, n4, grad);
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}
