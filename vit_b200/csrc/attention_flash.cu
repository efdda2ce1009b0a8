// Flash attention on tcgen05 for ANY sequence length (bf16, head_dim 16 / 32 / 64): ONE launch, the key/value loop runs
// inside the kernel with an online softmax -- no per-key-block launches, no partial outputs in HBM, no merge kernel.
//
// One CTA per (query tile of 128 rows, head, sample): 16 warps, warp w owns TMEM lane quarter (w & 3) = query rows
// 32 (w & 3) .. +31 and key-column group cg = w >> 2 (a query row's softmax is split over 4 threads, as in
// attention_tc.cu).  Per block of 128 keys:
//     S = Q K^T (UMMA, fp32 in TMEM)  ->  running row max m, alpha = exp(m_old - m_new)
//     P~ = dropout(exp(S - m))  (bf16, swizzled smem)   l = l alpha + rowsum(exp(S - m))
//     O_blk = P~ V  (UMMA, V as MN-major B)             o = o alpha + O_blk      <- the running output lives in REGISTERS
// (head_dim <= 64 means at most 16 output columns per thread, so nothing in TMEM is ever rescaled).
// K and V blocks are double-buffered TMA loads; S is double-buffered in TMEM: the scores of block k + 1 are issued while
// block k is still in its exp pass, and its P~ V product runs under the max pass of block k + 1.
// Dropout masks, RoPE and the log-sum-exp output are the ones of attention_tc.cu / attention.cu (same element indices),
// so the per-key-block backward and the SIMT kernels can be mixed with it.
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_host.cuh"

namespace vb {
using namespace vb::tc;

constexpr int FA_CG = 4;
constexpr int FA_THREADS = 128 * FA_CG;
constexpr float FA_LOG2E = 1.4426950408889634f;
constexpr uint32_t FA_Q = 0, FA_K = 16384, FA_V = FA_K + 2 * 16384, FA_P = FA_V + 2 * 16384, FA_BAR = FA_P + 32768,
                   FA_MX = FA_BAR + 256, FA_SM = FA_MX + 2048, FA_SMEM = FA_SM + 2048 + 1024;
constexpr uint32_t FC_S0 = 0, FC_S1 = 128, FC_O = 256, FC_COLS = 512;

__device__ __forceinline__ uint4 fa_pack8(const float* v) {
  __nv_bfloat162 t0 = __floats2bfloat162_rn(v[0], v[1]), t1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 t2 = __floats2bfloat162_rn(v[4], v[5]), t3 = __floats2bfloat162_rn(v[6], v[7]);
  uint4 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
  pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
  return pk;
}
__device__ __forceinline__ void fa_unpack8(uint4 pk, float* v) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
  for (int q = 0; q < 4; ++q) { float2 f = __bfloat1622float2(p[q]); v[2 * q] = f.x; v[2 * q + 1] = f.y; }
}
__device__ __forceinline__ uint8_t* fa_swz(uint8_t* tile, int r, int chunk) {
  return tile + (chunk >> 3) * 16384 + r * 128 + (((chunk & 7) ^ (r & 7)) << 4);
}
// RoPE on the D values of row r that start at chunk c0 of a swizzled 64-column block (rope.py:60-98)
template <int D>
__device__ __forceinline__ void fa_rope_row(uint8_t* blk, int r, int c0, const float* __restrict__ cosT, const float* __restrict__ sinT, int t) {
  float x[D];
#pragma unroll
  for (int c = 0; c < D / 8; ++c) fa_unpack8(*reinterpret_cast<const uint4*>(fa_swz(blk, r, c0 + c)), &x[c * 8]);
#pragma unroll
  for (int c = 0; c < D / 2; ++c) {
    const float cs = cosT[(size_t)t * (D / 2) + c], sn = sinT[(size_t)t * (D / 2) + c];
    const float lo = x[c], hi = x[c + D / 2];
    x[c] = lo * cs - hi * sn;
    x[c + D / 2] = hi * cs + lo * sn;
  }
#pragma unroll
  for (int c = 0; c < D / 8; ++c) *reinterpret_cast<uint4*>(fa_swz(blk, r, c0 + c)) = fa_pack8(&x[c * 8]);
}
__device__ __forceinline__ void fa_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

struct FlashParams {
  bf16* ctx; float* lse;
  const float* cosT; const float* sinT;
  int B, T, heads, H;
  float scale, p_drop; const uint64_t* rng; uint32_t site;
};

template <int D>
__global__ void __launch_bounds__(FA_THREADS, 1)
attn_flash_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const FlashParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t *sQ = base + FA_Q, *sK = base + FA_K, *sV = base + FA_V, *sP = base + FA_P;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + FA_BAR);
  uint64_t *b_q = bars, *b_k = bars + 1 /* [2] */, *b_v = bars + 3 /* [2] */, *b_s = bars + 5 /* [2] */, *b_pv = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  float* s_mx = reinterpret_cast<float*>(base + FA_MX);   // [4][128]
  float* s_sm = reinterpret_cast<float*>(base + FA_SM);   // [4][128]

  const int tid = threadIdx.x, warp = tid >> 5;
  const int r = ((warp & 3) << 5) | (tid & 31);
  const int cg = warp >> 2;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z, T = P.T;
  const int row0 = b * T, q0 = qt * 128, i = q0 + r;
  const bool valid = i < T;
  const int nkb = (T + 127) >> 7;
  const int colQ = h * D, colK = P.H + h * D, colV = 2 * P.H + h * D;
  const int cq = (colQ & 63) >> 3, ck = (colK & 63) >> 3, cv = (colV & 63) >> 3;   // chunk of the head inside the 64-column box

  if (tid == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int k = 0; k < 8; ++k) mbar_init(bars + k, 1);
    fence_barrier_init();
  }
  pdl_wait();     // q/k/v come from the previous kernel of the step
  pdl_trigger();
  auto load_k = [&](int kb) { mbar_expect_tx(b_k + (kb & 1), 16384); tma_load_2d(sK + (kb & 1) * 16384, &tmQKV, b_k + (kb & 1), colK & ~63, row0 + kb * 128); };
  auto load_v = [&](int kb) { mbar_expect_tx(b_v + (kb & 1), 16384); tma_load_2d(sV + (kb & 1) * 16384, &tmQKV, b_v + (kb & 1), colV & ~63, row0 + kb * 128); };
  if (tid == 0) {
    mbar_expect_tx(b_q, 16384);
    tma_load_2d(sQ, &tmQKV, b_q, colQ & ~63, row0 + q0);
    load_k(0); load_v(0);
    if (nkb > 1) { load_k(1); load_v(1); }
  }
  if (warp == 0) tmem_alloc(tmem_slot, FC_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t my_tmem = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t aQ = smem_u32(sQ) + cq * 16, aK = smem_u32(sK) + ck * 16, aV = smem_u32(sV) + cv * 16, aP = smem_u32(sP);
  const uint32_t id_s = make_idesc_bf16(128, 128, 0, 0), id_o = make_idesc_bf16(128, D, 0, 1);
  auto issue_s = [&](int kb) {   // thread 0: S(kb) = Q K(kb)^T
    for (int k = 0; k < D / 16; ++k)
      umma_bf16(tmem + ((kb & 1) ? FC_S1 : FC_S0), make_sdesc_sw128(aQ + k * 32, 16, 1024),
                make_sdesc_sw128(aK + (kb & 1) * 16384 + k * 32, 16, 1024), id_s, k > 0 ? 1u : 0u);
    umma_commit(b_s + (kb & 1));
  };
  const DropCtx dc = make_drop(P.p_drop, P.rng ? P.rng[0] : 0ull, P.rng ? (uint32_t)P.rng[1] : 0u, P.site);
  const int Tpad = attn_drop_tpad(T);
  const float sl2 = P.scale * FA_LOG2E;
  const uint64_t drow = ((uint64_t)(b * P.heads + h) * T + (valid ? i : 0)) * (uint64_t)Tpad;

  // ---- Q (and the first K block) ready; RoPE rotates rows in place ----
  mbar_wait(b_q, 0);
  if (P.cosT) {
    if (cg == 0) fa_rope_row<D>(sQ, r, cq, P.cosT, P.sinT, valid ? i : 0);
    mbar_wait(b_k, 0);
    if (cg == 1) fa_rope_row<D>(sK, r, ck, P.cosT, P.sinT, r < T ? r : 0);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    mbar_wait(b_k, 0);
    issue_s(0);
  }
  float m_run = -INFINITY, l_run = 0.f, alpha = 1.f, sum_blk = 0.f;
  float o[8];   // this thread's output columns (cg < D / 8): columns 8 cg .. 8 cg + 7;  D = 64: groups 0..3 take 16 each
  float o2[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) { o[c] = 0.f; o2[c] = 0.f; }
  uint32_t ph_pv = 0;

  for (int kb = 0; kb < nkb; ++kb) {
    const int st = kb & 1;
    const uint32_t cS = st ? FC_S1 : FC_S0;
    const int key0 = kb * 128, Tk = T - key0 < 128 ? T - key0 : 128;
    // ---- O of the previous block: o = o * alpha + O_blk ; l = l * alpha + rowsum ----
    if (kb > 0) {
      mbar_wait(b_pv, ph_pv); ph_pv ^= 1;
      tc_fence_after();
      if (D <= 32) {
        if (cg < D / 8) {
          float ob[8];
          tmem_ld_32x8(my_tmem + FC_O + cg * 8, ob);
#pragma unroll
          for (int c = 0; c < 8; ++c) o[c] = fmaf(o[c], alpha, ob[c]);
        }
      } else {
        float ob[16];
        tmem_ld_32x16(my_tmem + FC_O + cg * 16, ob);
#pragma unroll
        for (int c = 0; c < 8; ++c) { o[c] = fmaf(o[c], alpha, ob[c]); o2[c] = fmaf(o2[c], alpha, ob[8 + c]); }
      }
      l_run = fmaf(l_run, alpha, sum_blk);
      if (tid == 0 && kb + 1 < nkb) load_v(kb + 1);   // the V stage of block kb - 1 is free (its P~ V product is done)
    }
    // ---- scores of this block ----
    mbar_wait(b_s + st, (uint32_t)((kb >> 1) & 1));
    tc_fence_after();
    float mx = -INFINITY;
#pragma unroll
    for (int k2 = 0; k2 < 2; ++k2) {
      const int c0 = (cg + k2 * FA_CG) * 16;
      float v[16];
      tmem_ld_32x16(my_tmem + cS + c0, v);
      if (c0 + 16 <= Tk) {
#pragma unroll
        for (int j = 0; j < 16; ++j) mx = fmaxf(mx, v[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) if (c0 + j < Tk) mx = fmaxf(mx, v[j]);
      }
    }
    s_mx[cg * 128 + r] = mx;
    tc_fence_before();
    fa_bar();
    // all reads of O_blk (previous block) and of s_sm are done; the K stage of block kb is free once S(kb) completed
    if (tid == 0) {
      if (kb + 1 < nkb) {
        if (!P.cosT) {
          tc_fence_after();
          mbar_wait(b_k + (st ^ 1), (uint32_t)(((kb + 1) >> 1) & 1));
          issue_s(kb + 1);                       // scores of the next block run under this block's exp pass
        }
      }
    }
    mx = fmaxf(fmaxf(s_mx[r], s_mx[128 + r]), fmaxf(s_mx[256 + r], s_mx[384 + r]));
    const float m_new = fmaxf(m_run, mx);          // (a row always has at least one live key in block 0)
    alpha = ex2_approx((m_run - m_new) * sl2);      // exp2(-inf) = 0 for the first block
    m_run = m_new;
    const float nmxs = -m_new * sl2;
    float sum = 0.f;
#pragma unroll
    for (int k2 = 0; k2 < 2; ++k2) {
      const int c0 = (cg + k2 * FA_CG) * 16;
      float v[16];
      tmem_ld_32x16(my_tmem + cS + c0, v);
      if (c0 + 16 <= Tk) {
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
          float kp[8];
          drop8(dc, (drow + (uint64_t)(key0 + c0 + j)) >> 3, kp);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float p = ex2_approx(fmaf(v[j + q], sl2, nmxs));
            sum += p;
            v[j + q] = p * kp[q];
          }
          *reinterpret_cast<uint4*>(fa_swz(sP, r, (c0 + j) >> 3)) = fa_pack8(&v[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
          float kp[8];
          if (c0 + j < Tk) drop8(dc, (drow + (uint64_t)(key0 + c0 + j)) >> 3, kp);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float p = (c0 + j + q < Tk) ? ex2_approx(fmaf(v[j + q], sl2, nmxs)) : 0.f;
            sum += p;
            v[j + q] = (c0 + j < Tk) ? p * kp[q] : 0.f;
          }
          *reinterpret_cast<uint4*>(fa_swz(sP, r, (c0 + j) >> 3)) = fa_pack8(&v[j]);
        }
      }
    }
    s_sm[cg * 128 + r] = sum;
    if (P.cosT && kb + 1 < nkb) {   // RoPE: the next K block is rotated in place before its scores are issued
      mbar_wait(b_k + (st ^ 1), (uint32_t)(((kb + 1) >> 1) & 1));
      if (cg == 1) fa_rope_row<D>(sK + (st ^ 1) * 16384, r, ck, P.cosT, P.sinT, key0 + 128 + r < T ? key0 + 128 + r : 0);
    }
    fence_proxy_async();
    tc_fence_before();
    fa_bar();
    sum_blk = (s_sm[r] + s_sm[128 + r]) + (s_sm[256 + r] + s_sm[384 + r]);
    if (tid == 0) {
      tc_fence_after();
      if (P.cosT && kb + 1 < nkb) issue_s(kb + 1);
      mbar_wait(b_v + st, (uint32_t)((kb >> 1) & 1));
      for (int k = 0; k < 8; ++k)               // O_blk[i, c] = sum_j P~[i, j] v[j, c]
        umma_bf16(tmem + FC_O, make_sdesc_sw128(aP + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                  make_sdesc_sw128(aV + st * 16384 + k * 2048, 16384, 1024), id_o, k > 0 ? 1u : 0u);
      umma_commit(b_pv);
      if (kb + 2 < nkb) load_k(kb + 2);          // S(kb) is complete: its K stage takes block kb + 2
    }
  }
  // ---- last block's output, normalisation, stores ----
  mbar_wait(b_pv, ph_pv);
  tc_fence_after();
  l_run = fmaf(l_run, alpha, sum_blk);
  const float inv = 1.f / l_run;
  if (D <= 32) {
    if (cg < D / 8) {
      float ob[8];
      tmem_ld_32x8(my_tmem + FC_O + cg * 8, ob);
#pragma unroll
      for (int c = 0; c < 8; ++c) o[c] = fmaf(o[c], alpha, ob[c]) * inv;
      if (valid) *reinterpret_cast<uint4*>(P.ctx + (size_t)(row0 + i) * P.H + h * D + cg * 8) = fa_pack8(o);
    }
  } else {
    float ob[16];
    tmem_ld_32x16(my_tmem + FC_O + cg * 16, ob);
#pragma unroll
    for (int c = 0; c < 8; ++c) { o[c] = fmaf(o[c], alpha, ob[c]) * inv; o2[c] = fmaf(o2[c], alpha, ob[8 + c]) * inv; }
    if (valid) {
      uint4* dst = reinterpret_cast<uint4*>(P.ctx + (size_t)(row0 + i) * P.H + h * D + cg * 16);
      dst[0] = fa_pack8(o); dst[1] = fa_pack8(o2);
    }
  }
  if (valid && cg == 0) P.lse[(size_t)(b * P.heads + h) * T + i] = m_run * P.scale + logf(l_run);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, FC_COLS);
}


// ================================================================================================
// backward: ONE launch for any sequence length.  One CTA per (block of 128 keys, head, sample); it keeps K, V of its block
// in shared memory and dK, dV in TMEM and walks over all query tiles:
//     S = Q K^T, dP = dO V^T (UMMA)  ->  P = exp(S - lse), P~ = dropout(P), dS = P~ o dP - P D_i   (D_i = dO_i . O_i)
//     dQ_blk = dS K        -> fp32 partial of this key block (dq_part[kb]); attn_flash_dq_kernel sums the key blocks in
//                             block order (deterministic, no float atomics), rounds to bf16 and un-rotates (RoPE)
//     dK += dS^T Q, dV += P~^T dO    (MN-major A views of the dS / P~ tiles, accumulators resident in TMEM)
// Q / dO tiles are double-buffered TMA loads; the S / dP products of tile qt + 1 are issued right behind the dQ / dK / dV
// products of tile qt, so they run under the dQ epilogue and are usually complete when the next pass starts.
// ================================================================================================
constexpr uint32_t FB_Q = 0, FB_DO = 32768, FB_K = 65536, FB_V = FB_K + 16384, FB_DS = FB_V + 16384, FB_PT = FB_DS + 32768,
                   FB_BAR = FB_PT + 32768, FB_SMEM = FB_BAR + 256 + 1024;
constexpr uint32_t FBC_S = 0, FBC_DP = 128, FBC_DQ = 256, FBC_DK = 320, FBC_DV = 384, FBC_COLS = 512;

struct FOp { uint32_t addr, lbo, kblk; int mn; };   // shared-memory operand view: K-major (k-blocks of 64) or MN-major
__device__ __forceinline__ uint64_t fop_desc(const FOp& o, int k) {
  if (o.mn) return make_sdesc_sw128(o.addr + k * 2048, o.lbo, 1024);
  return make_sdesc_sw128(o.addr + (k >> 2) * o.kblk + (k & 3) * 32, 16, 1024);
}
__device__ __forceinline__ void fa_issue(uint32_t tmem_d, const FOp& A, const FOp& B, int N, int ksteps, bool acc) {
  const uint32_t idesc = make_idesc_bf16(128, N, A.mn, B.mn);
  for (int k = 0; k < ksteps; ++k) umma_bf16(tmem_d, fop_desc(A, k), fop_desc(B, k), idesc, (acc || k > 0) ? 1u : 0u);
}
// inverse rotation of a gradient row (the transpose of rope.py:60-98)
template <int D>
__device__ __forceinline__ void fa_rope_inv(float (&x)[D], const float* __restrict__ cosT, const float* __restrict__ sinT, int t) {
#pragma unroll
  for (int c = 0; c < D / 2; ++c) {
    const float cs = cosT[(size_t)t * (D / 2) + c], sn = -sinT[(size_t)t * (D / 2) + c];
    const float lo = x[c], hi = x[c + D / 2];
    x[c] = lo * cs - hi * sn;
    x[c + D / 2] = hi * cs + lo * sn;
  }
}

struct FlashBwdParams {
  const bf16* ctx; const float* lse; bf16* dqkv; float* dq_part;
  const float* cosT; const float* sinT;
  int B, T, heads, H;
  float scale, p_drop; const uint64_t* rng; uint32_t site;
};

template <int D>
__global__ void __launch_bounds__(FA_THREADS, 1)
attn_flash_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO, const FlashBwdParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t *sQ = base + FB_Q, *sDO = base + FB_DO, *sK = base + FB_K, *sV = base + FB_V, *sDS = base + FB_DS, *sPT = base + FB_PT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + FB_BAR);
  uint64_t *b_kv = bars, *b_q = bars + 1 /* [2] */, *b_s = bars + 3, *b_g = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int r = ((warp & 3) << 5) | (tid & 31);
  const int cg = warp >> 2;
  const int kb = blockIdx.x, h = blockIdx.y, b = blockIdx.z, T = P.T;
  const int row0 = b * T, key0 = kb * 128, Tk = T - key0 < 128 ? T - key0 : 128;
  const int nq = (T + 127) >> 7;
  const size_t M = (size_t)P.B * T;
  const int colQ = h * D, colK = P.H + h * D, colV = 2 * P.H + h * D;
  const int cq = (colQ & 63) >> 3, ck = (colK & 63) >> 3, cv = (colV & 63) >> 3;

  if (tid == 0) {
    tma_prefetch_desc(&tmQKV); tma_prefetch_desc(&tmDO);
    for (int k = 0; k < 5; ++k) mbar_init(bars + k, 1);
    fence_barrier_init();
  }
  pdl_wait();     // q/k/v, ctx, lse and dctx come from earlier kernels of the step
  pdl_trigger();
  auto load_q = [&](int qt) {
    const int s = qt & 1;
    mbar_expect_tx(b_q + s, 32768);
    tma_load_2d(sQ + s * 16384, &tmQKV, b_q + s, colQ & ~63, row0 + qt * 128);
    tma_load_2d(sDO + s * 16384, &tmDO, b_q + s, colQ & ~63, row0 + qt * 128);
  };
  if (tid == 0) {
    mbar_expect_tx(b_kv, 32768);
    tma_load_2d(sK, &tmQKV, b_kv, colK & ~63, row0 + key0);
    tma_load_2d(sV, &tmQKV, b_kv, colV & ~63, row0 + key0);
    load_q(0);
    if (nq > 1) load_q(1);
  }
  if (warp == 0) tmem_alloc(tmem_slot, FBC_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t my_tmem = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const FOp Kk{smem_u32(sK) + ck * 16, 16, 16384, 0}, Kmn{smem_u32(sK) + ck * 16, 16384, 0, 1};
  const FOp Vk{smem_u32(sV) + cv * 16, 16, 16384, 0};
  const FOp DSk{smem_u32(sDS), 16, 16384, 0}, DSmn{smem_u32(sDS), 16384, 0, 1}, PTmn{smem_u32(sPT), 16384, 0, 1};
  auto issue_s = [&](int qt) {   // thread 0: S = Q K^T, dP = dO V^T of query tile qt
    const int s = qt & 1;
    const FOp Qk{smem_u32(sQ) + s * 16384 + cq * 16, 16, 16384, 0}, DOk{smem_u32(sDO) + s * 16384 + cq * 16, 16, 16384, 0};
    fa_issue(tmem + FBC_S, Qk, Kk, 128, D / 16, false);
    fa_issue(tmem + FBC_DP, DOk, Vk, 128, D / 16, false);
    umma_commit(b_s);
  };
  const DropCtx dc = make_drop(P.p_drop, P.rng ? P.rng[0] : 0ull, P.rng ? (uint32_t)P.rng[1] : 0u, P.site);
  const int Tpad = attn_drop_tpad(T);
  const float sl2 = P.scale * FA_LOG2E;

  mbar_wait(b_kv, 0);
  mbar_wait(b_q, 0);
  if (P.cosT) {
    if (cg == 0) fa_rope_row<D>(sQ, r, cq, P.cosT, P.sinT, r < T ? r : 0);
    if (cg == 1) fa_rope_row<D>(sK, r, ck, P.cosT, P.sinT, r < Tk ? key0 + r : 0);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) { tc_fence_after(); issue_s(0); }

  for (int qt = 0; qt < nq; ++qt) {
    const int s = qt & 1;
    const int i = qt * 128 + r;
    const bool valid = i < T;
    const int ic = valid ? i : T - 1;
    mbar_wait(b_q + s, (uint32_t)((qt >> 1) & 1));   // (observed by every thread: the dO rows are read below)
    // row statistics (issued before the wait for the products): lse_i and D_i = dO_i . O_i
    const float lse2 = P.lse[(size_t)(b * P.heads + h) * T + ic] * FA_LOG2E;
    float Di = 0.f;
    {
      const bf16* op = P.ctx + (size_t)(row0 + ic) * P.H + h * D;
#pragma unroll
      for (int c = 0; c < D / 8; ++c) {
        float o8[8], d8[8];
        fa_unpack8(*reinterpret_cast<const uint4*>(op + c * 8), o8);
        fa_unpack8(*reinterpret_cast<const uint4*>(fa_swz(sDO + s * 16384, r, cq + c)), d8);
#pragma unroll
        for (int q = 0; q < 8; ++q) Di = fmaf(d8[q], o8[q], Di);
      }
    }
    const uint64_t drow = ((uint64_t)(b * P.heads + h) * T + ic) * (uint64_t)Tpad;
    mbar_wait(b_s, (uint32_t)(qt & 1));
    tc_fence_after();
#pragma unroll
    for (int k2 = 0; k2 < 2; ++k2) {
      const int c0 = (cg + k2 * FA_CG) * 16;
      float sv[16], dp[16];
      tmem_ld_32x16(my_tmem + FBC_S + c0, sv);
      tmem_ld_32x16(my_tmem + FBC_DP + c0, dp);
      if (valid && c0 + 16 <= Tk) {   // full chunk of a live query row: no masking
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
          float kp[8];
          drop8(dc, (drow + (uint64_t)(key0 + c0 + j)) >> 3, kp);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float p = ex2_approx(fmaf(sv[j + q], sl2, -lse2));
            const float pk = p * kp[q];                       // dropped probability
            sv[j + q] = fmaf(dp[j + q], pk, -p * Di);         // dS / scale (the scale is applied to dQ, dK)
            dp[j + q] = pk;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
          float kp[8];
          if (c0 + j < Tk) drop8(dc, (drow + (uint64_t)(key0 + c0 + j)) >> 3, kp);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const bool on = valid && (c0 + j + q < Tk);
            const float kq = (c0 + j < Tk) ? kp[q] : 0.f;
            const float p = on ? ex2_approx(fmaf(sv[j + q], sl2, -lse2)) : 0.f;
            const float pk = p * kq;
            sv[j + q] = on ? fmaf(dp[j + q], pk, -p * Di) : 0.f;   // (rows / keys outside the sample hold foreign data)
            dp[j + q] = pk;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 16; j += 8) {
        *reinterpret_cast<uint4*>(fa_swz(sDS, r, (c0 + j) >> 3)) = fa_pack8(&sv[j]);
        *reinterpret_cast<uint4*>(fa_swz(sPT, r, (c0 + j) >> 3)) = fa_pack8(&dp[j]);
      }
    }
    if (P.cosT && qt + 1 < nq) {   // RoPE: the next Q tile is rotated in place before its scores are issued
      mbar_wait(b_q + (s ^ 1), (uint32_t)(((qt + 1) >> 1) & 1));
      if (cg == 0) fa_rope_row<D>(sQ + (s ^ 1) * 16384, r, cq, P.cosT, P.sinT, i + 128 < T ? i + 128 : 0);
    }
    fence_proxy_async();
    tc_fence_before();
    fa_bar();
    if (tid == 0) {
      tc_fence_after();
      const FOp Qmn{smem_u32(sQ) + s * 16384 + cq * 16, 16384, 0, 1}, DOmn{smem_u32(sDO) + s * 16384 + cq * 16, 16384, 0, 1};
      fa_issue(tmem + FBC_DQ, DSk, Kmn, D, 8, false);        // dQ[i,:]  = sum_j dS[i,j] k_j
      fa_issue(tmem + FBC_DK, DSmn, Qmn, D, 8, qt > 0);      // dK[j,:] += sum_i dS[i,j] q_i
      fa_issue(tmem + FBC_DV, PTmn, DOmn, D, 8, qt > 0);     // dV[j,:] += sum_i P~[i,j] dO_i
      umma_commit(b_g);
      if (qt + 1 < nq) {
        mbar_wait(b_q + (s ^ 1), (uint32_t)(((qt + 1) >> 1) & 1));
        issue_s(qt + 1);                                      // runs behind the products above, under the dQ epilogue
      }
    }
    mbar_wait(b_g, (uint32_t)(qt & 1));
    tc_fence_after();
    if (tid == 0 && qt + 2 < nq) load_q(qt + 2);              // the Q / dO stage of this tile is free
    {
      float* dst = P.dq_part + ((size_t)kb * M + row0 + i) * P.H + h * D;
      if (D <= 32) {
        if (cg < D / 8) {
          float o[8];
          tmem_ld_32x8(my_tmem + FBC_DQ + cg * 8, o);
          if (valid) {
            float4* d4 = reinterpret_cast<float4*>(dst + cg * 8);
            d4[0] = make_float4(o[0] * P.scale, o[1] * P.scale, o[2] * P.scale, o[3] * P.scale);
            d4[1] = make_float4(o[4] * P.scale, o[5] * P.scale, o[6] * P.scale, o[7] * P.scale);
          }
        }
      } else {
        float o[16];
        tmem_ld_32x16(my_tmem + FBC_DQ + cg * 16, o);
        if (valid) {
          float4* d4 = reinterpret_cast<float4*>(dst + cg * 16);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            d4[c] = make_float4(o[4 * c] * P.scale, o[4 * c + 1] * P.scale, o[4 * c + 2] * P.scale, o[4 * c + 3] * P.scale);
        }
      }
    }
    // (no barrier here: thread 0 overwrites dQ / the dS tiles only behind the barrier in the middle of the next
    //  iteration, which every thread reaches after this epilogue)
  }
  // ---- dK, dV of this key block: thread = key row; the 2 D/8 eight-column pieces of [dK | dV] are spread over the groups ----
  tc_fence_after();
  {
    const int j = r;
    if (P.cosT) {
      if (cg < 2) {   // cg 0: dK row (inverse rotation needs the whole row), cg 1: dV row
        float x[D];
#pragma unroll
        for (int c = 0; c < D; c += 8) {
          float o[8];
          tmem_ld_32x8(my_tmem + (cg == 0 ? FBC_DK : FBC_DV) + c, o);
#pragma unroll
          for (int q = 0; q < 8; ++q) x[c + q] = o[q];
        }
        if (j < Tk) {
          if (cg == 0) {
#pragma unroll
            for (int c = 0; c < D; ++c) x[c] = bf16_round(x[c] * P.scale);
            fa_rope_inv<D>(x, P.cosT, P.sinT, key0 + j);
          }
          bf16* dst = P.dqkv + (size_t)(row0 + key0 + j) * (3 * P.H) + (cg == 0 ? P.H : 2 * P.H) + h * D;
#pragma unroll
          for (int c = 0; c < D; c += 8) *reinterpret_cast<uint4*>(dst + c) = fa_pack8(&x[c]);
        }
      }
    } else {
#pragma unroll 1
      for (int pc = cg; pc < 2 * (D / 8); pc += FA_CG) {
        const bool is_v = pc >= D / 8;
        const int c8 = (is_v ? pc - D / 8 : pc) * 8;
        float o[8];
        tmem_ld_32x8(my_tmem + (is_v ? FBC_DV : FBC_DK) + c8, o);
        if (j < Tk) {
          if (!is_v) {
#pragma unroll
            for (int c = 0; c < 8; ++c) o[c] *= P.scale;
          }
          *reinterpret_cast<uint4*>(P.dqkv + (size_t)(row0 + key0 + j) * (3 * P.H) + (is_v ? 2 * P.H : P.H) + h * D + c8) = fa_pack8(o);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, FBC_COLS);
}

// dQ = sum over key blocks (block order) of the fp32 partials -> bf16 (-> inverse RoPE).  One thread per (row, head).
template <int D>
__global__ void __launch_bounds__(256)
attn_flash_dq_kernel(const float* __restrict__ dq_part, bf16* __restrict__ dqkv, const float* __restrict__ cosT,
                     const float* __restrict__ sinT, int B, int T, int heads, int nkb) {
  pdl_wait();
  pdl_trigger();
  const int H = heads * D;
  const size_t M = (size_t)B * T;
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= M * heads) return;
  const size_t row = idx / heads;
  const int h = (int)(idx - row * heads);
  float x[D];
#pragma unroll
  for (int c = 0; c < D; ++c) x[c] = 0.f;
  for (int k = 0; k < nkb; ++k) {
    const float4* src = reinterpret_cast<const float4*>(dq_part + ((size_t)k * M + row) * H + h * D);
#pragma unroll
    for (int c = 0; c < D / 4; ++c) {
      const float4 a = __ldcs(src + c);
      x[4 * c] += a.x; x[4 * c + 1] += a.y; x[4 * c + 2] += a.z; x[4 * c + 3] += a.w;
    }
  }
  if (cosT) {
#pragma unroll
    for (int c = 0; c < D; ++c) x[c] = bf16_round(x[c]);
    fa_rope_inv<D>(x, cosT, sinT, (int)(row % T));
  }
  bf16* dst = dqkv + row * (3 * H) + h * D;
#pragma unroll
  for (int c = 0; c < D; c += 8) *reinterpret_cast<uint4*>(dst + c) = fa_pack8(&x[c]);
}

}  // namespace vb

using namespace vb;

// q, k, v = the three column blocks of one fused [B*T, 3H] bf16 buffer
extern "C" int vitb200_attn_flash_supported(int T, int d, int ld, int H) {
  if (!(d == 16 || d == 32 || d == 64)) return 0;
  if (ld != 3 * H || (H % 8) != 0) return 0;
  return T > 0 ? 1 : 0;
}

extern "C" int vitb200_attn_flash_fwd(const void* qkv, void* ctx, float* lse, const float* rope_cos, const float* rope_sin,
                                      int B, int T, int heads, int d, float scale, float p_drop, const uint64_t* rng,
                                      uint32_t site, void* stream) {
  if (!qkv || !ctx || !lse || B <= 0 || T <= 0 || heads <= 0) return VITB200_ERR_ARG;
  const int H = heads * d, ld = 3 * H;
  if (!vitb200_attn_flash_supported(T, d, ld, H)) return VITB200_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(ctx) & 15) != 0) return VITB200_ERR_ALIGN;
  CUtensorMap tm;
  int rc;
  if ((rc = get_tmap(qkv, ld, (uint64_t)B * T, 64, 128, &tm))) return rc;
  FlashParams P{(bf16*)ctx, lse, rope_cos, rope_sin, B, T, heads, H, scale, p_drop, rng, site};
  dim3 grid((T + 127) / 128, heads, B);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_FA(DD)                                                                                                  \
  {                                                                                                                    \
    static bool done = false;                                                                                          \
    if (!done) {                                                                                                       \
      cudaError_t e = cudaFuncSetAttribute(attn_flash_fwd_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM); \
      if (e != cudaSuccess) return vb_cuda_error(e);                                                                   \
      done = true;                                                                                                     \
    }                                                                                                                  \
    vb_launch_pdl(attn_flash_fwd_kernel<DD>, grid, dim3(FA_THREADS), FA_SMEM, st, tm, P);                              \
  }
  if (d == 16) LAUNCH_FA(16) else if (d == 32) LAUNCH_FA(32) else LAUNCH_FA(64)
#undef LAUNCH_FA
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

extern "C" size_t vitb200_attn_flash_bwd_ws_bytes(int B, int T, int heads, int d) {
  const size_t nkb = (size_t)(T + 127) / 128;
  return nkb * (size_t)B * T * heads * d * sizeof(float) + 256;
}

// dqkv [B*T, 3H] bf16 receives dq | dk | dv; ws: vitb200_attn_flash_bwd_ws_bytes() bytes (fp32 dQ partials per key block)
extern "C" int vitb200_attn_flash_bwd(const void* qkv, const void* ctx, const void* dctx, const float* lse, void* dqkv,
                                      const float* rope_cos, const float* rope_sin, int B, int T, int heads, int d,
                                      float scale, float p_drop, const uint64_t* rng, uint32_t site, void* ws, void* stream) {
  if (!qkv || !ctx || !dctx || !lse || !dqkv || !ws || B <= 0 || T <= 0 || heads <= 0) return VITB200_ERR_ARG;
  const int H = heads * d, ld = 3 * H;
  if (!vitb200_attn_flash_supported(T, d, ld, H)) return VITB200_ERR_SHAPE;
  if (((reinterpret_cast<uintptr_t>(ws) | reinterpret_cast<uintptr_t>(dqkv) | reinterpret_cast<uintptr_t>(ctx)) & 15) != 0)
    return VITB200_ERR_ALIGN;
  const int nkb = (T + 127) / 128;
  CUtensorMap tm, tmDO;
  int rc;
  if ((rc = get_tmap(qkv, ld, (uint64_t)B * T, 64, 128, &tm))) return rc;
  if ((rc = get_tmap(dctx, H, (uint64_t)B * T, 64, 128, &tmDO))) return rc;
  FlashBwdParams P{(const bf16*)ctx, lse, (bf16*)dqkv, (float*)ws, rope_cos, rope_sin, B, T, heads, H, scale, p_drop, rng, site};
  dim3 grid(nkb, heads, B);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t items = (size_t)B * T * heads;
  const unsigned rgrid = (unsigned)((items + 255) / 256);
#define LAUNCH_FB(DD)                                                                                                  \
  {                                                                                                                    \
    static bool done = false;                                                                                          \
    if (!done) {                                                                                                       \
      cudaError_t e = cudaFuncSetAttribute(attn_flash_bwd_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM); \
      if (e != cudaSuccess) return vb_cuda_error(e);                                                                   \
      done = true;                                                                                                     \
    }                                                                                                                  \
    vb_launch_pdl(attn_flash_bwd_kernel<DD>, grid, dim3(FA_THREADS), FB_SMEM, st, tm, tmDO, P);                        \
    VB_CHECK_LAUNCH();                                                                                                 \
    vb_launch_pdl(attn_flash_dq_kernel<DD>, dim3(rgrid), dim3(256), 0, st, (const float*)ws, (bf16*)dqkv, rope_cos,    \
                  rope_sin, B, T, heads, nkb);                                                                         \
  }
  if (d == 16) LAUNCH_FB(16) else if (d == 32) LAUNCH_FB(32) else LAUNCH_FB(64)
#undef LAUNCH_FB
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}
