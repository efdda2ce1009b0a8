// Multi-head self-attention on tcgen05 for sequences whose keys fit one TMEM tile (T <= 176), bf16.
//
// One CTA per (sample, head): 16 MMA-path warps (512 threads) + 1 side-row warp.  Q/K/V tiles are TMA-loaded straight
// out of the fused [B*T, 3H] QKV buffer (box = 64 columns starting at the head's column, so no head-major copy
// exists); only the first d columns of each 128-byte row take part in the MMAs.
// Warp w of the MMA path owns TMEM lane quarter (w & 3) (query rows 32 (w & 3) .. +31 of the tile) and key-column
// group cg = w >> 2: the softmax row of a query is split over four threads in chunks of 16 keys (chunk c belongs to
// group c & 3); row max / row sum are merged through shared memory.  (One thread per query row -- the first version --
// left each SM scheduler a single warp with ~6k dependent instructions: 12 us per launch at 128 CTAs.)
//   forward : S = Q K^T (UMMA, fp32 in TMEM) -> softmax in registers (exp2 with the scale folded in; Philox dropout)
//             -> P (bf16) to swizzled smem -> O = P V (UMMA, V as MN-major B)
//   backward: S and dP = dO V^T by UMMA; P, dS in registers; dS and dropped P go to smem ONCE and
//             are used as K-major A (dQ = dS K) and as MN-major A (dK = dS^T Q, dV = P^T dO), the latter
//             accumulating in TMEM over the query tiles.  No atomics: results are bitwise reproducible.
// Query tiles of 128 rows are looped inside the CTA; K/V are staged once.
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_host.cuh"

namespace vb {
using namespace vb::tc;

constexpr int AT_CG = 4;                          // key-column groups = threads per query row
constexpr int AT_TC_THREADS = 128 * AT_CG;        // MMA-path threads (512)
constexpr int AT_ALL_THREADS = AT_TC_THREADS + 32;  // + the side-row warp
constexpr float AT_LOG2E = 1.4426950408889634f;
VB_TL_DECL(tl_attn_fwd)

__device__ __forceinline__ uint4 at_pack8(const float* v) {
  __nv_bfloat162 t0 = __floats2bfloat162_rn(v[0], v[1]), t1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 t2 = __floats2bfloat162_rn(v[4], v[5]), t3 = __floats2bfloat162_rn(v[6], v[7]);
  uint4 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
  pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
  return pk;
}
__device__ __forceinline__ void at_unpack8(uint4 pk, float* v) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
  for (int q = 0; q < 4; ++q) { float2 f = __bfloat1622float2(p[q]); v[2 * q] = f.x; v[2 * q + 1] = f.y; }
}
__device__ __forceinline__ uint8_t* at_swz(uint8_t* tile, int r, int chunk) {
  return tile + (chunk >> 3) * 16384 + r * 128 + (((chunk & 7) ^ (r & 7)) << 4);
}
// row r (of a tile whose rows are 128 B) holds d values in chunks 0..d/8-1: load / store them
// (c0 = first 16-byte chunk of the head inside the 64-column TMA box: the box starts at a multiple of 64 columns)
template <int D>
__device__ __forceinline__ void at_load_row(uint8_t* tile, int r, float (&x)[D], int c0 = 0) {
#pragma unroll
  for (int c = 0; c < D / 8; ++c) at_unpack8(*reinterpret_cast<const uint4*>(at_swz(tile, r, c0 + c)), &x[c * 8]);
}
template <int D>
__device__ __forceinline__ void at_store_row(uint8_t* tile, int r, const float (&x)[D], int c0 = 0) {
#pragma unroll
  for (int c = 0; c < D / 8; ++c) *reinterpret_cast<uint4*>(at_swz(tile, r, c0 + c)) = at_pack8(&x[c * 8]);
}
// RoPE on a full head row held by one thread (rope.py:60-98); INV = transpose (for gradients)
template <int D, bool INV>
__device__ __forceinline__ void at_rope(float (&x)[D], const float* __restrict__ cosT, const float* __restrict__ sinT, int t) {
#pragma unroll
  for (int c = 0; c < D / 2; ++c) {
    const float cs = cosT[(size_t)t * (D / 2) + c];
    float sn = sinT[(size_t)t * (D / 2) + c];
    if (INV) sn = -sn;
    const float lo = x[c], hi = x[c + D / 2];
    x[c] = lo * cs - hi * sn;
    x[c + D / 2] = hi * cs + lo * sn;
  }
}

struct AOp { uint32_t addr, lbo, kblk; int mn; };
__device__ __forceinline__ uint64_t aop_desc(const AOp& o, int k) {
  if (o.mn) return make_sdesc_sw128(o.addr + k * 2048, o.lbo, 1024);
  return make_sdesc_sw128(o.addr + (k >> 2) * o.kblk + (k & 3) * 32, 16, 1024);
}
__device__ __forceinline__ void at_issue(uint32_t tmem_d, const AOp& A, const AOp& B, int N, int ksteps, bool acc) {
  const uint32_t idesc = make_idesc_bf16(128, N, A.mn, B.mn);
  for (int k = 0; k < ksteps; ++k) umma_bf16(tmem_d, aop_desc(A, k), aop_desc(B, k), idesc, (acc || k > 0) ? 1u : 0u);
}

__device__ __forceinline__ void bar_main() { asm volatile("bar.sync 1, 512;" ::: "memory"); }    // the 16 MMA-path warps
__device__ __forceinline__ void bar_all() { asm volatile("bar.sync 2, 544;" ::: "memory"); }     // + the side-row warp
__device__ __forceinline__ float warp_allsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_allmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
constexpr int AT_SIDE_KEYS = 6;  // keys per lane of the side-row warp: covers T <= 192
constexpr int AT_MAXCH = 3;      // 16-key chunks per thread: covers KP <= 192

struct AttnTcParams {
  const bf16* qkv;       // q pointer (k = q + H, v = q + 2H inside the same rows)
  bf16* ctx; float* lse;
  const bf16* dctx; bf16* dqkv; int ld_d;
  const float* cosT; const float* sinT;
  int B, T, heads, H, ld, KP;   // KP = keys (of this launch) padded to a multiple of 16
  float scale, p_drop; const uint64_t* rng; uint32_t site;
  // key-blocked mode (long sequences): this launch handles keys [key0, key0 + Tk) of every sample only.
  //   forward : not available (long sequences run attention_flash.cu)
  //   backward: dK / dV of the block's keys are complete; dQ is accumulated over the launches in an fp32 buffer
  //             (first: store, otherwise add; last: the bf16 dq rows are written)
  int kb, key0, Tk, first, last;
  float* dq_acc;
};

// ================================================================================================
// forward
// ================================================================================================
template <int D>
__global__ void __launch_bounds__(AT_ALL_THREADS, 1)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AttnTcParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const int KP = P.KP, nch = KP / 16;
  const uint32_t kv_bytes = (uint32_t)KP * 128;
  uint8_t* sQ = base;                       // 16 KB
  uint8_t* sK = sQ + 16384;                 // 32 KB reserved
  uint8_t* sV = sK + 32768;                 // 32 KB reserved
  uint8_t* sP = sV + 32768;                 // up to 3 x 16 KB (4 reserved)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * 16384);
  uint64_t *b_kv = bars, *b_q = bars + 1, *b_mma = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  float* s_mx = reinterpret_cast<float*>(bars + 8);   // [AT_CG][128] row-max exchange
  float* s_sm = s_mx + AT_CG * 128;                   // [AT_CG][128] row-sum exchange
  constexpr uint32_t TMEM_COLS = 512;
  const uint32_t cS = 0, cO = 256;

  const int tid = threadIdx.x, warp = tid >> 5;
  VB_TL(tl_attn_fwd, 0);
  const int r = ((warp & 3) << 5) | (tid & 31);   // query row of the tile = TMEM lane
  const int cg = warp >> 2;                       // key-column group (4 = side-row warp)
  const int h = blockIdx.x, b = blockIdx.y, T = P.T;
  const int row0 = b * T;
  const int key0 = P.kb ? P.key0 : 0, Tk = P.kb ? P.Tk : T;   // keys of this launch: [key0, key0 + Tk)
  // TMA boxes are 64 columns wide and start at multiples of 64 columns; a head's d columns sit at 16-byte chunk
  // cq / ck / cv inside the 128-byte rows (UMMA descriptors and row accessors start there: the 128B swizzle is a
  // function of the address bits, so a start offset inside the swizzle atom selects the columns)
  const int colQ = h * D, colK = P.H + h * D, colV = 2 * P.H + h * D;
  const int cq = (colQ & 63) >> 3, ck = (colK & 63) >> 3, cv = (colV & 63) >> 3;
  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmKV);
    mbar_init(b_kv, 1); mbar_init(b_q, 1); mbar_init(b_mma, 1);
    fence_barrier_init();
  }
  VB_TL(tl_attn_fwd, 1);
  pdl_wait();     // q/k/v come from the previous kernel of the step
  pdl_trigger();
  VB_TL(tl_attn_fwd, 2);
  if (tid == 0) {    // loads first (same thread that initialised the barriers): they fly while TMEM is allocated
    mbar_expect_tx(b_kv, 2 * kv_bytes);
    tma_load_2d(sK, &tmKV, b_kv, colK & ~63, row0 + key0);
    tma_load_2d(sV, &tmKV, b_kv, colV & ~63, row0 + key0);
    mbar_expect_tx(b_q, 16384);                       // first query tile
    tma_load_2d(sQ, &tmQ, b_q, colQ & ~63, row0);
  }
  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t my_tmem = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  VB_TL(tl_attn_fwd, 3);
  const AOp Qk{smem_u32(sQ) + cq * 16, 16, 16384, 0}, Kk{smem_u32(sK) + ck * 16, 16, 16384, 0};
  const AOp Pk{smem_u32(sP), 16, 16384, 0}, Vmn{smem_u32(sV) + cv * 16, 16384, 0, 1};
  const DropCtx dc = make_drop(P.p_drop, P.rng ? P.rng[0] : 0ull, P.rng ? (uint32_t)P.rng[1] : 0u, P.site);
  const int Tpad = attn_drop_tpad(T);
  const float sl2 = P.scale * AT_LOG2E;
  uint32_t ph_q = 0, ph_mma = 0;
  // T = 128 n + 1 (the CLS token makes every configured sequence one row longer than a tile): the last query row
  // is handled by a 17th warp with plain FMAs, concurrently with the tensor-core tiles, instead of a whole extra tile.
  const bool side = (T % 128 == 1) && T > 1 && P.cosT == nullptr;
  const int nq = side ? T / 128 : (T + 127) / 128;

  if (cg == AT_CG) {
    if (side) {
      const int lane = tid & 31, i = T - 1;
      mbar_wait(b_kv, 0);
      float qf[D];
      {
        const bf16* qp = P.qkv + (size_t)(row0 + i) * P.ld + h * D;
#pragma unroll
        for (int c = 0; c < D; c += 8) at_unpack8(*reinterpret_cast<const uint4*>(qp + c), &qf[c]);
      }
      float sc[AT_SIDE_KEYS];
      float mx = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < AT_SIDE_KEYS; ++jj) {
        const int j = lane + 32 * jj;
        float acc = -INFINITY;
        if (j < Tk) {
          float kr[D];
          at_load_row<D>(sK, j, kr, ck);
          acc = 0.f;
#pragma unroll
          for (int c = 0; c < D; ++c) acc = fmaf(qf[c], kr[c], acc);
        }
        sc[jj] = acc;
        mx = fmaxf(mx, acc);
      }
      mx = warp_allmax(mx);
      const uint64_t drow = ((uint64_t)(b * P.heads + h) * T + i) * (uint64_t)Tpad;
      float sum = 0.f, o[D];
#pragma unroll
      for (int c = 0; c < D; ++c) o[c] = 0.f;
#pragma unroll
      for (int jj = 0; jj < AT_SIDE_KEYS; ++jj) {
        const int j = lane + 32 * jj;
        if (j < Tk) {
          float p = exp2f((sc[jj] - mx) * sl2);
          sum += p;
          p *= drop1(dc, drow + (uint64_t)(key0 + j));
          p = bf16_round(p);
          float vr[D];
          at_load_row<D>(sV, j, vr, cv);
#pragma unroll
          for (int c = 0; c < D; ++c) o[c] = fmaf(p, vr[c], o[c]);
        }
      }
      sum = warp_allsum(sum);
#pragma unroll
      for (int c = 0; c < D; ++c) o[c] = warp_allsum(o[c]);
      if (lane == 0) {
        const float inv = 1.f / sum;
#pragma unroll
        for (int c = 0; c < D; ++c) o[c] *= inv;
        bf16* dst = P.ctx + (size_t)(row0 + i) * P.H + h * D;
#pragma unroll
        for (int c = 0; c < D; c += 8) *reinterpret_cast<uint4*>(dst + c) = at_pack8(&o[c]);
        P.lse[(size_t)(b * P.heads + h) * T + i] = mx * P.scale + logf(sum);
      }
    }
  } else {
  for (int qt = 0; qt < nq; ++qt) {
    const int q0 = qt * 128, i = q0 + r;
    const bool valid = i < T;
    if (tid == 0 && qt > 0) {
      mbar_expect_tx(b_q, 16384);
      tma_load_2d(sQ, &tmQ, b_q, colQ & ~63, row0 + q0);
    }
    if (qt == 0) mbar_wait(b_kv, 0);
    mbar_wait(b_q, ph_q); ph_q ^= 1;
  VB_TL(tl_attn_fwd, 4);
    if (P.cosT) {  // rotate the query rows (and, once, the key rows) in place
      float x[D];
      if (cg == 0) {
        at_load_row<D>(sQ, r, x, cq);
        at_rope<D, false>(x, P.cosT, P.sinT, valid ? i : 0);
        at_store_row<D>(sQ, r, x, cq);
      }
      if (qt == 0) {
        for (int rr = tid; rr < KP; rr += AT_TC_THREADS) {
          at_load_row<D>(sK, rr, x, ck);
          at_rope<D, false>(x, P.cosT, P.sinT, rr < Tk ? key0 + rr : 0);
          at_store_row<D>(sK, rr, x, ck);
        }
      }
      fence_proxy_async();
    }
    tc_fence_before();
    bar_main();
    if (tid == 0) {
      tc_fence_after();
      at_issue(tmem + cS, Qk, Kk, KP, D / 16, false);   // S[i, j] = q_i . k_j
      umma_commit(b_mma);
    }
    mbar_wait(b_mma, ph_mma); ph_mma ^= 1;
  VB_TL(tl_attn_fwd, 5);
    tc_fence_after();
    // ---- softmax: this thread holds the 16-key chunks cg, cg + 4, cg + 8 of its query row ----
    float v[AT_MAXCH][16];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < AT_MAXCH; ++k) {
      const int c0 = (cg + k * AT_CG) * 16;
      if (cg + k * AT_CG < nch) {
        tmem_ld_32x16(my_tmem + cS + c0, v[k]);
        if (c0 + 16 <= Tk) {   // full chunk: no key masking
#pragma unroll
          for (int j = 0; j < 16; ++j) mx = fmaxf(mx, v[k][j]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) if (c0 + j < Tk) mx = fmaxf(mx, v[k][j]);
        }
      }
    }
    s_mx[cg * 128 + r] = mx;
    bar_main();
  VB_TL(tl_attn_fwd, 6);
    mx = fmaxf(fmaxf(s_mx[r], s_mx[128 + r]), fmaxf(s_mx[256 + r], s_mx[384 + r]));
    const float nmxs = -mx * sl2;
    float sum = 0.f;
    const uint64_t drow = ((uint64_t)(b * P.heads + h) * T + (valid ? i : 0)) * (uint64_t)Tpad;
#pragma unroll
    for (int k = 0; k < AT_MAXCH; ++k) {
      const int c0 = (cg + k * AT_CG) * 16;
      if (cg + k * AT_CG < nch) {
        if (c0 + 16 <= Tk) {
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            float kp[8];
            drop8(dc, (drow + (uint64_t)(key0 + c0 + j)) >> 3, kp);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float p = ex2_approx(fmaf(v[k][j + q], sl2, nmxs));
              sum += p;
              v[k][j + q] = p * kp[q];
            }
            *reinterpret_cast<uint4*>(at_swz(sP, r, (c0 + j) >> 3)) = at_pack8(&v[k][j]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            float kp[8];
            if (c0 + j < Tk) drop8(dc, (drow + (uint64_t)(key0 + c0 + j)) >> 3, kp);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float p = (c0 + j + q < Tk) ? ex2_approx(fmaf(v[k][j + q], sl2, nmxs)) : 0.f;
              sum += p;
              v[k][j + q] = (c0 + j < Tk) ? p * kp[q] : 0.f;
            }
            *reinterpret_cast<uint4*>(at_swz(sP, r, (c0 + j) >> 3)) = at_pack8(&v[k][j]);
          }
        }
      }
    }
    s_sm[cg * 128 + r] = sum;
    fence_proxy_async();
    tc_fence_before();
    bar_main();
  VB_TL(tl_attn_fwd, 7);
    if (tid == 0) {
      tc_fence_after();
      at_issue(tmem + cO, Pk, Vmn, D, KP / 16, false);  // O[i, c] = sum_j P[i, j] v[j, c]
      umma_commit(b_mma);
    }
    sum = (s_sm[r] + s_sm[128 + r]) + (s_sm[256 + r] + s_sm[384 + r]);
    mbar_wait(b_mma, ph_mma); ph_mma ^= 1;
  VB_TL(tl_attn_fwd, 8);
    tc_fence_after();
    if (cg < D / 8) {  // 8 output columns per thread
      float o[8];
      tmem_ld_32x8(my_tmem + cO + cg * 8, o);
      if (valid) {
        const float inv = 1.f / sum;
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] *= inv;
        *reinterpret_cast<uint4*>(P.ctx + (size_t)(row0 + i) * P.H + h * D + cg * 8) = at_pack8(o);
        if (cg == 0) P.lse[(size_t)(b * P.heads + h) * T + i] = mx * P.scale + logf(sum);
      }
    }
    tc_fence_before();
    bar_main();
  }
  }
  VB_TL(tl_attn_fwd, 9);
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

// ================================================================================================
// backward
// ================================================================================================
template <int D>
__global__ void __launch_bounds__(AT_ALL_THREADS, 1)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   const __grid_constant__ CUtensorMap tmDO, const AttnTcParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const int KP = P.KP, nch = KP / 16;
  const uint32_t kv_bytes = (uint32_t)KP * 128;
  // nblk <= 3 blocks of dS / P~ are written.  The MN-major view of key tile 1 spans blocks 2 and 3; block 3 then aliases
  // the NEXT buffer (sPT block 0, resp. the first 16 KB of sK): finite bf16 data whose product rows (keys >= 192)
  // are never read.
  uint8_t* sQ = base;                        // 16 KB
  uint8_t* sDO = sQ + 16384;                 // 16 KB
  uint8_t* sDS = sDO + 16384;                // 3 x 16 KB
  uint8_t* sPT = sDS + 49152;                // 3 x 16 KB
  uint8_t* sK = sPT + 49152;                 // 32 KB reserved
  uint8_t* sV = sK + 32768;                  // 32 KB reserved
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 32768);
  uint64_t *b_kv = bars, *b_q = bars + 1, *b_mma = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  float* xds = reinterpret_cast<float*>(bars + 8);  // side row: dS[j] (scale folded), dropped P[j], q row, dO row
  float* xpt = xds + 256;
  float* xq = xpt + 256;
  float* xdo = xq + 32;
  constexpr uint32_t TMEM_COLS = 512;
  // TMEM columns: S and dP take round_up(KP, 32) columns each, then dQ and the dK / dV accumulators of up to two key tiles
  const uint32_t KPa = (uint32_t)((KP + 31) / 32 * 32);
  const uint32_t cS = 0, cDP = KPa, cDQ = 2 * KPa, cDK0 = cDQ + 32, cDK1 = cDK0 + 32, cDV0 = cDK1 + 32, cDV1 = cDV0 + 32;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int r = ((warp & 3) << 5) | (tid & 31);
  const int cg = warp >> 2;
  const int h = blockIdx.x, b = blockIdx.y, T = P.T;
  const int row0 = b * T;
  const int key0 = P.kb ? P.key0 : 0, Tk = P.kb ? P.Tk : T;   // keys of this launch: [key0, key0 + Tk)
  const int colQ = h * D, colK = P.H + h * D, colV = 2 * P.H + h * D;   // see the forward kernel
  const int cq = (colQ & 63) >> 3, ck = (colK & 63) >> 3, cv = (colV & 63) >> 3;
  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmKV); tma_prefetch_desc(&tmDO);
    mbar_init(b_kv, 1); mbar_init(b_q, 1); mbar_init(b_mma, 1);
    fence_barrier_init();
  }
  pdl_wait();     // q/k/v and dctx come from earlier kernels of the step
  pdl_trigger();
  if (tid == 0) {    // loads first (same thread that initialised the barriers): they fly while TMEM is allocated
    mbar_expect_tx(b_kv, 2 * kv_bytes);
    tma_load_2d(sK, &tmKV, b_kv, colK & ~63, row0 + key0);
    tma_load_2d(sV, &tmKV, b_kv, colV & ~63, row0 + key0);
    mbar_expect_tx(b_q, 32768);                       // first query / dO tiles
    tma_load_2d(sQ, &tmQ, b_q, colQ & ~63, row0);
    tma_load_2d(sDO, &tmDO, b_q, colQ & ~63, row0);
  }
  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t my_tmem = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const AOp Qk{smem_u32(sQ) + cq * 16, 16, 16384, 0}, Qmn{smem_u32(sQ) + cq * 16, 16384, 0, 1};
  const AOp DOk{smem_u32(sDO) + cq * 16, 16, 16384, 0}, DOmn{smem_u32(sDO) + cq * 16, 16384, 0, 1};
  const AOp Kk{smem_u32(sK) + ck * 16, 16, 16384, 0}, Kmn{smem_u32(sK) + ck * 16, 16384, 0, 1};
  const AOp Vk{smem_u32(sV) + cv * 16, 16, 16384, 0};
  const AOp DSk{smem_u32(sDS), 16, 16384, 0};
  const DropCtx dc = make_drop(P.p_drop, P.rng ? P.rng[0] : 0ull, P.rng ? (uint32_t)P.rng[1] : 0u, P.site);
  const int Tpad = attn_drop_tpad(T);
  const float sl2 = P.scale * AT_LOG2E;
  uint32_t ph_q = 0, ph_mma = 0;
  const bool side = (T % 128 == 1) && T > 1 && P.cosT == nullptr;  // see the forward kernel
  const int nq = side ? T / 128 : (T + 127) / 128;
  const int nkt = (KP + 127) / 128;  // key tiles of the dK / dV accumulators (1 or 2)

  if (cg == AT_CG) {
    if (side) {
      // last query row with plain FMAs: dq directly; its rank-1 contributions to dK / dV are handed to the key-row
      // epilogue through shared memory (xds, xpt, xq, xdo)
      const int lane = tid & 31, i = T - 1;
      mbar_wait(b_kv, 0);
      float qf[D], dof[D];
      float Di = 0.f;
      {
        const bf16* qp = P.qkv + (size_t)(row0 + i) * P.ld + h * D;
        const bf16* dp_ = P.dctx + (size_t)(row0 + i) * P.H + h * D;
        const bf16* op = P.ctx + (size_t)(row0 + i) * P.H + h * D;
#pragma unroll
        for (int c = 0; c < D; c += 8) {
          float o8[8];
          at_unpack8(*reinterpret_cast<const uint4*>(qp + c), &qf[c]);
          at_unpack8(*reinterpret_cast<const uint4*>(dp_ + c), &dof[c]);
          at_unpack8(*reinterpret_cast<const uint4*>(op + c), o8);
#pragma unroll
          for (int q = 0; q < 8; ++q) Di = fmaf(dof[c + q], o8[q], Di);
        }
      }
      const float lse2 = P.lse[(size_t)(b * P.heads + h) * T + i] * AT_LOG2E;
      const uint64_t drow = ((uint64_t)(b * P.heads + h) * T + i) * (uint64_t)Tpad;
      float dq[D];
#pragma unroll
      for (int c = 0; c < D; ++c) dq[c] = 0.f;
#pragma unroll
      for (int jj = 0; jj < AT_SIDE_KEYS; ++jj) {
        const int j = lane + 32 * jj;
        if (j < Tk) {
          float kr[D], vr[D];
          at_load_row<D>(sK, j, kr, ck);
          at_load_row<D>(sV, j, vr, cv);
          float sdot = 0.f, dp = 0.f;
#pragma unroll
          for (int c = 0; c < D; ++c) { sdot = fmaf(qf[c], kr[c], sdot); dp = fmaf(dof[c], vr[c], dp); }
          const float p = exp2f(sdot * sl2 - lse2);
          const float keep = drop1(dc, drow + (uint64_t)(key0 + j));
          const float ds = bf16_round(p * (dp * keep - Di) * P.scale);
          xds[j] = ds;
          xpt[j] = bf16_round(p * keep);
#pragma unroll
          for (int c = 0; c < D; ++c) dq[c] = fmaf(ds, kr[c], dq[c]);
        }
      }
#pragma unroll
      for (int c = 0; c < D; ++c) dq[c] = warp_allsum(dq[c]);
      if (lane == 0) {
        if (P.kb) {
          float* acc = P.dq_acc + (size_t)(row0 + i) * P.H + h * D;
#pragma unroll
          for (int c = 0; c < D; ++c) { if (!P.first) dq[c] += acc[c]; if (!P.last) acc[c] = dq[c]; }
        }
        if (!P.kb || P.last) {
          bf16* dst = P.dqkv + (size_t)(row0 + i) * P.ld_d + h * D;
#pragma unroll
          for (int c = 0; c < D; c += 8) *reinterpret_cast<uint4*>(dst + c) = at_pack8(&dq[c]);
        }
#pragma unroll
        for (int c = 0; c < D; ++c) { xq[c] = qf[c]; xdo[c] = dof[c]; }
      }
    }
    bar_all();  // side-row results visible to the key-row epilogue
  } else {
  for (int qt = 0; qt < nq; ++qt) {
    const int q0 = qt * 128, i = q0 + r;
    const bool valid = i < T;
    const int ic = valid ? i : T - 1;
    if (tid == 0 && qt > 0) {
      mbar_expect_tx(b_q, 32768);
      tma_load_2d(sQ, &tmQ, b_q, colQ & ~63, row0 + q0);
      tma_load_2d(sDO, &tmDO, b_q, colQ & ~63, row0 + q0);
    }
    // row statistics: issued before the waits (lse_i, and the ctx row for D_i = dO_i . O_i)
    const float lse2 = P.lse[(size_t)(b * P.heads + h) * T + ic] * AT_LOG2E;
    uint4 oraw[D / 8];
    {
      const bf16* op = P.ctx + (size_t)(row0 + ic) * P.H + h * D;
#pragma unroll
      for (int c = 0; c < D / 8; ++c) oraw[c] = *reinterpret_cast<const uint4*>(op + c * 8);
    }
    if (qt == 0) mbar_wait(b_kv, 0);
    mbar_wait(b_q, ph_q); ph_q ^= 1;
    // rows of Q / dO beyond this sample must not leak into the dK / dV contractions: zero them
    float dof[D];
    if (!valid) {
#pragma unroll
      for (int c = 2 * cg; c < 2 * cg + 2; ++c) {
        *reinterpret_cast<uint4*>(at_swz(sQ, r, c)) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(at_swz(sDO, r, c)) = make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int c = 0; c < D; ++c) dof[c] = 0.f;
    } else {
      at_load_row<D>(sDO, r, dof, cq);
    }
    if (P.cosT) {
      float x[D];
      if (valid && cg == 0) {
        at_load_row<D>(sQ, r, x, cq);
        at_rope<D, false>(x, P.cosT, P.sinT, i);
        at_store_row<D>(sQ, r, x, cq);
      }
      if (qt == 0) {
        for (int rr = tid; rr < KP; rr += AT_TC_THREADS) {
          at_load_row<D>(sK, rr, x, ck);
          at_rope<D, false>(x, P.cosT, P.sinT, rr < Tk ? key0 + rr : 0);
          at_store_row<D>(sK, rr, x, ck);
        }
      }
    }
    fence_proxy_async();
    tc_fence_before();
    bar_main();
    if (tid == 0) {
      tc_fence_after();
      at_issue(tmem + cS, Qk, Kk, KP, D / 16, false);    // S  = Q K^T
      at_issue(tmem + cDP, DOk, Vk, KP, D / 16, false);  // dP = dO V^T
      umma_commit(b_mma);
    }
    // D_i = dO_i . O_i (flash-attention backward's row statistic); every thread of the row computes it
    float Di = 0.f;
#pragma unroll
    for (int c = 0; c < D / 8; ++c) {
      float o8[8];
      at_unpack8(oraw[c], o8);
#pragma unroll
      for (int q = 0; q < 8; ++q) Di = fmaf(dof[c * 8 + q], o8[q], Di);
    }
    const uint64_t drow = ((uint64_t)(b * P.heads + h) * T + ic) * (uint64_t)Tpad;
    mbar_wait(b_mma, ph_mma); ph_mma ^= 1;
    tc_fence_after();
#pragma unroll 1
    for (int ch = cg; ch < nch; ch += AT_CG) {
      const int c0 = ch * 16;
      float s[16], dp[16];
      tmem_ld_32x16(my_tmem + cS + c0, s);
      tmem_ld_32x16(my_tmem + cDP + c0, dp);
      if (valid && c0 + 16 <= Tk) {   // full chunk of a live query row: no masking
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
          float kpa[8];
          drop8(dc, (drow + (uint64_t)(key0 + c0 + j)) >> 3, kpa);
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
          if (c0 + j < Tk) drop8(dc, (drow + (uint64_t)(key0 + c0 + j)) >> 3, kpa);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const bool on = valid && (c0 + j + q < Tk);
            const float kq = (c0 + j < Tk) ? kpa[q] : 0.f;
            const float p = on ? ex2_approx(fmaf(s[j + q], sl2, -lse2)) : 0.f;
            const float pk = p * kq;
            s[j + q] = on ? fmaf(dp[j + q], pk, -p * Di) : 0.f;  // (columns past T hold stale TMEM data)
            dp[j + q] = pk;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 16; j += 8) {
        *reinterpret_cast<uint4*>(at_swz(sDS, r, (c0 + j) >> 3)) = at_pack8(&s[j]);
        *reinterpret_cast<uint4*>(at_swz(sPT, r, (c0 + j) >> 3)) = at_pack8(&dp[j]);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    bar_main();
    if (tid == 0) {
      tc_fence_after();
      at_issue(tmem + cDQ, DSk, Kmn, D, KP / 16, false);                       // dQ[i,:]  = sum_j dS[i,j] k_j
      for (int kt = 0; kt < nkt; ++kt) {
        const AOp DSmn{smem_u32(sDS) + kt * 32768, 16384, 0, 1}, PTmn{smem_u32(sPT) + kt * 32768, 16384, 0, 1};
        at_issue(tmem + (kt ? cDK1 : cDK0), DSmn, Qmn, D, 8, qt > 0);          // dK[j,:] += sum_i dS[i,j] q_i
        at_issue(tmem + (kt ? cDV1 : cDV0), PTmn, DOmn, D, 8, qt > 0);         // dV[j,:] += sum_i P~[i,j] dO_i
      }
      umma_commit(b_mma);
    }
    mbar_wait(b_mma, ph_mma); ph_mma ^= 1;
    tc_fence_after();
    if (P.cosT) {  // the inverse rotation pairs columns c and c + D/2: one thread takes the whole row
      if (cg == 0) {
        float o[32];
        tmem_ld_32x32(my_tmem + cDQ, o);
        if (valid) {
          float x[D];
#pragma unroll
          for (int c = 0; c < D; ++c) x[c] = o[c] * P.scale;
          if (P.kb) {
            float* acc = P.dq_acc + (size_t)(row0 + i) * P.H + h * D;
#pragma unroll
            for (int c = 0; c < D; ++c) { if (!P.first) x[c] += acc[c]; if (!P.last) acc[c] = x[c]; }
          }
          if (!P.kb || P.last) {
#pragma unroll
            for (int c = 0; c < D; ++c) x[c] = bf16_round(x[c]);
            at_rope<D, true>(x, P.cosT, P.sinT, i);
            bf16* dst = P.dqkv + (size_t)(row0 + i) * P.ld_d + h * D;
#pragma unroll
            for (int c = 0; c < D; c += 8) *reinterpret_cast<uint4*>(dst + c) = at_pack8(&x[c]);
          }
        }
      }
    } else if (cg < D / 8) {
      float o[8];
      tmem_ld_32x8(my_tmem + cDQ + cg * 8, o);
#pragma unroll
      for (int c = 0; c < 8; ++c) o[c] *= P.scale;
      if (valid && P.kb) {
        float4* acc = reinterpret_cast<float4*>(P.dq_acc + (size_t)(row0 + i) * P.H + h * D + cg * 8);
        if (!P.first) {
          const float4 a0 = acc[0], a1 = acc[1];
          o[0] += a0.x; o[1] += a0.y; o[2] += a0.z; o[3] += a0.w; o[4] += a1.x; o[5] += a1.y; o[6] += a1.z; o[7] += a1.w;
        }
        if (!P.last) { acc[0] = make_float4(o[0], o[1], o[2], o[3]); acc[1] = make_float4(o[4], o[5], o[6], o[7]); }
      }
      if (valid && (!P.kb || P.last))
        *reinterpret_cast<uint4*>(P.dqkv + (size_t)(row0 + i) * P.ld_d + h * D + cg * 8) = at_pack8(o);
    }
    tc_fence_before();
    bar_main();
  }
  // ---- dK, dV: thread = key row of key tile kt; the 2 D/8 eight-column pieces of [dK | dV] are spread over the groups ----
  bar_all();
  tc_fence_after();
  for (int kt = 0; kt < nkt; ++kt) {
    const int j = kt * 128 + r;
    if (P.cosT) {
      if (cg < 2) {  // cg 0: dK row (inverse rotation), cg 1: dV row
        float o[32];
        tmem_ld_32x32(my_tmem + (cg == 0 ? (kt ? cDK1 : cDK0) : (kt ? cDV1 : cDV0)), o);
        if (j < Tk) {
          float x[D];
#pragma unroll
          for (int c = 0; c < D; ++c) x[c] = o[c];
          if (cg == 0) {
#pragma unroll
            for (int c = 0; c < D; ++c) x[c] = bf16_round(x[c] * P.scale);
            at_rope<D, true>(x, P.cosT, P.sinT, key0 + j);
          }
          bf16* dst = P.dqkv + (size_t)(row0 + key0 + j) * P.ld_d + (cg == 0 ? P.H : 2 * P.H) + h * D;
#pragma unroll
          for (int c = 0; c < D; c += 8) *reinterpret_cast<uint4*>(dst + c) = at_pack8(&x[c]);
        }
      }
    } else {
#pragma unroll 1
      for (int pc = cg; pc < 2 * (D / 8); pc += AT_CG) {
        const bool is_v = pc >= D / 8;
        const int c8 = (is_v ? pc - D / 8 : pc) * 8;
        float o[8];
        tmem_ld_32x8(my_tmem + (is_v ? (kt ? cDV1 : cDV0) : (kt ? cDK1 : cDK0)) + c8, o);
        if (j < Tk) {
          if (!is_v) {
#pragma unroll
            for (int c = 0; c < 8; ++c) o[c] *= P.scale;
          }
          if (side) {
            const float a = is_v ? xpt[j] : xds[j];
            const float* xr = is_v ? xdo : xq;
#pragma unroll
            for (int c = 0; c < 8; ++c) o[c] = fmaf(a, xr[c8 + c], o[c]);
          }
          *reinterpret_cast<uint4*>(P.dqkv + (size_t)(row0 + key0 + j) * P.ld_d + (is_v ? 2 * P.H : P.H) + h * D + c8) = at_pack8(o);
        }
      }
    }
  }
  tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

constexpr int AT_KB = 128;   // keys per launch in key-blocked mode

constexpr int AT_FWD_SMEM = 16384 + 32768 + 32768 + 4 * 16384 + 1024 + 64 + 2 * AT_CG * 128 * 4 + 1024;
constexpr int AT_BWD_SMEM = 16384 + 16384 + 49152 + 49152 + 32768 + 32768 + 1024 + 4096;

static inline int at_kp(int T) { return (T + 15) / 16 * 16; }

}  // namespace vb

using namespace vb;

VB_TL_EXPORT(vitb200_tl_attn_fwd, vb::tl_attn_fwd)

// q, k, v must be the three column blocks of one fused [B*T, 3H] bf16 buffer (k = q + H, v = q + 2H, ld = 3H)
extern "C" int vitb200_attn_tc_supported(int T, int d, int ld, int H) {
  if (!(d == 16 || d == 32)) return 0;
  if (ld != 3 * H || (ld % 8) != 0) return 0;
  const int KP = at_kp(T), KPa = (KP + 31) / 32 * 32;
  if (KP > 256 || 2 * KPa + 32 * 5 > 512) return 0;  // TMEM columns of the backward kernel (T <= 176)
  return 1;
}

extern "C" int vitb200_attn_tc_fwd(const void* qkv, void* ctx, float* lse, const float* rope_cos, const float* rope_sin,
                                   int B, int T, int heads, int d, float scale, float p_drop, const uint64_t* rng,
                                   uint32_t site, void* stream) {
  if (!qkv || !ctx || !lse || B <= 0 || T <= 0 || heads <= 0) return VITB200_ERR_ARG;
  const int H = heads * d, ld = 3 * H;
  if (!vitb200_attn_tc_supported(T, d, ld, H)) return VITB200_ERR_SHAPE;
  const int M = B * T, KP = at_kp(T);
  CUtensorMap tQ, tKV;
  int rc;
  if ((rc = get_tmap(qkv, ld, M, 64, 128, &tQ))) return rc;
  if ((rc = get_tmap(qkv, ld, M, 64, KP, &tKV))) return rc;
  AttnTcParams P{(const bf16*)qkv, (bf16*)ctx, lse, nullptr, nullptr, 0, rope_cos, rope_sin, B, T, heads, H, ld, KP,
                 scale, p_drop, rng, site, 0, 0, 0, 0, 0, nullptr};
  dim3 grid(heads, B);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_F(DD)                                                                                             \
  {                                                                                                              \
    static bool done = false;                                                                                    \
    if (!done) {                                                                                                 \
      cudaError_t e = cudaFuncSetAttribute(attn_tc_fwd_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_FWD_SMEM); \
      if (e != cudaSuccess) return vb_cuda_error(e);                                                             \
      done = true;                                                                                               \
    }                                                                                                            \
    vb_launch_pdl(attn_tc_fwd_kernel<DD>, grid, dim3(AT_ALL_THREADS), AT_FWD_SMEM, st, tQ, tKV, P);                                \
  }
  if (d == 16) LAUNCH_F(16) else LAUNCH_F(32)
#undef LAUNCH_F
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

extern "C" int vitb200_attn_tc_bwd(const void* qkv, const void* ctx, const void* dctx, const float* lse, void* dqkv,
                                   const float* rope_cos, const float* rope_sin, int B, int T, int heads, int d,
                                   float scale, float p_drop, const uint64_t* rng, uint32_t site, void* stream) {
  if (!qkv || !ctx || !dctx || !lse || !dqkv || B <= 0 || T <= 0 || heads <= 0) return VITB200_ERR_ARG;
  const int H = heads * d, ld = 3 * H;
  if (!vitb200_attn_tc_supported(T, d, ld, H)) return VITB200_ERR_SHAPE;
  const int M = B * T, KP = at_kp(T);
  CUtensorMap tQ, tKV, tDO;
  int rc;
  if ((rc = get_tmap(qkv, ld, M, 64, 128, &tQ))) return rc;
  if ((rc = get_tmap(qkv, ld, M, 64, KP, &tKV))) return rc;
  if ((rc = get_tmap(dctx, H, M, 64, 128, &tDO))) return rc;
  AttnTcParams P{(const bf16*)qkv, (bf16*)const_cast<void*>(ctx), const_cast<float*>(lse), (const bf16*)dctx, (bf16*)dqkv,
                 ld, rope_cos, rope_sin, B, T, heads, H, ld, KP, scale, p_drop, rng, site, 0, 0, 0, 0, 0, nullptr};
  dim3 grid(heads, B);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_B(DD)                                                                                             \
  {                                                                                                              \
    static bool done = false;                                                                                    \
    if (!done) {                                                                                                 \
      cudaError_t e = cudaFuncSetAttribute(attn_tc_bwd_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_BWD_SMEM); \
      if (e != cudaSuccess) return vb_cuda_error(e);                                                             \
      done = true;                                                                                               \
    }                                                                                                            \
    vb_launch_pdl(attn_tc_bwd_kernel<DD>, grid, dim3(AT_ALL_THREADS), AT_BWD_SMEM, st, tQ, tKV, tDO, P);                           \
  }
  if (d == 16) LAUNCH_B(16) else LAUNCH_B(32)
#undef LAUNCH_B
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

// ---- key-blocked mode: any sequence length (the long-sequence sweep: T = 510 / 513 / 2034 / 2049) -------------------
// The single-tile kernels above keep a whole key range in one TMEM tile.  Longer sequences run the one-launch flash
// kernels of attention_flash.cu (forward and backward).  What remains here is the per-key-block BACKWARD: the single-tile
// backward kernel launched once per block of 128 keys (P recomputed from the lse, dK / dV of a block complete, dQ
// accumulated over the launches in fp32).  It is the independent cross-check of the flash backward in the tests (same
// products, same accumulation order => bit-identical results) and is not used by the engine.
extern "C" int vitb200_attn_tc_blocked_supported(int T, int d, int ld, int H) {
  if (!(d == 16 || d == 32)) return 0;
  if (ld != 3 * H || (ld % 8) != 0 || (H % 8) != 0) return 0;
  return T > 0 ? 1 : 0;
}
extern "C" size_t vitb200_attn_tc_blocked_ws_bytes(int B, int T, int heads, int d) {
  return sizeof(float) * (size_t)B * T * heads * d + 256;   // the fp32 dQ accumulator
}

extern "C" int vitb200_attn_tc_blocked_bwd(const void* qkv, const void* ctx, const void* dctx, const float* lse, void* dqkv,
                                           const float* rope_cos, const float* rope_sin, int B, int T, int heads, int d,
                                           float scale, float p_drop, const uint64_t* rng, uint32_t site, void* ws,
                                           void* stream) {
  if (!qkv || !ctx || !dctx || !lse || !dqkv || !ws || B <= 0 || T <= 0 || heads <= 0) return VITB200_ERR_ARG;
  const int H = heads * d, ld = 3 * H;
  if (!vitb200_attn_tc_blocked_supported(T, d, ld, H)) return VITB200_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(ws) & 15) != 0) return VITB200_ERR_ALIGN;
  const int M = B * T, nblk = (T + AT_KB - 1) / AT_KB;
  float* dq_acc = reinterpret_cast<float*>(ws);
  CUtensorMap tQ, tKV, tDO;
  int rc;
  if ((rc = get_tmap(qkv, ld, M, 64, 128, &tQ))) return rc;
  if ((rc = get_tmap(dctx, H, M, 64, 128, &tDO))) return rc;
  dim3 grid(heads, B);
  cudaStream_t st = (cudaStream_t)stream;
  static bool done16 = false, done32 = false;
  bool& done = d == 16 ? done16 : done32;
  if (!done) {
    cudaError_t e = d == 16 ? cudaFuncSetAttribute(attn_tc_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_BWD_SMEM)
                            : cudaFuncSetAttribute(attn_tc_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_BWD_SMEM);
    if (e != cudaSuccess) return vb_cuda_error(e);
    done = true;
  }
  for (int kb = 0; kb < nblk; ++kb) {
    const int key0 = kb * AT_KB, Tk = T - key0 < AT_KB ? T - key0 : AT_KB, KP = at_kp(Tk);
    if ((rc = get_tmap(qkv, ld, M, 64, KP, &tKV))) return rc;
    AttnTcParams P{(const bf16*)qkv, (bf16*)const_cast<void*>(ctx), const_cast<float*>(lse), (const bf16*)dctx, (bf16*)dqkv,
                   ld, rope_cos, rope_sin, B, T, heads, H, ld, KP, scale, p_drop, rng, site, 1, key0, Tk, kb == 0,
                   kb == nblk - 1, dq_acc};
    if (d == 16) vb_launch_pdl(attn_tc_bwd_kernel<16>, grid, dim3(AT_ALL_THREADS), AT_BWD_SMEM, st, tQ, tKV, tDO, P);
    else vb_launch_pdl(attn_tc_bwd_kernel<32>, grid, dim3(AT_ALL_THREADS), AT_BWD_SMEM, st, tQ, tKV, tDO, P);
    VB_CHECK_LAUNCH();
  }
  return VITB200_OK;
}
