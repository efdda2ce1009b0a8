// Multi-head self-attention, flash style (online softmax, probabilities never stored), SIMT fp32 math.
// Used for FP32 mode (1e-4 parity bar) and as the comparator / fallback-free general path for head dims
// the tcgen05 kernel does not cover.  One query (or key) row per group of TPR lanes, DPT dims per lane,
// head_dim = DPT * TPR; K/V (or Q/dO) tiles are staged in shared memory and read as broadcasts.
#include "common.cuh"

namespace vb {

constexpr int AT_THREADS = 128;
constexpr float LOG2E = 1.4426950408889634f;

template <int D> struct AttnTile { static constexpr int KT = D >= 128 ? 32 : 64; };

// rope on a register fragment: dims [part*DPT, part*DPT+DPT) of a row at position t.
// fwd: x' = x*cos + rot(x)*sin ; inverse (transpose): dx = dx'*cos - rot(dx')*sin   (rope.py:60-98)
template <int DPT, int TPR, bool INVERSE>
__device__ __forceinline__ void rope_frag(float (&x)[DPT], const float* __restrict__ cosT, const float* __restrict__ sinT,
                                          int t, int part) {
  constexpr int D = DPT * TPR, HALF = D / 2;
  if (TPR == 1) {
#pragma unroll
    for (int c = 0; c < DPT / 2; ++c) {
      float cs = cosT[(size_t)t * HALF + c], sn = sinT[(size_t)t * HALF + c];
      if (INVERSE) sn = -sn;
      float lo = x[c], hi = x[c + DPT / 2];
      x[c] = lo * cs - hi * sn;
      x[c + DPT / 2] = hi * cs + lo * sn;
    }
  } else {
    const bool first = part < TPR / 2;
    const int jbase = (first ? part : part - TPR / 2) * DPT;
#pragma unroll
    for (int c = 0; c < DPT; ++c) {
      float other = __shfl_xor_sync(0xffffffffu, x[c], TPR / 2 > 0 ? TPR / 2 : 1);
      float cs = cosT[(size_t)t * HALF + jbase + c], sn = sinT[(size_t)t * HALF + jbase + c];
      if (INVERSE) sn = -sn;
      x[c] = first ? x[c] * cs - other * sn : x[c] * cs + other * sn;
    }
  }
}

// Stage `rows` rows [r0, r0+KT) of one head into smem as fp32 (rows >= T zero-filled), optional RoPE.
template <typename T, int D, int KT, bool ROUND>
__device__ __forceinline__ void load_tile(float (*S)[D], const T* __restrict__ base, int ld, int r0, int Tlen,
                                          const float* __restrict__ cosT, const float* __restrict__ sinT, int tid) {
  if constexpr (D == 4) {   // head_dim 4 (hidden 32 / 8 heads of the sweep grid): one 4-vector per row, rotation pairs (0,2), (1,3)
    for (int j = tid; j < KT; j += AT_THREADS) {
      const int t = r0 + j;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < Tlen) {
        x = Vec4<T>::ld(base + (size_t)t * ld);
        if (cosT) {
          const float c0 = cosT[(size_t)t * 2], c1 = cosT[(size_t)t * 2 + 1], s0 = sinT[(size_t)t * 2], s1 = sinT[(size_t)t * 2 + 1];
          float4 y = make_float4(x.x * c0 - x.z * s0, x.y * c1 - x.w * s1, x.z * c0 + x.x * s0, x.w * c1 + x.y * s1);
          if (ROUND) { y.x = round_to<T>(y.x); y.y = round_to<T>(y.y); y.z = round_to<T>(y.z); y.w = round_to<T>(y.w); }
          x = y;
        }
      }
      *reinterpret_cast<float4*>(&S[j][0]) = x;
    }
    return;
  }
  constexpr int HALF = D / 2, CH = HALF / 4 > 0 ? HALF / 4 : 1;
  for (int idx = tid; idx < KT * CH; idx += AT_THREADS) {
    const int j = idx / CH, c = (idx % CH) * 4;
    const int t = r0 + j;
    float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
    if (t < Tlen) {
      const T* p = base + (size_t)t * ld;
      lo = Vec4<T>::ld(p + c);
      hi = Vec4<T>::ld(p + c + HALF);
      if (cosT) {
        float4 cs = *reinterpret_cast<const float4*>(cosT + (size_t)t * HALF + c);
        float4 sn = *reinterpret_cast<const float4*>(sinT + (size_t)t * HALF + c);
        float4 l2, h2;
        l2.x = lo.x * cs.x - hi.x * sn.x; h2.x = hi.x * cs.x + lo.x * sn.x;
        l2.y = lo.y * cs.y - hi.y * sn.y; h2.y = hi.y * cs.y + lo.y * sn.y;
        l2.z = lo.z * cs.z - hi.z * sn.z; h2.z = hi.z * cs.z + lo.z * sn.z;
        l2.w = lo.w * cs.w - hi.w * sn.w; h2.w = hi.w * cs.w + lo.w * sn.w;
        if (ROUND) {
          l2.x = round_to<T>(l2.x); l2.y = round_to<T>(l2.y); l2.z = round_to<T>(l2.z); l2.w = round_to<T>(l2.w);
          h2.x = round_to<T>(h2.x); h2.y = round_to<T>(h2.y); h2.z = round_to<T>(h2.z); h2.w = round_to<T>(h2.w);
        }
        lo = l2; hi = h2;
      }
    }
    *reinterpret_cast<float4*>(&S[j][c]) = lo;
    *reinterpret_cast<float4*>(&S[j][c + HALF]) = hi;
  }
}

template <typename T, int DPT>
__device__ __forceinline__ void load_frag(float (&x)[DPT], const T* __restrict__ p) {
#pragma unroll
  for (int c = 0; c < DPT; c += 4) {
    float4 v = Vec4<T>::ld(p + c);
    x[c] = v.x; x[c + 1] = v.y; x[c + 2] = v.z; x[c + 3] = v.w;
  }
}
template <typename T, int DPT>
__device__ __forceinline__ void store_frag(T* __restrict__ p, const float (&x)[DPT]) {
#pragma unroll
  for (int c = 0; c < DPT; c += 4) Vec4<T>::st(p + c, make_float4(x[c], x[c + 1], x[c + 2], x[c + 3]));
}
template <int DPT>
__device__ __forceinline__ float dot_frag(const float (&a)[DPT], const float* __restrict__ s) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < DPT; c += 4) {
    float4 v = *reinterpret_cast<const float4*>(s + c);
    acc = fmaf(a[c], v.x, acc); acc = fmaf(a[c + 1], v.y, acc);
    acc = fmaf(a[c + 2], v.z, acc); acc = fmaf(a[c + 3], v.w, acc);
  }
  return acc;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <typename T, int DPT, int TPR>
__global__ void __launch_bounds__(AT_THREADS)
attn_fwd_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, int ld, T* __restrict__ ctx,
                float* __restrict__ lse, const float* __restrict__ cosT, const float* __restrict__ sinT, int Tlen,
                int heads, float scale, float p_drop, const uint64_t* __restrict__ rng, uint32_t site) {
  constexpr int D = DPT * TPR, ROWS = AT_THREADS / TPR, KT = AttnTile<D>::KT;
  __shared__ __align__(16) float Ks[KT][D];
  __shared__ __align__(16) float Vs[KT][D];
  const int tid = threadIdx.x, r = tid / TPR, part = tid % TPR;
  const int h = blockIdx.y, b = blockIdx.z;
  const int i = blockIdx.x * ROWS + r;
  const bool valid = i < Tlen;
  const int ic = valid ? i : Tlen - 1;
  const size_t row0 = (size_t)b * Tlen;
  const int Hd = heads * D;
  const int Tpad = attn_drop_tpad(Tlen);
  const DropCtx dc = make_drop(p_drop, rng ? rng[0] : 0ull, rng ? (uint32_t)rng[1] : 0u, site);
  const float sl2 = scale * LOG2E;

  float qf[DPT], o[DPT];
  load_frag<T, DPT>(qf, q + (row0 + ic) * ld + h * D + part * DPT);
  if (cosT) {
    rope_frag<DPT, TPR, false>(qf, cosT, sinT, ic, part);
#pragma unroll
    for (int c = 0; c < DPT; ++c) qf[c] = round_to<T>(qf[c]);
  }
#pragma unroll
  for (int c = 0; c < DPT; ++c) o[c] = 0.f;
  float m = -INFINITY, l = 0.f;  // running max (log2 units) and sum
  const uint64_t drow = ((uint64_t)(b * heads + h) * Tlen + ic) * (uint64_t)Tpad;

  for (int kt0 = 0; kt0 < Tlen; kt0 += KT) {
    __syncthreads();
    load_tile<T, D, KT, true>(Ks, k + row0 * ld + h * D, ld, kt0, Tlen, cosT, sinT, tid);
    load_tile<T, D, KT, false>(Vs, v + row0 * ld + h * D, ld, kt0, Tlen, nullptr, nullptr, tid);
    __syncthreads();
    const int nk = min(KT, Tlen - kt0);
    for (int j0 = 0; j0 < nk; j0 += 4) {
      float s[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        float d = group_sum<TPR>(dot_frag<DPT>(qf, &Ks[j0 + jj][part * DPT]));
        s[jj] = (j0 + jj < nk) ? d * sl2 : -INFINITY;
      }
      const float mx = fmaxf(fmaxf(m, fmaxf(s[0], s[1])), fmaxf(s[2], s[3]));
      const float corr = exp2f(m - mx);
      float pj[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) pj[jj] = exp2f(s[jj] - mx);
      l = l * corr + (pj[0] + pj[1]) + (pj[2] + pj[3]);
      m = mx;
      if (dc.on) {
        const float4 kp = drop4(dc, (drow + (uint64_t)(kt0 + j0)) >> 2);
        pj[0] *= kp.x; pj[1] *= kp.y; pj[2] *= kp.z; pj[3] *= kp.w;
      }
#pragma unroll
      for (int c = 0; c < DPT; ++c) {
        float a = o[c] * corr;
        a = fmaf(pj[0], Vs[j0 + 0][part * DPT + c], a);
        a = fmaf(pj[1], Vs[j0 + 1][part * DPT + c], a);
        a = fmaf(pj[2], Vs[j0 + 2][part * DPT + c], a);
        a = fmaf(pj[3], Vs[j0 + 3][part * DPT + c], a);
        o[c] = a;
      }
    }
  }
  if (valid) {
    const float inv = 1.f / l;
#pragma unroll
    for (int c = 0; c < DPT; ++c) o[c] *= inv;
    store_frag<T, DPT>(ctx + (row0 + i) * Hd + h * D + part * DPT, o);
    if (part == 0) lse[(size_t)(b * heads + h) * Tlen + i] = (m + log2f(l)) * (1.f / LOG2E);
  }
}

// ------------------------------------------------------------------------------------------------
// backward, pass 1: dQ (one query row per lane group) and dsum_i = dO_i . O_i
// ------------------------------------------------------------------------------------------------
template <typename T, int DPT, int TPR>
__global__ void __launch_bounds__(AT_THREADS)
attn_bwd_dq_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, int ld,
                   const T* __restrict__ ctx, const T* __restrict__ dctx, const float* __restrict__ lse,
                   float* __restrict__ dsum, T* __restrict__ dq, int ld_d, const float* __restrict__ cosT,
                   const float* __restrict__ sinT, int Tlen, int heads, float scale, float p_drop,
                   const uint64_t* __restrict__ rng, uint32_t site) {
  constexpr int D = DPT * TPR, ROWS = AT_THREADS / TPR, KT = AttnTile<D>::KT;
  __shared__ __align__(16) float Ks[KT][D];
  __shared__ __align__(16) float Vs[KT][D];
  const int tid = threadIdx.x, r = tid / TPR, part = tid % TPR;
  const int h = blockIdx.y, b = blockIdx.z;
  const int i = blockIdx.x * ROWS + r;
  const bool valid = i < Tlen;
  const int ic = valid ? i : Tlen - 1;
  const size_t row0 = (size_t)b * Tlen;
  const int Hd = heads * D;
  const int Tpad = attn_drop_tpad(Tlen);
  const DropCtx dc = make_drop(p_drop, rng ? rng[0] : 0ull, rng ? (uint32_t)rng[1] : 0u, site);
  const float sl2 = scale * LOG2E;

  float qf[DPT], dof[DPT], acc[DPT];
  load_frag<T, DPT>(qf, q + (row0 + ic) * ld + h * D + part * DPT);
  if (cosT) {
    rope_frag<DPT, TPR, false>(qf, cosT, sinT, ic, part);
#pragma unroll
    for (int c = 0; c < DPT; ++c) qf[c] = round_to<T>(qf[c]);
  }
  load_frag<T, DPT>(dof, dctx + (row0 + ic) * Hd + h * D + part * DPT);
  float Di;
  {
    float of[DPT];
    load_frag<T, DPT>(of, ctx + (row0 + ic) * Hd + h * D + part * DPT);
    float t = 0.f;
#pragma unroll
    for (int c = 0; c < DPT; ++c) t = fmaf(dof[c], of[c], t);
    Di = group_sum<TPR>(t);
  }
  const size_t si = (size_t)(b * heads + h) * Tlen + ic;
  const float lse2 = lse[si] * LOG2E;
  if (valid && part == 0) dsum[si] = Di;
#pragma unroll
  for (int c = 0; c < DPT; ++c) acc[c] = 0.f;
  const uint64_t drow = ((uint64_t)(b * heads + h) * Tlen + ic) * (uint64_t)Tpad;

  for (int kt0 = 0; kt0 < Tlen; kt0 += KT) {
    __syncthreads();
    load_tile<T, D, KT, true>(Ks, k + row0 * ld + h * D, ld, kt0, Tlen, cosT, sinT, tid);
    load_tile<T, D, KT, false>(Vs, v + row0 * ld + h * D, ld, kt0, Tlen, nullptr, nullptr, tid);
    __syncthreads();
    const int nk = min(KT, Tlen - kt0);
    for (int j0 = 0; j0 < nk; j0 += 4) {
      float4 kp = make_float4(1.f, 1.f, 1.f, 1.f);
      if (dc.on) kp = drop4(dc, (drow + (uint64_t)(kt0 + j0)) >> 2);
      const float kpa[4] = {kp.x, kp.y, kp.z, kp.w};
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float s = group_sum<TPR>(dot_frag<DPT>(qf, &Ks[j0 + jj][part * DPT]));
        const float dp = group_sum<TPR>(dot_frag<DPT>(dof, &Vs[j0 + jj][part * DPT]));
        const float p = (j0 + jj < nk) ? exp2f(s * sl2 - lse2) : 0.f;
        const float ds = p * (dp * kpa[jj] - Di);
#pragma unroll
        for (int c = 0; c < DPT; ++c) acc[c] = fmaf(ds, Ks[j0 + jj][part * DPT + c], acc[c]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < DPT; ++c) acc[c] *= scale;
  if (cosT) {
#pragma unroll
    for (int c = 0; c < DPT; ++c) acc[c] = round_to<T>(acc[c]);
    rope_frag<DPT, TPR, true>(acc, cosT, sinT, ic, part);
  }
  if (valid) store_frag<T, DPT>(dq + (row0 + i) * ld_d + h * D + part * DPT, acc);
}

// ------------------------------------------------------------------------------------------------
// backward, pass 2: dK, dV (one key row per lane group; loops over query tiles)
// ------------------------------------------------------------------------------------------------
template <typename T, int DPT, int TPR>
__global__ void __launch_bounds__(AT_THREADS)
attn_bwd_dkv_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, int ld,
                    const T* __restrict__ dctx, const float* __restrict__ lse, const float* __restrict__ dsum,
                    T* __restrict__ dk, T* __restrict__ dv, int ld_d, const float* __restrict__ cosT,
                    const float* __restrict__ sinT, int Tlen, int heads, float scale, float p_drop,
                    const uint64_t* __restrict__ rng, uint32_t site) {
  constexpr int D = DPT * TPR, ROWS = AT_THREADS / TPR, QT = AttnTile<D>::KT;
  __shared__ __align__(16) float Qs[QT][D];
  __shared__ __align__(16) float Os[QT][D];
  __shared__ float Ls[QT], Ds[QT];
  const int tid = threadIdx.x, r = tid / TPR, part = tid % TPR;
  const int h = blockIdx.y, b = blockIdx.z;
  const int j = blockIdx.x * ROWS + r;
  const bool valid = j < Tlen;
  const int jc = valid ? j : Tlen - 1;
  const size_t row0 = (size_t)b * Tlen;
  const int Hd = heads * D;
  const int Tpad = attn_drop_tpad(Tlen);
  const DropCtx dc = make_drop(p_drop, rng ? rng[0] : 0ull, rng ? (uint32_t)rng[1] : 0u, site);
  const float sl2 = scale * LOG2E;

  float kf[DPT], vf[DPT], ak[DPT], av[DPT];
  load_frag<T, DPT>(kf, k + (row0 + jc) * ld + h * D + part * DPT);
  if (cosT) {
    rope_frag<DPT, TPR, false>(kf, cosT, sinT, jc, part);
#pragma unroll
    for (int c = 0; c < DPT; ++c) kf[c] = round_to<T>(kf[c]);
  }
  load_frag<T, DPT>(vf, v + (row0 + jc) * ld + h * D + part * DPT);
#pragma unroll
  for (int c = 0; c < DPT; ++c) { ak[c] = 0.f; av[c] = 0.f; }
  const uint64_t dbase = (uint64_t)(b * heads + h) * Tlen;

  for (int it0 = 0; it0 < Tlen; it0 += QT) {
    __syncthreads();
    load_tile<T, D, QT, true>(Qs, q + row0 * ld + h * D, ld, it0, Tlen, cosT, sinT, tid);
    load_tile<T, D, QT, false>(Os, dctx + row0 * Hd + h * D, Hd, it0, Tlen, nullptr, nullptr, tid);
    for (int x = tid; x < QT; x += AT_THREADS) {
      const int t = it0 + x;
      Ls[x] = t < Tlen ? lse[dbase + t] * LOG2E : 0.f;
      Ds[x] = t < Tlen ? dsum[dbase + t] : 0.f;
    }
    __syncthreads();
    const int nq = min(QT, Tlen - it0);
    for (int x = 0; x < nq; ++x) {
      const float s = group_sum<TPR>(dot_frag<DPT>(kf, &Qs[x][part * DPT]));
      const float dp = group_sum<TPR>(dot_frag<DPT>(vf, &Os[x][part * DPT]));
      const float p = exp2f(s * sl2 - Ls[x]);
      const float keep = dc.on ? drop1(dc, (dbase + (uint64_t)(it0 + x)) * (uint64_t)Tpad + (uint64_t)jc) : 1.f;
      const float pd = p * keep;
      const float ds = p * (dp * keep - Ds[x]);
#pragma unroll
      for (int c = 0; c < DPT; ++c) {
        av[c] = fmaf(pd, Os[x][part * DPT + c], av[c]);
        ak[c] = fmaf(ds, Qs[x][part * DPT + c], ak[c]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < DPT; ++c) ak[c] *= scale;
  if (cosT) {
#pragma unroll
    for (int c = 0; c < DPT; ++c) ak[c] = round_to<T>(ak[c]);
    rope_frag<DPT, TPR, true>(ak, cosT, sinT, jc, part);
  }
  if (valid) {
    store_frag<T, DPT>(dk + (row0 + j) * ld_d + h * D + part * DPT, ak);
    store_frag<T, DPT>(dv + (row0 + j) * ld_d + h * D + part * DPT, av);
  }
}

// ------------------------------------------------------------------------------------------------
// attention probabilities (viz / output_attentions path only)
// ------------------------------------------------------------------------------------------------
template <typename T, int DPT, int TPR>
__global__ void __launch_bounds__(AT_THREADS)
attn_probs_kernel(const T* __restrict__ q, const T* __restrict__ k, int ld, const float* __restrict__ lse,
                  float* __restrict__ probs, const float* __restrict__ cosT, const float* __restrict__ sinT, int Tlen,
                  int heads, float scale) {
  constexpr int D = DPT * TPR, ROWS = AT_THREADS / TPR, KT = AttnTile<D>::KT;
  __shared__ __align__(16) float Ks[KT][D];
  const int tid = threadIdx.x, r = tid / TPR, part = tid % TPR;
  const int h = blockIdx.y, b = blockIdx.z;
  const int i = blockIdx.x * ROWS + r;
  const bool valid = i < Tlen;
  const int ic = valid ? i : Tlen - 1;
  const size_t row0 = (size_t)b * Tlen;
  const float sl2 = scale * LOG2E;
  float qf[DPT];
  load_frag<T, DPT>(qf, q + (row0 + ic) * ld + h * D + part * DPT);
  if (cosT) {
    rope_frag<DPT, TPR, false>(qf, cosT, sinT, ic, part);
#pragma unroll
    for (int c = 0; c < DPT; ++c) qf[c] = round_to<T>(qf[c]);
  }
  const size_t si = (size_t)(b * heads + h) * Tlen + ic;
  const float lse2 = lse[si] * LOG2E;
  for (int kt0 = 0; kt0 < Tlen; kt0 += KT) {
    __syncthreads();
    load_tile<T, D, KT, true>(Ks, k + row0 * ld + h * D, ld, kt0, Tlen, cosT, sinT, tid);
    __syncthreads();
    const int nk = min(KT, Tlen - kt0);
    for (int jj = 0; jj < nk; ++jj) {
      const float s = group_sum<TPR>(dot_frag<DPT>(qf, &Ks[jj][part * DPT]));
      if (valid && part == 0) probs[si * Tlen + kt0 + jj] = exp2f(s * sl2 - lse2);
    }
  }
}

template <typename T>
static int attn_dispatch(int which, const void* q, const void* k, const void* v, int ld, const void* ctx, const void* dctx,
                         void* out_ctx, float* lse, float* dsum, void* dq, void* dk, void* dv, int ld_d, float* probs,
                         const float* cosT, const float* sinT, int B, int Tlen, int heads, int d, float scale,
                         float p_drop, const uint64_t* rng, uint32_t site, cudaStream_t st) {
#define ATTN_CASE(DPT, TPR)                                                                                          \
  {                                                                                                                  \
    constexpr int ROWS = AT_THREADS / TPR;                                                                           \
    dim3 grid(ceil_div(Tlen, ROWS), heads, B);                                                                       \
    if (which == 0)                                                                                                  \
      attn_fwd_kernel<T, DPT, TPR><<<grid, AT_THREADS, 0, st>>>((const T*)q, (const T*)k, (const T*)v, ld,            \
                                                                (T*)out_ctx, lse, cosT, sinT, Tlen, heads, scale,     \
                                                                p_drop, rng, site);                                   \
    else if (which == 1) {                                                                                           \
      attn_bwd_dq_kernel<T, DPT, TPR><<<grid, AT_THREADS, 0, st>>>((const T*)q, (const T*)k, (const T*)v, ld,         \
                                                                   (const T*)ctx, (const T*)dctx, lse, dsum, (T*)dq,  \
                                                                   ld_d, cosT, sinT, Tlen, heads, scale, p_drop, rng, \
                                                                   site);                                             \
      attn_bwd_dkv_kernel<T, DPT, TPR><<<grid, AT_THREADS, 0, st>>>((const T*)q, (const T*)k, (const T*)v, ld,        \
                                                                    (const T*)dctx, lse, dsum, (T*)dk, (T*)dv, ld_d,  \
                                                                    cosT, sinT, Tlen, heads, scale, p_drop, rng,      \
                                                                    site);                                            \
    } else                                                                                                           \
      attn_probs_kernel<T, DPT, TPR><<<grid, AT_THREADS, 0, st>>>((const T*)q, (const T*)k, ld, lse, probs, cosT,     \
                                                                  sinT, Tlen, heads, scale);                          \
  }
  switch (d) {
    case 4: ATTN_CASE(4, 1) break;
    case 8: ATTN_CASE(8, 1) break;
    case 16: ATTN_CASE(16, 1) break;
    case 32: ATTN_CASE(16, 2) break;
    case 64: ATTN_CASE(16, 4) break;
    case 128: ATTN_CASE(16, 8) break;
    default: return VITB200_ERR_SHAPE;
  }
#undef ATTN_CASE
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

}  // namespace vb

using namespace vb;

extern "C" int vitb200_attn_tc_supported(int T, int d, int ld, int H);
extern "C" int vitb200_attn_tc_fwd(const void*, void*, float*, const float*, const float*, int, int, int, int, float, float,
                                   const uint64_t*, uint32_t, void*);
extern "C" int vitb200_attn_tc_bwd(const void*, const void*, const void*, const float*, void*, const float*, const float*,
                                   int, int, int, int, float, float, const uint64_t*, uint32_t, void*);
// 0 = automatic (tcgen05 kernels when q/k/v are the column blocks of one fused bf16 QKV buffer), 1 = SIMT only
static int g_attn_mode = 0;
extern "C" int vitb200_set_attn_mode(int mode) {
  int old = g_attn_mode;
  if (mode == 0 || mode == 1) g_attn_mode = mode;
  return old;
}
static inline bool fused_qkv(const void* q, const void* k, const void* v, int ld, int heads, int d) {
  const int H = heads * d;
  return ld == 3 * H && (const char*)k == (const char*)q + (size_t)H * 2 && (const char*)v == (const char*)q + (size_t)H * 4 &&
         (reinterpret_cast<uintptr_t>(q) & 15) == 0;
}

static int attn_check(int ld, int B, int Tlen, int heads, int d) {
  if (B < 0 || Tlen <= 0 || heads <= 0 || d <= 0) return VITB200_ERR_ARG;
  if (!(d == 4 || d == 8 || d == 16 || d == 32 || d == 64 || d == 128)) return VITB200_ERR_SHAPE;
  if (ld % 4 != 0 || ld < heads * d) return VITB200_ERR_SHAPE;
  if (heads > 65535 || B > 65535) return VITB200_ERR_SHAPE;
  return VITB200_OK;
}

extern "C" int vitb200_attn_fwd(const void* q, const void* k, const void* v, int ld, void* ctx, float* lse,
                                const float* rope_cos, const float* rope_sin, int B, int T, int heads, int d,
                                float scale, float p_drop, const uint64_t* rng, uint32_t site, int dtype,
                                void* stream) {
  int rc = attn_check(ld, B, T, heads, d);
  if (rc) return rc;
  if (!q || !k || !v || !ctx || !lse || ((rope_cos == nullptr) != (rope_sin == nullptr))) return VITB200_ERR_ARG;
  if (B == 0) return VITB200_OK;
  if (dtype == VITB200_BF16 && g_attn_mode == 0 && fused_qkv(q, k, v, ld, heads, d) &&
      vitb200_attn_tc_supported(T, d, ld, heads * d) && (reinterpret_cast<uintptr_t>(ctx) & 15) == 0)
    return vitb200_attn_tc_fwd(q, ctx, lse, rope_cos, rope_sin, B, T, heads, d, scale, p_drop, rng, site, stream);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VITB200_F32)
    return attn_dispatch<float>(0, q, k, v, ld, nullptr, nullptr, ctx, lse, nullptr, nullptr, nullptr, nullptr, 0,
                                nullptr, rope_cos, rope_sin, B, T, heads, d, scale, p_drop, rng, site, st);
  if (dtype == VITB200_BF16)
    return attn_dispatch<bf16>(0, q, k, v, ld, nullptr, nullptr, ctx, lse, nullptr, nullptr, nullptr, nullptr, 0,
                               nullptr, rope_cos, rope_sin, B, T, heads, d, scale, p_drop, rng, site, st);
  return VITB200_ERR_ARG;
}

extern "C" int vitb200_attn_bwd(const void* q, const void* k, const void* v, int ld, const void* ctx, const void* dctx,
                                const float* lse, float* dsum, void* dq, void* dk, void* dv, int ld_d,
                                const float* rope_cos, const float* rope_sin, int B, int T, int heads, int d,
                                float scale, float p_drop, const uint64_t* rng, uint32_t site, int dtype,
                                void* stream) {
  int rc = attn_check(ld, B, T, heads, d);
  if (rc) return rc;
  if (ld_d % 4 != 0) return VITB200_ERR_SHAPE;
  if (!q || !k || !v || !ctx || !dctx || !lse || !dsum || !dq || !dk || !dv) return VITB200_ERR_ARG;
  if ((rope_cos == nullptr) != (rope_sin == nullptr)) return VITB200_ERR_ARG;
  if (B == 0) return VITB200_OK;
  if (dtype == VITB200_BF16 && g_attn_mode == 0 && fused_qkv(q, k, v, ld, heads, d) && ld_d == ld &&
      fused_qkv(dq, dk, dv, ld_d, heads, d) && vitb200_attn_tc_supported(T, d, ld, heads * d) &&
      ((reinterpret_cast<uintptr_t>(ctx) | reinterpret_cast<uintptr_t>(dctx)) & 15) == 0)
    return vitb200_attn_tc_bwd(q, ctx, dctx, lse, dq, rope_cos, rope_sin, B, T, heads, d, scale, p_drop, rng, site, stream);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VITB200_F32)
    return attn_dispatch<float>(1, q, k, v, ld, ctx, dctx, nullptr, const_cast<float*>(lse), dsum, dq, dk, dv, ld_d,
                                nullptr, rope_cos, rope_sin, B, T, heads, d, scale, p_drop, rng, site, st);
  if (dtype == VITB200_BF16)
    return attn_dispatch<bf16>(1, q, k, v, ld, ctx, dctx, nullptr, const_cast<float*>(lse), dsum, dq, dk, dv, ld_d,
                               nullptr, rope_cos, rope_sin, B, T, heads, d, scale, p_drop, rng, site, st);
  return VITB200_ERR_ARG;
}

extern "C" int vitb200_attn_probs(const void* q, const void* k, int ld, const float* lse, float* probs,
                                  const float* rope_cos, const float* rope_sin, int B, int T, int heads, int d,
                                  float scale, int dtype, void* stream) {
  int rc = attn_check(ld, B, T, heads, d);
  if (rc) return rc;
  if (!q || !k || !lse || !probs) return VITB200_ERR_ARG;
  if (B == 0) return VITB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VITB200_F32)
    return attn_dispatch<float>(2, q, k, nullptr, ld, nullptr, nullptr, nullptr, const_cast<float*>(lse), nullptr,
                                nullptr, nullptr, nullptr, 0, probs, rope_cos, rope_sin, B, T, heads, d, scale, 0.f,
                                nullptr, 0, st);
  if (dtype == VITB200_BF16)
    return attn_dispatch<bf16>(2, q, k, nullptr, ld, nullptr, nullptr, nullptr, const_cast<float*>(lse), nullptr,
                               nullptr, nullptr, nullptr, 0, probs, rope_cos, rope_sin, B, T, heads, d, scale, 0.f,
                               nullptr, 0, st);
  return VITB200_ERR_ARG;
}
