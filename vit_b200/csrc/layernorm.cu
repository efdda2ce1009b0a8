// Residual add + hidden dropout + LayerNorm, forward and backward (HBM-bound: 128-bit loads, a group
// of `lpr` lanes per row with shuffle reductions, the row kept in registers between the two passes).
#include "common.cuh"

namespace vb {

constexpr int LN_THREADS = 256;
constexpr int LN_MAX_BLOCKS = 592;  // 4 x 148 SMs

__device__ __forceinline__ float group_sum_rt(float v, int lpr) {
  for (int o = lpr >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static inline int ln_lpr(int H) {
  int chunks = H / 4, l = 1;
  while (l < chunks && l < 32) l <<= 1;
  return l;
}

template <typename T, int MAXC>
__global__ void __launch_bounds__(LN_THREADS)
add_ln_fwd_kernel(const float* __restrict__ z_in, const T* __restrict__ delta, float* __restrict__ z_out,
                  T* __restrict__ u, float* __restrict__ mean, float* __restrict__ rstd,
                  const float* __restrict__ gamma, const float* __restrict__ beta, int M, int H, int cls_T, float eps,
                  float p_drop, const uint64_t* __restrict__ rng, uint32_t site, int lpr) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rpw = 32 / lpr, sub = lane / lpr, l = lane % lpr;
  const int nchunk = H >> 2;
  const DropCtx dc = make_drop(delta ? p_drop : 0.f, rng ? rng[0] : 0ull, rng ? (uint32_t)rng[1] : 0u, site);
  const float invH = 1.f / (float)H;
  const long long rows_per_block = (long long)(LN_THREADS / 32) * rpw;
  for (long long base = (long long)blockIdx.x * rows_per_block; base < M; base += (long long)gridDim.x * rows_per_block) {
    const long long row = base + (long long)warp * rpw + sub;
    const bool active = row < M;
    float4 x[MAXC];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ci = l + c * lpr;
      x[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (active && ci < nchunk) {
        const size_t o = (size_t)row * H + (size_t)ci * 4;
        float4 v = *reinterpret_cast<const float4*>(z_in + o);
        if (delta) {
          float4 d = Vec4<T>::ld(delta + o);
          float4 k = drop4(dc, o >> 2);
          v.x += round_to<T>(d.x * k.x); v.y += round_to<T>(d.y * k.y);
          v.z += round_to<T>(d.z * k.z); v.w += round_to<T>(d.w * k.w);
          *reinterpret_cast<float4*>(z_out + o) = v;
        }
        x[c] = v;
        s += (v.x + v.y) + (v.z + v.w);
      }
    }
    const float mu = group_sum_rt(s, lpr) * invH;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ci = l + c * lpr;
      if (active && ci < nchunk) {
        float a = x[c].x - mu, b = x[c].y - mu, cc = x[c].z - mu, d = x[c].w - mu;
        q += (a * a + b * b) + (cc * cc + d * d);
      }
    }
    const float var = group_sum_rt(q, lpr) * invH;
    const float rs = rsqrtf(var + eps);
    const bool do_ln = active && (cls_T <= 0 || (row % cls_T) == 0);
    if (do_ln) {
      const long long orow = cls_T > 0 ? row / cls_T : row;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const int ci = l + c * lpr;
        if (ci < nchunk) {
          float4 g = *reinterpret_cast<const float4*>(gamma + ci * 4);
          float4 b = *reinterpret_cast<const float4*>(beta + ci * 4);
          float4 y;
          y.x = (x[c].x - mu) * rs * g.x + b.x; y.y = (x[c].y - mu) * rs * g.y + b.y;
          y.z = (x[c].z - mu) * rs * g.z + b.z; y.w = (x[c].w - mu) * rs * g.w + b.w;
          Vec4<T>::st(u + (size_t)orow * H + (size_t)ci * 4, y);
        }
      }
      if (l == 0) { mean[orow] = mu; rstd[orow] = rs; }
    }
  }
}

template <typename T, int MAXC>
__global__ void __launch_bounds__(LN_THREADS)
add_ln_bwd_kernel(const T* __restrict__ du, const float* __restrict__ z, const float* __restrict__ mean,
                  const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ dres,
                  float* __restrict__ dz, T* __restrict__ ddelta, float* __restrict__ dgamma,
                  float* __restrict__ dbeta, int M, int H, int cls_T, float p_drop, const uint64_t* __restrict__ rng,
                  uint32_t site, int accumulate, int lpr, float* __restrict__ partial, unsigned int* counter) {
  __shared__ float red[2][1024];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rpw = 32 / lpr, sub = lane / lpr, l = lane % lpr;
  const int nchunk = H >> 2;
  const DropCtx dc = make_drop(ddelta ? p_drop : 0.f, rng ? rng[0] : 0ull, rng ? (uint32_t)rng[1] : 0u, site);
  const float invH = 1.f / (float)H;
  const long long rows_per_block = (long long)(LN_THREADS / 32) * rpw;
  float4 ag[MAXC], ab[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) { ag[c] = make_float4(0.f, 0.f, 0.f, 0.f); ab[c] = ag[c]; }

  for (long long base = (long long)blockIdx.x * rows_per_block; base < M; base += (long long)gridDim.x * rows_per_block) {
    const long long row = base + (long long)warp * rpw + sub;
    const bool active = row < M;
    const bool has_ln = active && (cls_T <= 0 || (row % cls_T) == 0);
    const long long irow = cls_T > 0 ? row / cls_T : row;
    float mu = 0.f, rs = 0.f;
    if (has_ln) { mu = mean[irow]; rs = rstd[irow]; }
    float4 xh[MAXC], g[MAXC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ci = l + c * lpr;
      xh[c] = make_float4(0.f, 0.f, 0.f, 0.f); g[c] = xh[c];
      if (has_ln && ci < nchunk) {
        float4 xv = *reinterpret_cast<const float4*>(z + (size_t)row * H + (size_t)ci * 4);
        float4 dy = Vec4<T>::ld(du + (size_t)irow * H + (size_t)ci * 4);
        float4 gm = *reinterpret_cast<const float4*>(gamma + ci * 4);
        xh[c].x = (xv.x - mu) * rs; xh[c].y = (xv.y - mu) * rs; xh[c].z = (xv.z - mu) * rs; xh[c].w = (xv.w - mu) * rs;
        ag[c].x += dy.x * xh[c].x; ag[c].y += dy.y * xh[c].y; ag[c].z += dy.z * xh[c].z; ag[c].w += dy.w * xh[c].w;
        ab[c].x += dy.x; ab[c].y += dy.y; ab[c].z += dy.z; ab[c].w += dy.w;
        g[c].x = dy.x * gm.x; g[c].y = dy.y * gm.y; g[c].z = dy.z * gm.z; g[c].w = dy.w * gm.w;
        s1 += (g[c].x + g[c].y) + (g[c].z + g[c].w);
        s2 += (g[c].x * xh[c].x + g[c].y * xh[c].y) + (g[c].z * xh[c].z + g[c].w * xh[c].w);
      }
    }
    const float c1 = group_sum_rt(s1, lpr) * invH;
    const float c2 = group_sum_rt(s2, lpr) * invH;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ci = l + c * lpr;
      if (active && ci < nchunk) {
        const size_t o = (size_t)row * H + (size_t)ci * 4;
        float4 r = dres ? *reinterpret_cast<const float4*>(dres + o) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_ln) {
          r.x += rs * (g[c].x - c1 - xh[c].x * c2); r.y += rs * (g[c].y - c1 - xh[c].y * c2);
          r.z += rs * (g[c].z - c1 - xh[c].z * c2); r.w += rs * (g[c].w - c1 - xh[c].w * c2);
        }
        *reinterpret_cast<float4*>(dz + o) = r;
        if (ddelta) {
          float4 k = drop4(dc, o >> 2);
          // autocast: the gradient reaching the bf16 branch output is cast to bf16, then dropout-backward
          float4 t;
          t.x = round_to<T>(r.x) * k.x; t.y = round_to<T>(r.y) * k.y;
          t.z = round_to<T>(r.z) * k.z; t.w = round_to<T>(r.w) * k.w;
          Vec4<T>::st(ddelta + o, t);
        }
      }
    }
  }

  // ---- dgamma / dbeta: lanes of a warp that hold the same columns, then warps in warp order,
  //      then CTAs in CTA order (all fixed => deterministic) ----
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    for (int o = lpr; o < 32; o <<= 1) {
      ag[c].x += __shfl_xor_sync(0xffffffffu, ag[c].x, o); ag[c].y += __shfl_xor_sync(0xffffffffu, ag[c].y, o);
      ag[c].z += __shfl_xor_sync(0xffffffffu, ag[c].z, o); ag[c].w += __shfl_xor_sync(0xffffffffu, ag[c].w, o);
      ab[c].x += __shfl_xor_sync(0xffffffffu, ab[c].x, o); ab[c].y += __shfl_xor_sync(0xffffffffu, ab[c].y, o);
      ab[c].z += __shfl_xor_sync(0xffffffffu, ab[c].z, o); ab[c].w += __shfl_xor_sync(0xffffffffu, ab[c].w, o);
    }
  }
  for (int i = threadIdx.x; i < H; i += LN_THREADS) { red[0][i] = 0.f; red[1][i] = 0.f; }
  __syncthreads();
  for (int w = 0; w < LN_THREADS / 32; ++w) {
    if (warp == w && sub == 0) {
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const int ci = l + c * lpr;
        if (ci < nchunk) {
          float* r0 = &red[0][ci * 4]; float* r1 = &red[1][ci * 4];
          r0[0] += ag[c].x; r0[1] += ag[c].y; r0[2] += ag[c].z; r0[3] += ag[c].w;
          r1[0] += ab[c].x; r1[1] += ab[c].y; r1[2] += ab[c].z; r1[3] += ab[c].w;
        }
      }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < H; i += LN_THREADS) {
    partial[(size_t)blockIdx.x * 2 * H + i] = red[0][i];
    partial[(size_t)blockIdx.x * 2 * H + H + i] = red[1][i];
  }
  if (!last_block_ticket(counter, gridDim.x)) return;
  // 2H output columns over all threads; 8 partial loads in flight per thread, summed in CTA order
  for (int i = threadIdx.x; i < 2 * H; i += LN_THREADS) {
    float sum = 0.f;
    unsigned int b = 0;
    const unsigned int nb = gridDim.x;
    for (; b + 8 <= nb; b += 8) {
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = __ldcg(&partial[(size_t)(b + q) * 2 * H + i]);
#pragma unroll
      for (int q = 0; q < 8; ++q) sum += v[q];
    }
    for (; b < nb; ++b) sum += __ldcg(&partial[(size_t)b * 2 * H + i]);
    float* dst = i < H ? dgamma : dbeta;
    const int c = i < H ? i : i - H;
    if (dst) dst[c] = accumulate ? dst[c] + sum : sum;
  }
}

static inline int ln_grid(int M, int lpr, int max_blocks = LN_MAX_BLOCKS) {
  long long rows_per_block = (LN_THREADS / 32) * (32 / lpr);
  long long g = (M + rows_per_block - 1) / rows_per_block;
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace vb

using namespace vb;

static int ln_check(int M, int H) {
  if (M < 0 || H <= 0) return VITB200_ERR_ARG;
  if (H % 4 != 0 || H > 1024) return VITB200_ERR_SHAPE;
  return VITB200_OK;
}

extern "C" int vitb200_add_ln_fwd(const float* z_in, const void* delta, float* z_out, void* u, float* mean, float* rstd,
                                  const float* gamma, const float* beta, int M, int H, int cls_T, float eps,
                                  float p_drop, const uint64_t* rng, uint32_t site, int dtype, void* stream) {
  int rc = ln_check(M, H);
  if (rc) return rc;
  if (!z_in || !u || !mean || !rstd || !gamma || !beta || (delta && !z_out)) return VITB200_ERR_ARG;
  if (M == 0) return VITB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int lpr = ln_lpr(H), grid = ln_grid(M, lpr);
  const bool small = (H / 4) <= 32;
#define LAUNCH(T, C)                                                                                              \
  add_ln_fwd_kernel<T, C><<<grid, LN_THREADS, 0, st>>>(z_in, (const T*)delta, z_out, (T*)u, mean, rstd, gamma, beta, \
                                                       M, H, cls_T, eps, p_drop, rng, site, lpr)
  if (dtype == VITB200_F32) { if (small) LAUNCH(float, 1); else LAUNCH(float, 8); }
  else if (dtype == VITB200_BF16) { if (small) LAUNCH(bf16, 1); else LAUNCH(bf16, 8); }
  else return VITB200_ERR_ARG;
#undef LAUNCH
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

extern "C" size_t vitb200_add_ln_bwd_ws_bytes(int M, int H) {
  (void)M;
  return 4096 + (size_t)LN_MAX_BLOCKS * 2 * H * sizeof(float);
}

extern "C" int vitb200_add_ln_bwd(const void* du, const float* z, const float* mean, const float* rstd,
                                  const float* gamma, const float* dres, float* dz, void* ddelta, float* dgamma,
                                  float* dbeta, int M, int H, int cls_T, float p_drop, const uint64_t* rng,
                                  uint32_t site, int accumulate, int dtype, void* ws, void* stream) {
  int rc = ln_check(M, H);
  if (rc) return rc;
  if (!du || !z || !mean || !rstd || !gamma || !dz || !ws) return VITB200_ERR_ARG;
  if (M == 0) return VITB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int lpr = ln_lpr(H), grid = ln_grid(M, lpr, 148);  // one CTA per SM: the second stage reads `grid` partials
  const bool small = (H / 4) <= 32;
  unsigned int* counter = reinterpret_cast<unsigned int*>(ws);
  float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + 4096);
#define LAUNCH(T, C)                                                                                            \
  add_ln_bwd_kernel<T, C><<<grid, LN_THREADS, 0, st>>>((const T*)du, z, mean, rstd, gamma, dres, dz, (T*)ddelta, \
                                                       dgamma, dbeta, M, H, cls_T, p_drop, rng, site, accumulate, \
                                                       lpr, partial, counter)
  if (dtype == VITB200_F32) { if (small) LAUNCH(float, 1); else LAUNCH(float, 8); }
  else if (dtype == VITB200_BF16) { if (small) LAUNCH(bf16, 1); else LAUNCH(bf16, 8); }
  else return VITB200_ERR_ARG;
#undef LAUNCH
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}
