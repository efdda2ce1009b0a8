// Regression / classification head on the CLS rows + loss, forward and backward
// (src/models/specvit.py:78-89).  Tiny problem (B x H x C): plain SIMT, fixed-order reductions.
#include "common.cuh"

namespace vb {

constexpr int HL_THREADS = 256;

// one warp per sample: logits[b, c] = s[b, :] . w[c, :] + bias[c]
template <typename T>
__global__ void __launch_bounds__(HL_THREADS)
head_logits_kernel(const T* __restrict__ s, const T* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ logits, int B, int H, int C) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (HL_THREADS / 32) + (threadIdx.x >> 5);
  if (b >= B) return;
  for (int c = 0; c < C; ++c) {
    float acc = 0.f;
    for (int h = lane; h < H; h += 32) acc = fmaf(to_f<T>(s[(size_t)b * H + h]), to_f<T>(w[(size_t)c * H + h]), acc);
    acc = warp_sum(acc);
    if (lane == 0) logits[(size_t)b * C + c] = round_to<T>(acc + (bias ? bias[c] : 0.f));
  }
}

__device__ __forceinline__ float loss_term(const float* __restrict__ logits, const void* __restrict__ labels, int b,
                                           int C, int kind) {
  if (kind == VITB200_LOSS_CE) {
    const long long y = reinterpret_cast<const long long*>(labels)[b];
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, logits[(size_t)b * C + c]);
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(logits[(size_t)b * C + c] - mx);
    return mx + logf(se) - logits[(size_t)b * C + y];
  }
  const float* yl = reinterpret_cast<const float*>(labels);
  float t = 0.f;
  for (int c = 0; c < C; ++c) {
    float d = logits[(size_t)b * C + c] - yl[(size_t)b * C + c];
    t += kind == VITB200_LOSS_L1 ? fabsf(d) : d * d;
  }
  return t;
}

// single block: thread-strided partial sums, then a fixed-order tree
__global__ void __launch_bounds__(1024)
head_loss_kernel(const float* __restrict__ logits, const void* __restrict__ labels, float* __restrict__ loss, int B,
                 int C, int kind) {
  __shared__ float red[1024];
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += 1024) acc += loss_term(logits, labels, b, C, kind);
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float denom = kind == VITB200_LOSS_CE ? (float)B : (float)B * (float)C;
    loss[0] = red[0] / denom;
  }
}

template <typename T>
__device__ __forceinline__ float dlogit(const float* __restrict__ logits, const void* __restrict__ labels, int b, int c,
                                        int B, int C, int kind, float g) {
  float d;
  if (kind == VITB200_LOSS_GIVEN) {  // `labels` already holds d(objective)/d(logits)
    return round_to<T>(reinterpret_cast<const float*>(labels)[(size_t)b * C + c]);
  }
  if (kind == VITB200_LOSS_CE) {
    const long long y = reinterpret_cast<const long long*>(labels)[b];
    float mx = -INFINITY;
    for (int k = 0; k < C; ++k) mx = fmaxf(mx, logits[(size_t)b * C + k]);
    float se = 0.f;
    for (int k = 0; k < C; ++k) se += expf(logits[(size_t)b * C + k] - mx);
    d = (expf(logits[(size_t)b * C + c] - mx) / se - (c == (int)y ? 1.f : 0.f)) / (float)B;
  } else {
    const float e = logits[(size_t)b * C + c] - reinterpret_cast<const float*>(labels)[(size_t)b * C + c];
    const float n = (float)B * (float)C;
    d = kind == VITB200_LOSS_L1 ? ((e > 0.f) - (e < 0.f)) / n : 2.f * e / n;
  }
  return round_to<T>(d * g);  // autocast: the gradient of the bf16 logits is bf16
}

// ds[b, h] = sum_c dl[b, c] w[c, h]   (one warp per sample)
template <typename T>
__global__ void __launch_bounds__(HL_THREADS)
head_ds_kernel(const T* __restrict__ w, const float* __restrict__ logits, const void* __restrict__ labels,
               const float* __restrict__ gloss, T* __restrict__ ds, int B, int H, int C, int kind) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (HL_THREADS / 32) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float g = gloss ? gloss[0] : 1.f;
  for (int h = lane; h < H; h += 32) {
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(dlogit<T>(logits, labels, b, c, B, C, kind, g), to_f<T>(w[(size_t)c * H + h]), acc);
    ds[(size_t)b * H + h] = from_f<T>(acc);
  }
}

// dw[c, h] = sum_b dl[b, c] s[b, h] ; dbias[c] = sum_b dl[b, c].  One block per class c; column h (and the
// virtual column H for the bias) per thread; NS batch slices reduced in slice order.
template <typename T>
__global__ void __launch_bounds__(HL_THREADS)
head_dw_kernel(const T* __restrict__ s, const float* __restrict__ logits, const void* __restrict__ labels,
               const float* __restrict__ gloss, float* __restrict__ dw, float* __restrict__ dbias, int B, int H, int C,
               int kind, int accumulate) {
  __shared__ float red[HL_THREADS];
  const int c = blockIdx.x;
  const float g = gloss ? gloss[0] : 1.f;
  const int HB = min(H + 1, HL_THREADS);   // columns handled per pass (incl. the bias column)
  const int NS = HL_THREADS / HB;          // batch slices
  const int col_l = threadIdx.x % HB, sl = threadIdx.x / HB;
  for (int h0 = 0; h0 < H + 1; h0 += HB) {
    const int h = h0 + col_l;
    float acc = 0.f;
    if (sl < NS && h <= H) {
      for (int b = sl; b < B; b += NS) {
        const float dl = dlogit<T>(logits, labels, b, c, B, C, kind, g);
        acc = fmaf(dl, h < H ? to_f<T>(s[(size_t)b * H + h]) : 1.f, acc);
      }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    if (sl == 0 && h <= H) {
      float t = 0.f;
      for (int k = 0; k < NS; ++k) t += red[k * HB + col_l];
      if (h < H) {
        float* o = dw + (size_t)c * H + h;
        *o = accumulate ? *o + t : t;
      } else if (dbias) {
        dbias[c] = accumulate ? dbias[c] + t : t;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Single-CTA fused variants (the head is a [B, H] x [H, C] problem: one launch instead of two / three)
// ------------------------------------------------------------------------------------------------
constexpr int HF_THREADS = 1024;  // 32 warps: two samples per warp at B = 64 (the kernel is a chain of dependent loads)
constexpr int HF_WARPS = HF_THREADS / 32;
constexpr int HF_MAXC = 4;

// logits + loss in one kernel: warp w handles samples w, w + 8, ... ; loss terms summed in sample order
template <typename T>
__device__ __forceinline__ void head_fwd_body(const T* __restrict__ s, const T* __restrict__ w, const float* __restrict__ bias,
                                              const void* __restrict__ labels, float* __restrict__ logits,
                                              float* __restrict__ loss, int B, int H, int C, int kind) {
  __shared__ float wsum[HF_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc_loss = 0.f;
  for (int b = warp; b < B; b += HF_WARPS) {
    for (int c = 0; c < C; ++c) {
      float a = 0.f;
      for (int h = lane; h < H; h += 32) a = fmaf(to_f<T>(s[(size_t)b * H + h]), to_f<T>(w[(size_t)c * H + h]), a);
      a = warp_sum(a);
      if (lane == 0) logits[(size_t)b * C + c] = round_to<T>(a + (bias ? bias[c] : 0.f));
    }
    __syncwarp();
    if (labels && lane == 0) acc_loss += loss_term(logits, labels, b, C, kind);
  }
  if (!labels) return;
  if (lane == 0) wsum[warp] = acc_loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < HF_WARPS; ++k) t += wsum[k];
    loss[0] = t / (kind == VITB200_LOSS_CE ? (float)B : (float)B * (float)C);
  }
}
template <typename T>
__global__ void __launch_bounds__(HF_THREADS)
head_fused_fwd_kernel(const T* __restrict__ s, const T* __restrict__ w, const float* __restrict__ bias,
                      const void* __restrict__ labels, float* __restrict__ logits, float* __restrict__ loss, int B, int H,
                      int C, int kind) {
  pdl_wait();
  pdl_trigger();
  head_fwd_body<T>(s, w, bias, labels, logits, loss, B, H, C, kind);
}

// head backward + final-LayerNorm backward of the CLS rows in one kernel (C <= HF_MAXC):
//   dl = dloss/dlogits ; ds = dl . W ; dz_cls = LN'(ds) ; dgamma, dbeta, dW, dbias reduced over samples in a
//   fixed order (warp-strided sample order, then warps in order).
template <typename T>
__device__ __forceinline__ void head_bwd_body(const T* __restrict__ s, const T* __restrict__ w, const float* __restrict__ logits,
                      const void* __restrict__ labels, const float* __restrict__ gloss, const float* __restrict__ z,
                      size_t z_row_stride, const float* __restrict__ mean, const float* __restrict__ rstd,
                      const float* __restrict__ gamma, float* __restrict__ dz_cls, float* __restrict__ dgamma,
                      float* __restrict__ dbeta, float* __restrict__ dw, float* __restrict__ dbias, int B, int H, int C,
                      int kind, int accumulate) {
  extern __shared__ float red[];  // [HF_WARPS][(2 + C) * H + C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float g = gloss ? gloss[0] : 1.f;
  constexpr int HPL = 4;   // columns per lane (H <= 128)
  const int npl = (H + 31) / 32;
  float ag[HPL], ab[HPL];      // dgamma / dbeta partials for columns lane, lane+32, ...
  float aw[HF_MAXC][4];        // dW partials: only H <= 128 uses registers; larger H goes through the slow path below
  float abias[HF_MAXC];
#pragma unroll
  for (int i = 0; i < HPL; ++i) { ag[i] = 0.f; ab[i] = 0.f; }
#pragma unroll
  for (int c = 0; c < HF_MAXC; ++c) { abias[c] = 0.f; for (int i = 0; i < 4; ++i) aw[c][i] = 0.f; }
  for (int b = warp; b < B; b += HF_WARPS) {
    float dl[HF_MAXC];
#pragma unroll
    for (int c = 0; c < HF_MAXC; ++c) dl[c] = c < C ? dlogit<T>(logits, labels, b, c, B, C, kind, g) : 0.f;
    const float mu = mean[b], rs = rstd[b];
    float s1 = 0.f, s2 = 0.f;
    float dsv[4], xh[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int h = lane + 32 * i;
      dsv[i] = 0.f; xh[i] = 0.f;
      if (i < npl && h < H) {
        float a = 0.f;
#pragma unroll
        for (int c = 0; c < HF_MAXC; ++c) if (c < C) a = fmaf(dl[c], to_f<T>(w[(size_t)c * H + h]), a);
        a = round_to<T>(a);
        dsv[i] = a;
        xh[i] = (z[(size_t)b * z_row_stride + h] - mu) * rs;
        const float gg = a * gamma[h];
        s1 += gg;
        s2 = fmaf(gg, xh[i], s2);
        ag[i] = fmaf(a, xh[i], ag[i]);
        ab[i] += a;
        const float sv = to_f<T>(s[(size_t)b * H + h]);
#pragma unroll
        for (int c = 0; c < HF_MAXC; ++c) aw[c][i] = fmaf(dl[c], sv, aw[c][i]);
      }
    }
#pragma unroll
    for (int c = 0; c < HF_MAXC; ++c) abias[c] += dl[c];
    const float c1 = warp_sum(s1) / (float)H, c2 = warp_sum(s2) / (float)H;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int h = lane + 32 * i;
      if (i < npl && h < H) dz_cls[(size_t)b * H + h] = rs * (dsv[i] * gamma[h] - c1 - xh[i] * c2);
    }
  }
  // cross-warp reduction in warp order
  const int per = (2 + C) * H + C;
  float* mine = red + (size_t)warp * per;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int h = lane + 32 * i;
    if (i < npl && h < H) {
      mine[h] = ag[i]; mine[H + h] = ab[i];
      for (int c = 0; c < C; ++c) mine[(2 + c) * H + h] = aw[c][i];
    }
  }
  if (lane == 0) for (int c = 0; c < C; ++c) mine[(2 + C) * H + c] = abias[c];
  __syncthreads();
  for (int e = threadIdx.x; e < per; e += HF_THREADS) {
    float t = 0.f;
    for (int k = 0; k < HF_WARPS; ++k) t += red[(size_t)k * per + e];
    float* dst;
    if (e < H) dst = dgamma + e;
    else if (e < 2 * H) dst = dbeta + (e - H);
    else if (e < (2 + C) * H) dst = dw + (e - 2 * H);
    else dst = dbias ? dbias + (e - (2 + C) * H) : nullptr;
    if (dst) *dst = accumulate ? *dst + t : t;
  }
}
template <typename T>
__global__ void __launch_bounds__(HF_THREADS)
head_fused_bwd_kernel(const T* __restrict__ s, const T* __restrict__ w, const float* __restrict__ logits,
                      const void* __restrict__ labels, const float* __restrict__ gloss, const float* __restrict__ z,
                      size_t z_row_stride, const float* __restrict__ mean, const float* __restrict__ rstd,
                      const float* __restrict__ gamma, float* __restrict__ dz_cls, float* __restrict__ dgamma,
                      float* __restrict__ dbeta, float* __restrict__ dw, float* __restrict__ dbias, int B, int H, int C,
                      int kind, int accumulate) {
  pdl_wait();
  pdl_trigger();
  head_bwd_body<T>(s, w, logits, labels, gloss, z, z_row_stride, mean, rstd, gamma, dz_cls, dgamma, dbeta, dw, dbias, B, H, C,
                   kind, accumulate);
}
// forward (logits, loss) and backward of the head in one launch: a training step knows dloss/dloss = 1 in advance.
// Sample b is handled by the same warp in both halves, so the logits only need warp-level ordering.
template <typename T>
__global__ void __launch_bounds__(HF_THREADS)
head_fused_fwd_bwd_kernel(const T* __restrict__ s, const T* __restrict__ w, const float* __restrict__ bias,
                          const void* __restrict__ labels, float* __restrict__ logits, float* __restrict__ loss,
                          const float* __restrict__ z, size_t z_row_stride, const float* __restrict__ mean,
                          const float* __restrict__ rstd, const float* __restrict__ gamma, float* __restrict__ dz_cls,
                          float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dw,
                          float* __restrict__ dbias, int B, int H, int C, int kind) {
  pdl_wait();
  pdl_trigger();
  head_fwd_body<T>(s, w, bias, labels, logits, loss, B, H, C, kind);
  __syncthreads();
  head_bwd_body<T>(s, w, logits, labels, nullptr, z, z_row_stride, mean, rstd, gamma, dz_cls, dgamma, dbeta, dw, dbias, B, H, C,
                   kind, 0);
}

}  // namespace vb

using namespace vb;

extern "C" int vitb200_head_fused_supported(int H, int C) { return (H <= 128 && C <= HF_MAXC) ? 1 : 0; }

extern "C" int vitb200_head_fused_fwd(const void* s, const void* w, const float* bias, const void* labels, float* logits,
                                      float* loss, int B, int H, int C, int loss_kind, int dtype, void* stream) {
  if (!s || !w || !logits || B <= 0 || H <= 0 || C <= 0 || (labels && !loss)) return VITB200_ERR_ARG;
  if (loss_kind < 0 || loss_kind > 2) return VITB200_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VITB200_F32)
    vb_launch_pdl(head_fused_fwd_kernel<float>, dim3(1), dim3(HF_THREADS), 0, st, (const float*)s, (const float*)w, bias, labels, logits, loss, B, H, C, loss_kind);
  else if (dtype == VITB200_BF16)
    vb_launch_pdl(head_fused_fwd_kernel<bf16>, dim3(1), dim3(HF_THREADS), 0, st, (const bf16*)s, (const bf16*)w, bias, labels, logits, loss, B, H, C, loss_kind);
  else
    return VITB200_ERR_ARG;
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

extern "C" int vitb200_head_fused_fwd_bwd(const void* s, const void* w, const float* bias, const void* labels, float* logits,
                                          float* loss, const float* z, size_t z_row_stride, const float* mean,
                                          const float* rstd, const float* gamma, float* dz_cls, float* dgamma, float* dbeta,
                                          float* dw, float* dbias, int B, int H, int C, int loss_kind, int dtype, void* stream) {
  if (!s || !w || !logits || !labels || !loss || !z || !mean || !rstd || !gamma || !dz_cls || !dgamma || !dbeta || !dw)
    return VITB200_ERR_ARG;
  if (B <= 0 || !vitb200_head_fused_supported(H, C)) return VITB200_ERR_SHAPE;
  if (loss_kind < 0 || loss_kind > 2) return VITB200_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)HF_WARPS * ((2 + C) * H + C) * sizeof(float);
  if (dtype == VITB200_F32)
    vb_launch_pdl(head_fused_fwd_bwd_kernel<float>, dim3(1), dim3(HF_THREADS), smem, st, (const float*)s, (const float*)w, bias,
                  labels, logits, loss, z, z_row_stride, mean, rstd, gamma, dz_cls, dgamma, dbeta, dw, dbias, B, H, C, loss_kind);
  else if (dtype == VITB200_BF16)
    vb_launch_pdl(head_fused_fwd_bwd_kernel<bf16>, dim3(1), dim3(HF_THREADS), smem, st, (const bf16*)s, (const bf16*)w, bias,
                  labels, logits, loss, z, z_row_stride, mean, rstd, gamma, dz_cls, dgamma, dbeta, dw, dbias, B, H, C, loss_kind);
  else
    return VITB200_ERR_ARG;
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

extern "C" int vitb200_head_fused_bwd(const void* s, const void* w, const float* logits, const void* labels,
                                      const float* gloss, const float* z, size_t z_row_stride, const float* mean,
                                      const float* rstd, const float* gamma, float* dz_cls, float* dgamma, float* dbeta,
                                      float* dw, float* dbias, int B, int H, int C, int loss_kind, int accumulate,
                                      int dtype, void* stream) {
  if (!s || !w || !logits || !labels || !z || !mean || !rstd || !gamma || !dz_cls || !dgamma || !dbeta || !dw)
    return VITB200_ERR_ARG;
  if (B <= 0 || !vitb200_head_fused_supported(H, C)) return VITB200_ERR_SHAPE;
  if (loss_kind < 0 || loss_kind > 3) return VITB200_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)HF_WARPS * ((2 + C) * H + C) * sizeof(float);
  if (dtype == VITB200_F32)
    vb_launch_pdl(head_fused_bwd_kernel<float>, dim3(1), dim3(HF_THREADS), smem, st, (const float*)s, (const float*)w, logits,
                  labels, gloss, z, z_row_stride, mean, rstd, gamma, dz_cls, dgamma, dbeta, dw, dbias, B, H, C, loss_kind,
                  accumulate);
  else if (dtype == VITB200_BF16)
    vb_launch_pdl(head_fused_bwd_kernel<bf16>, dim3(1), dim3(HF_THREADS), smem, st, (const bf16*)s, (const bf16*)w, logits,
                  labels, gloss, z, z_row_stride, mean, rstd, gamma, dz_cls, dgamma, dbeta, dw, dbias, B, H, C, loss_kind,
                  accumulate);
  else
    return VITB200_ERR_ARG;
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

extern "C" int vitb200_head_loss_fwd(const void* s, const void* w, const float* bias, const void* labels, float* logits,
                                     float* loss, int B, int H, int C, int loss_kind, int dtype, void* stream) {
  if (!s || !w || !logits || B < 0 || H <= 0 || C <= 0) return VITB200_ERR_ARG;
  if (labels && !loss) return VITB200_ERR_ARG;
  if (loss_kind < 0 || loss_kind > 2) return VITB200_ERR_ARG;
  if (B == 0) return VITB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ceil_div(B, HL_THREADS / 32);
  if (dtype == VITB200_F32)
    head_logits_kernel<float><<<grid, HL_THREADS, 0, st>>>((const float*)s, (const float*)w, bias, logits, B, H, C);
  else if (dtype == VITB200_BF16)
    head_logits_kernel<bf16><<<grid, HL_THREADS, 0, st>>>((const bf16*)s, (const bf16*)w, bias, logits, B, H, C);
  else
    return VITB200_ERR_ARG;
  VB_CHECK_LAUNCH();
  if (labels) {
    head_loss_kernel<<<1, 1024, 0, st>>>(logits, labels, loss, B, C, loss_kind);
    VB_CHECK_LAUNCH();
  }
  return VITB200_OK;
}

extern "C" int vitb200_head_loss_bwd(const void* s, const void* w, const float* logits, const void* labels,
                                     const float* gloss, void* ds, float* dw, float* dbias, int B, int H, int C,
                                     int loss_kind, int accumulate, int dtype, void* stream) {
  if (!s || !w || !logits || !labels || !ds || !dw || B < 0 || H <= 0 || C <= 0) return VITB200_ERR_ARG;
  if (loss_kind < 0 || loss_kind > 3) return VITB200_ERR_ARG;
  if (B == 0) return VITB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ceil_div(B, HL_THREADS / 32);
  if (dtype == VITB200_F32) {
    head_ds_kernel<float><<<grid, HL_THREADS, 0, st>>>((const float*)w, logits, labels, gloss, (float*)ds, B, H, C, loss_kind);
    head_dw_kernel<float><<<C, HL_THREADS, 0, st>>>((const float*)s, logits, labels, gloss, dw, dbias, B, H, C, loss_kind, accumulate);
  } else if (dtype == VITB200_BF16) {
    head_ds_kernel<bf16><<<grid, HL_THREADS, 0, st>>>((const bf16*)w, logits, labels, gloss, (bf16*)ds, B, H, C, loss_kind);
    head_dw_kernel<bf16><<<C, HL_THREADS, 0, st>>>((const bf16*)s, logits, labels, gloss, dw, dbias, B, H, C, loss_kind, accumulate);
  } else {
    return VITB200_ERR_ARG;
  }
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}
