// Kernels for the callers on either side of the encoder step (SURVEY.md 8f "next" rows): the input hand-off from a
// device-resident dataset, the pixel gradient a trainable input preprocessor needs, and the evaluation metrics.
// All of them are HBM-bound byte movers / reductions: coalesced 128-bit accesses, grids sized to cover the 148 SMs.
#include "common.cuh"

namespace vb {

constexpr int PL_THREADS = 256;

// ---------------------------------------------------------------------------------------------
// bf16 -> fp32 (the bf16 output of the preprocessor GEMM feeds the fp32 pixel buffer of the embedding kernel)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PL_THREADS)
cast_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * PL_THREADS + threadIdx.x; i < n4; i += (size_t)gridDim.x * PL_THREADS)
    *reinterpret_cast<float4*>(dst + i * 4) = Vec4<bf16>::ld(src + i * 4);
}

// ---------------------------------------------------------------------------------------------
// d loss / d pixel of the patch embedding (the `unfold` + Linear backward w.r.t. its input):
//   dx[b, l] = sum over windows n covering l (n*S <= l < n*S + P, n < n_valid) of
//              sum_h drop'(dz[b, 1+n, h]) * w[h, l - n*S]
// One thread per pixel.  T = GEMM operand type: in bf16 mode the masked gradient, each window's contribution and the
// result are rounded to bf16 like autocast's bf16 Linear backward.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(PL_THREADS)
patch_embed_dgrad_kernel(const float* __restrict__ dz, const T* __restrict__ w, float* __restrict__ dx, int B, int L,
                         int P, int S, int Np, int n_valid, int H, float p_drop, const uint64_t* __restrict__ rng,
                         uint32_t site) {
  const DropCtx dc = make_drop(p_drop, rng ? rng[0] : 0ull, rng ? (uint32_t)rng[1] : 0u, site);
  const int T1 = Np + 1;
  const size_t total = (size_t)B * L;
  for (size_t i = (size_t)blockIdx.x * PL_THREADS + threadIdx.x; i < total; i += (size_t)gridDim.x * PL_THREADS) {
    const int b = (int)(i / L), l = (int)(i - (size_t)b * L);
    const int n_lo = l >= P ? (l - P) / S + 1 : 0;
    int n_hi = l / S;
    if (n_hi > n_valid - 1) n_hi = n_valid - 1;
    float tot = 0.f;
    for (int n = n_lo; n <= n_hi; ++n) {
      const int j = l - n * S;
      const size_t o = ((size_t)b * T1 + 1 + n) * H;
      float acc = 0.f;
      for (int h = 0; h < H; h += 4) {
        const float4 g = *reinterpret_cast<const float4*>(dz + o + h);
        const float4 kp = drop4(dc, (o + h) >> 2);
        acc = fmaf(round_to<T>(g.x * kp.x), to_f<T>(w[(size_t)(h + 0) * P + j]), acc);
        acc = fmaf(round_to<T>(g.y * kp.y), to_f<T>(w[(size_t)(h + 1) * P + j]), acc);
        acc = fmaf(round_to<T>(g.z * kp.z), to_f<T>(w[(size_t)(h + 2) * P + j]), acc);
        acc = fmaf(round_to<T>(g.w * kp.w), to_f<T>(w[(size_t)(h + 3) * P + j]), acc);
      }
      tot += round_to<T>(acc);
    }
    dx[i] = round_to<T>(tot);
  }
}

// ---------------------------------------------------------------------------------------------
// Batch gather from a device-resident dataset (src/dataloader/base.py:219-245 keeps the whole set in RAM as fp32
// tensors; here it lives in HBM): x[i, :] = flux[idx[i], :] (+ N(0,1) * error[idx[i], :] * noise_level, src/vit.py:86-88)
// and y[i] = labels[idx[i]] (label rows copied as raw bytes: fp32 [C] or int64).  blockIdx.x = sample of the batch,
// blockIdx.y = slice of the row.  Normal deviates: Philox4x32-7 keyed like the dropout masks (site 0x4e5a) + Box-Muller.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float u01(uint32_t r) { return ((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f); }

__global__ void __launch_bounds__(PL_THREADS)
gather_batch_kernel(const float* __restrict__ flux, const float* __restrict__ error, const unsigned char* __restrict__ labels,
                    const int64_t* __restrict__ idx, float* __restrict__ x, unsigned char* __restrict__ y, int L,
                    int label_bytes, long long n_rows, float noise_level, const uint64_t* __restrict__ rng) {
  const int i = blockIdx.x;
  long long r = idx ? idx[i] : (long long)i;
  if (r < 0 || r >= n_rows) r = 0;  // the host validates the index range; never read out of bounds
  const float* src = flux + (size_t)r * L;
  float* dst = x + (size_t)i * L;
  const bool noisy = noise_level > 0.f && error != nullptr;
  const float* err = noisy ? error + (size_t)r * L : nullptr;
  const uint64_t seed = rng ? rng[0] : 0ull;
  const uint32_t step = rng ? (uint32_t)rng[1] : 0u;
  const int L4 = L >> 2;
  const int per = (L4 + gridDim.y - 1) / gridDim.y;
  const int q0 = blockIdx.y * per, q1 = min(L4, q0 + per);
  for (int q = q0 + threadIdx.x; q < q1; q += PL_THREADS) {
    float4 v = *reinterpret_cast<const float4*>(src + q * 4);
    if (noisy) {
      const float4 e = *reinterpret_cast<const float4*>(err + q * 4);
      const uint64_t c = (uint64_t)i * L4 + q;
      const uint4 u = philox4x32_7((uint32_t)c, (uint32_t)(c >> 32), 0x4e5au, step, (uint32_t)seed, (uint32_t)(seed >> 32));
      const float r0 = sqrtf(-2.f * __logf(u01(u.x))), r1 = sqrtf(-2.f * __logf(u01(u.z)));
      float s0, c0, s1, c1;
      __sincosf(6.283185307179586f * u01(u.y), &s0, &c0);
      __sincosf(6.283185307179586f * u01(u.w), &s1, &c1);
      v.x = fmaf(r0 * c0 * noise_level, e.x, v.x);
      v.y = fmaf(r0 * s0 * noise_level, e.y, v.y);
      v.z = fmaf(r1 * c1 * noise_level, e.z, v.z);
      v.w = fmaf(r1 * s1 * noise_level, e.w, v.w);
    }
    *reinterpret_cast<float4*>(dst + q * 4) = v;
  }
  if (blockIdx.y == 0 && labels != nullptr && y != nullptr)
    for (int k = threadIdx.x; k < label_bytes; k += PL_THREADS) y[(size_t)i * label_bytes + k] = labels[(size_t)r * label_bytes + k];
}

// ---------------------------------------------------------------------------------------------
// Evaluation metrics accumulated on the device (src/vit.py:94-125 uses torchmetrics MeanAbsoluteError /
// MeanSquaredError / R2Score / Accuracy, each a host-visible state update per batch).  One CTA, fixed-order reduction
// in double precision (deterministic).  acc (DEVICE, doubles):
//   acc[0] += B ; acc[1] += B * loss (loss may be NULL)
//   regression (labels f32 [B, C]):  acc[2 + 4c + {0,1,2,3}] += sum_b {|e|, e^2, y, y^2} with e = logits - y
//   classification (labels int64 [B]): acc[2] += #(argmax_c logits[b, c] == labels[b])
// ---------------------------------------------------------------------------------------------
constexpr int MT_MAX_C = 16;

__global__ void __launch_bounds__(PL_THREADS)
eval_metrics_kernel(const float* __restrict__ logits, const void* __restrict__ labels, const float* __restrict__ loss,
                    double* __restrict__ acc, int B, int C, int is_cls) {
  __shared__ double sh[PL_THREADS];
  const int nq = is_cls ? 1 : 4 * C;
  for (int q = 0; q < nq; ++q) {
    double s = 0.0;
    if (is_cls) {
      const int64_t* lab = reinterpret_cast<const int64_t*>(labels);
      for (int b = threadIdx.x; b < B; b += PL_THREADS) {
        int best = 0;
        float bv = logits[(size_t)b * C];
        for (int c = 1; c < C; ++c) {
          const float v = logits[(size_t)b * C + c];
          if (v > bv) { bv = v; best = c; }   // first maximum wins, like torch.argmax
        }
        s += (best == (int)lab[b]) ? 1.0 : 0.0;
      }
    } else {
      const float* lab = reinterpret_cast<const float*>(labels);
      const int c = q >> 2, kind = q & 3;
      for (int b = threadIdx.x; b < B; b += PL_THREADS) {
        const double yv = (double)lab[(size_t)b * C + c];
        const double e = (double)logits[(size_t)b * C + c] - yv;
        s += kind == 0 ? fabs(e) : kind == 1 ? e * e : kind == 2 ? yv : yv * yv;
      }
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = PL_THREADS / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) acc[2 + q] += sh[0];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    acc[0] += (double)B;
    if (loss) acc[1] += (double)B * (double)loss[0];
  }
}

static inline int pl_grid(size_t n) {
  size_t g = (n + PL_THREADS - 1) / PL_THREADS;
  return (int)(g > 1184 ? 1184 : (g < 1 ? 1 : g));
}

}  // namespace vb

using namespace vb;

extern "C" int vitb200_cast_f32(const void* src, float* dst, size_t n, void* stream) {
  if (!src || !dst) return VITB200_ERR_ARG;
  if (n % 4 != 0) return VITB200_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(dst) & 15) != 0 || (reinterpret_cast<uintptr_t>(src) & 7) != 0) return VITB200_ERR_ALIGN;
  if (n == 0) return VITB200_OK;
  cast_f32_kernel<<<pl_grid(n / 4), PL_THREADS, 0, (cudaStream_t)stream>>>((const bf16*)src, dst, n / 4);
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

extern "C" int vitb200_patch_embed_dgrad(const float* dz, const void* w, float* dx, int B, int L, int P, int S, int Np,
                                         int n_valid, int H, float p_drop, const uint64_t* rng, uint32_t site, int dtype,
                                         void* stream) {
  if (!dz || !w || !dx) return VITB200_ERR_ARG;
  if (B < 0 || L <= 0 || P <= 0 || S <= 0 || Np <= 0 || n_valid < 0 || n_valid > Np || H <= 0) return VITB200_ERR_ARG;
  if (H % 4 != 0) return VITB200_ERR_SHAPE;
  if (n_valid > 0 && (long long)(n_valid - 1) * S + P > L) return VITB200_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(dz) & 15) != 0) return VITB200_ERR_ALIGN;
  if (B == 0) return VITB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = pl_grid((size_t)B * L);
  if (dtype == VITB200_F32)
    patch_embed_dgrad_kernel<float><<<grid, PL_THREADS, 0, st>>>(dz, (const float*)w, dx, B, L, P, S, Np, n_valid, H,
                                                                 p_drop, rng, site);
  else if (dtype == VITB200_BF16)
    patch_embed_dgrad_kernel<bf16><<<grid, PL_THREADS, 0, st>>>(dz, (const bf16*)w, dx, B, L, P, S, Np, n_valid, H,
                                                                p_drop, rng, site);
  else
    return VITB200_ERR_ARG;
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

extern "C" int vitb200_gather_batch(const float* flux, const float* error, const void* labels, const int64_t* idx,
                                    float* x, void* y, int B, int L, int label_bytes, long long n_rows,
                                    float noise_level, const uint64_t* rng, void* stream) {
  if (!flux || !x || B < 0 || L <= 0 || n_rows <= 0 || label_bytes < 0) return VITB200_ERR_ARG;
  if ((labels == nullptr) != (y == nullptr)) return VITB200_ERR_ARG;
  if (L % 4 != 0) return VITB200_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(flux) & 15) != 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0 ||
      (error && (reinterpret_cast<uintptr_t>(error) & 15) != 0))
    return VITB200_ERR_ALIGN;
  if (B == 0) return VITB200_OK;
  // enough CTAs to cover the 148 SMs about twice, at least 256 float4 per CTA
  int slices = (296 + B - 1) / B;
  const int max_slices = ((L >> 2) + PL_THREADS - 1) / PL_THREADS;
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  dim3 grid(B, slices);
  gather_batch_kernel<<<grid, PL_THREADS, 0, (cudaStream_t)stream>>>(
      flux, error, (const unsigned char*)labels, idx, x, (unsigned char*)y, L, label_bytes, n_rows, noise_level, rng);
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

extern "C" int vitb200_eval_metrics_accum(const float* logits, const void* labels, const float* loss, double* acc, int B,
                                          int C, int is_cls, void* stream) {
  if (!logits || !labels || !acc || B < 0 || C <= 0) return VITB200_ERR_ARG;
  if (!is_cls && C > MT_MAX_C) return VITB200_ERR_SHAPE;
  if (B == 0) return VITB200_OK;
  eval_metrics_kernel<<<1, PL_THREADS, 0, (cudaStream_t)stream>>>(logits, labels, loss, acc, B, C, is_cls ? 1 : 0);
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}
