// Patch embedding (sliding-window tokenizer as an implicit-unfold GEMM) with the CLS row, learned
// positions and embedding dropout fused into the epilogue; and its backward.
#include "gemm_simt.cuh"

namespace vb {

template <typename T>
struct EpiPatch {
  float* z; const float* bias; const float* cls; const float* pos; int Np, H; DropCtx dc;
  template <int TN>
  __device__ __forceinline__ void apply(int m, int n0, const float (&v)[TN]) const {
    static_assert(TN == 4, "patch epilogue expects 4 consecutive columns");
    if (n0 >= H) return;  // H % 4 == 0, so a group is entirely in or out
    const int b = m / Np, p = m - b * Np;
    const int T1 = Np + 1;
    {
      const size_t o = ((size_t)b * T1 + 1 + p) * H + n0;
      const float4 kp = drop4(dc, o >> 2);
      const float kpa[4] = {kp.x, kp.y, kp.z, kp.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float t = round_to<T>(v[j] + bias[n0 + j]);
        if (pos) t += pos[(size_t)(1 + p) * H + n0 + j];
        z[o + j] = t * kpa[j];
      }
    }
    if (p == 0) {  // the CLS row of this sample (embedding.py:86-88)
      const size_t o = ((size_t)b * T1) * H + n0;
      const float4 kp = drop4(dc, o >> 2);
      const float kpa[4] = {kp.x, kp.y, kp.z, kp.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float t = cls[n0 + j];
        if (pos) t += pos[n0 + j];
        z[o + j] = t * kpa[j];
      }
    }
  }
};

// dz -> (dropout mask) -> dtok rows (GEMM operand type) + per-chunk partial sums over the batch for
// the CLS token and the learned positions; the last CTA reduces the partials in chunk order.
constexpr int EB_THREADS = 256;
constexpr int EB_MAX_CHUNKS = 64;

template <typename T>
__global__ void __launch_bounds__(EB_THREADS)
embed_bwd_prep_kernel(const float* __restrict__ dz, T* __restrict__ dtok, float* __restrict__ partial,
                      float* __restrict__ dcls, float* __restrict__ dpos, int B, int Np, int H, int per_chunk,
                      float p_drop, const uint64_t* __restrict__ rng, uint32_t site, int accumulate,
                      unsigned int* counter) {
  const int T1 = Np + 1;
  const int rows_red = dpos ? T1 : 1;          // rows whose batch-sum is needed
  const size_t red4 = (size_t)rows_red * H / 4;
  const size_t all4 = (size_t)T1 * H / 4;
  const DropCtx dc = make_drop(p_drop, rng ? rng[0] : 0ull, rng ? (uint32_t)rng[1] : 0u, site);
  const int b0 = blockIdx.x * per_chunk, b1 = min(B, b0 + per_chunk);
  float4* part4 = reinterpret_cast<float4*>(partial) + (size_t)blockIdx.x * red4;
  for (size_t e4 = threadIdx.x; e4 < all4; e4 += EB_THREADS) {
    const size_t e = e4 * 4;
    const int t = (int)(e / H), hcol = (int)(e - (size_t)t * H);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = b0; b < b1; ++b) {
      const size_t o = (size_t)b * T1 * H + e;
      float4 g = *reinterpret_cast<const float4*>(dz + o);
      const float4 kp = drop4(dc, o >> 2);
      g.x *= kp.x; g.y *= kp.y; g.z *= kp.z; g.w *= kp.w;
      acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
      if (t >= 1) Vec4<T>::st(dtok + ((size_t)b * Np + (t - 1)) * H + hcol, g);
    }
    if (e4 < red4) part4[e4] = acc;
  }
  if (!last_block_ticket(counter, gridDim.x)) return;
  for (size_t e = threadIdx.x; e < (size_t)rows_red * H; e += EB_THREADS) {
    float s = 0.f;
    for (unsigned int c = 0; c < gridDim.x; ++c) s += __ldcg(&partial[(size_t)c * red4 * 4 + e]);
    if (e < (size_t)H) dcls[e] = accumulate ? dcls[e] + s : s;
    if (dpos) dpos[e] = accumulate ? dpos[e] + s : s;
  }
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static inline int eb_chunks(int B, int* per_chunk) {
  int pc = ceil_div(B, EB_MAX_CHUNKS);
  if (pc < 1) pc = 1;
  *per_chunk = pc;
  int n = ceil_div(B, pc);
  return n < 1 ? 1 : n;
}

// The DropCtx of an epilogue functor is built on the device from the rng pointer; the functor is
// passed by value, so it carries the pointer and builds the context lazily.
template <typename T>
struct EpiPatchLazy {
  static constexpr bool kSplit = false;
  float* z; const float* bias; const float* cls; const float* pos; int Np, H; float p_drop; const uint64_t* rng;
  uint32_t site;
  template <int TN>
  __device__ __forceinline__ void apply(int m, int n0, const float (&v)[TN]) const {
    EpiPatch<T> e{z, bias, cls, pos, Np, H,
                  make_drop(p_drop, rng ? rng[0] : 0ull, rng ? (uint32_t)rng[1] : 0u, site)};
    e.template apply<TN>(m, n0, v);
  }
};

struct EpiWgradPE {  // output [H, P + 1]: column P is the bias gradient (ones-column trick)
  static constexpr bool kSplit = true;
  float* dw; float* db; int Kw; int accumulate;
  __device__ __forceinline__ void apply1(int m, int n, float v) const {
    if (n < Kw) {
      size_t o = (size_t)m * Kw + n;
      dw[o] = accumulate ? dw[o] + v : v;
    } else if (n == Kw) {
      db[m] = accumulate ? db[m] + v : v;
    }
  }
  template <int TN>
  __device__ __forceinline__ void apply(int m, int n0, const float (&v)[TN]) const {
#pragma unroll
    for (int j = 0; j < TN; ++j) apply1(m, n0 + j, v[j]);
  }
};

static inline int pe_wgrad_bn(int P) { return (P + 1) <= 32 ? 32 : 64; }

template <typename T, bool ROUND>
static int patch_bwd_t(const float* dz, const float* x, float* dw, float* dbias, float* dcls, float* dpos, int B, int L,
                       int P, int S, int Np, int n_valid, int H, float p_drop, const uint64_t* rng, uint32_t site,
                       int accumulate, void* ws, cudaStream_t st) {
  int pc;
  const int chunks = eb_chunks(B, &pc);
  char* base = reinterpret_cast<char*>(ws);
  unsigned int* counter = reinterpret_cast<unsigned int*>(base);
  T* dtok = reinterpret_cast<T*>(base + 4096);
  const size_t dtok_b = align256((size_t)B * Np * H * sizeof(float));
  float* partial = reinterpret_cast<float*>(base + 4096 + dtok_b);
  const size_t part_b = align256((size_t)chunks * (Np + 1) * H * sizeof(float));
  float* gemm_partial = reinterpret_cast<float*>(base + 4096 + dtok_b + part_b);
  const int M = B * Np;
  embed_bwd_prep_kernel<T><<<chunks, EB_THREADS, 0, st>>>(dz, dtok, partial, dcls, dpos, B, Np, H, pc, p_drop, rng,
                                                          site, accumulate, counter);
  VB_CHECK_LAUNCH();
  // dw[h, j] = sum_{b,p} dtok[(b,p), h] * x[b, p*S + j] ; dbias[h] = sum dtok[(b,p), h]
  AccMNMajor<T, false> A{dtok, H, H, M};
  AccUnfoldMN<ROUND> Bm{x, L, S, Np, n_valid, P + 1, M};
  EpiWgradPE epi{dw, dbias, P, accumulate};
  const int bn = pe_wgrad_bn(P);
  const int splits = gemm_splits(H, P + 1, M, bn);
  // tickets live in the first 4096 bytes of ws for every kernel (they run back to back and reset themselves)
  if (bn == 32) return launch_gemm<32>(A, Bm, epi, H, P + 1, M, splits, counter, gemm_partial, st);
  return launch_gemm<64>(A, Bm, epi, H, P + 1, M, splits, counter, gemm_partial, st);
}

}  // namespace vb

using namespace vb;

extern "C" int vitb200_patch_embed_fwd(const float* x, const void* w, const float* bias, const float* cls,
                                       const float* pos, float* z, int B, int L, int P, int S, int Np, int n_valid,
                                       int H, float p_drop, const uint64_t* rng, uint32_t site, int dtype,
                                       void* stream) {
  if (!x || !w || !bias || !cls || !z) return VITB200_ERR_ARG;
  if (B < 0 || L <= 0 || P <= 0 || S <= 0 || Np <= 0 || n_valid < 0 || n_valid > Np || H <= 0) return VITB200_ERR_ARG;
  if (H % 4 != 0) return VITB200_ERR_SHAPE;
  if (n_valid > 0 && (long long)(n_valid - 1) * S + P > L) return VITB200_ERR_ARG;
  if (B == 0) return VITB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int M = B * Np;
  if (dtype == VITB200_F32) {
    AccUnfoldK<false> A{x, L, S, Np, n_valid, M, P};
    AccKMajor<float> Bm{(const float*)w, P, H, P};
    EpiPatchLazy<float> epi{z, bias, cls, pos, Np, H, p_drop, rng, site};
    return launch_gemm<64>(A, Bm, epi, M, H, P, 1, nullptr, nullptr, st);
  }
  if (dtype == VITB200_BF16) {
    AccUnfoldK<true> A{x, L, S, Np, n_valid, M, P};
    AccKMajor<bf16> Bm{(const bf16*)w, P, H, P};
    EpiPatchLazy<bf16> epi{z, bias, cls, pos, Np, H, p_drop, rng, site};
    return launch_gemm<64>(A, Bm, epi, M, H, P, 1, nullptr, nullptr, st);
  }
  return VITB200_ERR_ARG;
}

extern "C" size_t vitb200_patch_embed_bwd_ws_bytes(int B, int Np, int P, int H) {
  int pc;
  int chunks = eb_chunks(B, &pc);
  size_t dtok = align256((size_t)B * Np * H * sizeof(float));
  size_t part = align256((size_t)chunks * (Np + 1) * H * sizeof(float));
  int splits = gemm_splits(H, P + 1, B * Np, pe_wgrad_bn(P));
  size_t gemm = splits > 1 ? (size_t)splits * H * (P + 1) * sizeof(float) : 0;
  return 4096 + dtok + part + gemm;
}

extern "C" int vitb200_patch_embed_bwd(const float* dz, const float* x, float* dw, float* dbias, float* dcls,
                                       float* dpos, int B, int L, int P, int S, int Np, int n_valid, int H,
                                       float p_drop, const uint64_t* rng, uint32_t site, int accumulate, int dtype,
                                       void* ws, void* stream) {
  if (!dz || !x || !dw || !dbias || !dcls || !ws) return VITB200_ERR_ARG;
  if (B < 0 || L <= 0 || P <= 0 || S <= 0 || Np <= 0 || n_valid < 0 || n_valid > Np || H <= 0) return VITB200_ERR_ARG;
  if (H % 4 != 0) return VITB200_ERR_SHAPE;
  if (B == 0) return VITB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VITB200_F32)
    return patch_bwd_t<float, false>(dz, x, dw, dbias, dcls, dpos, B, L, P, S, Np, n_valid, H, p_drop, rng, site,
                                     accumulate, ws, st);
  if (dtype == VITB200_BF16)
    return patch_bwd_t<bf16, true>(dz, x, dw, dbias, dcls, dpos, B, L, P, S, Np, n_valid, H, p_drop, rng, site,
                                   accumulate, ws, st);
  return VITB200_ERR_ARG;
}
