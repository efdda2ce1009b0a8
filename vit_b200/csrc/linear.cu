// nn.Linear forward / dgrad / wgrad on the SIMT GEMM (FP32 mode, and BF16 mode until the tcgen05
// path takes over for shapes it supports).  See include/vit_b200.h for the contract.
#include "gemm_simt.cuh"

namespace vb {

template <typename T>
struct EpiBiasAct {
  static constexpr bool kSplit = false;
  T* y; T* y_act; const float* bias; int N; int act;
  template <int TN>
  __device__ __forceinline__ void apply(int m, int n0, const float (&v)[TN]) const {
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + j;
      if (n >= N) continue;
      float a = v[j] + (bias ? bias[n] : 0.f);
      size_t o = (size_t)m * N + n;
      y[o] = from_f<T>(a);
      if (act == VITB200_ACT_GELU) y_act[o] = from_f<T>(gelu_f(round_to<T>(a)));
    }
  }
};

template <typename T>
struct EpiDgrad {
  static constexpr bool kSplit = false;
  T* dx; const T* pre; int K;
  template <int TN>
  __device__ __forceinline__ void apply(int m, int n0, const float (&v)[TN]) const {
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + j;
      if (n >= K) continue;
      size_t o = (size_t)m * K + n;
      float g = v[j];
      if (pre) g = round_to<T>(g) * gelu_grad_f(to_f<T>(pre[o]));
      dx[o] = from_f<T>(g);
    }
  }
};

struct EpiWgrad {  // output [N_w, K_w + 1]: column K_w is the bias gradient (ones-column trick)
  static constexpr bool kSplit = true;
  float* dw; float* db; int Kw; int accumulate;
  __device__ __forceinline__ void apply1(int m, int n, float v) const {
    if (n < Kw) {
      size_t o = (size_t)m * Kw + n;
      dw[o] = accumulate ? dw[o] + v : v;
    } else if (n == Kw && db) {
      db[m] = accumulate ? db[m] + v : v;
    }
  }
  template <int TN>
  __device__ __forceinline__ void apply(int m, int n0, const float (&v)[TN]) const {
#pragma unroll
    for (int j = 0; j < TN; ++j) apply1(m, n0 + j, v[j]);
  }
};

template <typename T>
static int linear_fwd_t(const void* x, const void* w, const float* bias, void* y, void* y_act, int M, int N, int K,
                        int act, cudaStream_t st) {
  AccKMajor<T> A{(const T*)x, K, M, K};
  AccKMajor<T> B{(const T*)w, K, N, K};
  EpiBiasAct<T> epi{(T*)y, (T*)y_act, bias, N, act};
  if (N <= 32) return launch_gemm<32>(A, B, epi, M, N, K, 1, nullptr, nullptr, st);
  return launch_gemm<64>(A, B, epi, M, N, K, 1, nullptr, nullptr, st);
}

template <typename T>
static int linear_dgrad_t(const void* dy, const void* w, const void* pre, void* dx, int M, int N, int K,
                          cudaStream_t st) {
  // dx[m,k] = sum_n dy[m,n] w[n,k] : A = dy (contraction n contiguous), B(k, n) = w[n*K + k]
  AccKMajor<T> A{(const T*)dy, N, M, N};
  AccMNMajor<T, false> B{(const T*)w, K, K, N};
  EpiDgrad<T> epi{(T*)dx, (const T*)pre, K};
  if (K <= 32) return launch_gemm<32>(A, B, epi, M, K, N, 1, nullptr, nullptr, st);
  return launch_gemm<64>(A, B, epi, M, K, N, 1, nullptr, nullptr, st);
}

static inline int wgrad_bn(int K) { return (K + 1) <= 32 ? 32 : 64; }

template <typename T>
static int linear_wgrad_t(const void* dy, const void* x, float* dw, float* db, int M, int N, int K, int accumulate,
                          void* ws, cudaStream_t st) {
  // dw[n,k] = sum_m dy[m,n] x[m,k] : A(n, m) = dy[m*N + n], B(k, m) = x[m*K + k] (+ ones column k == K)
  AccMNMajor<T, false> A{(const T*)dy, N, N, M};
  AccMNMajor<T, true> B{(const T*)x, K, K + 1, M};
  EpiWgrad epi{dw, db, K, accumulate};
  int bn = wgrad_bn(K);
  int splits = gemm_splits(N, K + 1, M, bn);
  unsigned int* counters = reinterpret_cast<unsigned int*>(ws);
  float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + 4096);
  if (bn == 32) return launch_gemm<32>(A, B, epi, N, K + 1, M, splits, counters, partial, st);
  return launch_gemm<64>(A, B, epi, N, K + 1, M, splits, counters, partial, st);
}

// ---- wgrad with a SMALL output and a long contraction (the encoder's Linear layers: N x K <= 128 x 128, M = B T rows) ----
// The split-K GEMM above pays one L2 round trip per 16-row k-step and a serial second stage in the last CTA of a tile
// (38 us per call at M = 8256).  Here every CTA takes one contiguous slab of rows, stages dy[slab, N] and
// [x | 1][slab, K + 1] in shared memory with ONE batch of loads and forms the complete [N, K + 1] partial product from
// there (4 x 4 register micro-tiles); a second launch sums the slabs' partials per output element in slab order
// (deterministic) with all loads of a thread in flight.
constexpr int WS_THREADS = 256;
constexpr int WS_MAX_SLABS = 148;

template <typename T>
__global__ void __launch_bounds__(WS_THREADS)
wgrad_slab_kernel(const T* __restrict__ dy, const T* __restrict__ x, float* __restrict__ partial, int M, int N, int K, int rows) {
  extern __shared__ __align__(16) float ws_smem[];
  const int Kp = (K + 4) & ~3;                 // K real columns, one column of ones (bias gradient), zero padding
  float* sdy = ws_smem;                         // [rows][N]
  float* sx = ws_smem + (size_t)rows * N;       // [rows][Kp]
  const int r0 = blockIdx.x * rows, nr = min(rows, M - r0);
  const int tid = threadIdx.x;
  for (int i = tid; i < rows * N; i += WS_THREADS) {
    const int r = i / N;
    sdy[i] = r < nr ? to_f<T>(dy[(size_t)(r0 + r) * N + (i - r * N)]) : 0.f;
  }
  for (int i = tid; i < rows * Kp; i += WS_THREADS) {
    const int r = i / Kp, k = i - r * Kp;
    sx[i] = r < nr ? (k < K ? to_f<T>(x[(size_t)(r0 + r) * K + k]) : (k == K ? 1.f : 0.f)) : 0.f;
  }
  __syncthreads();
  const int tk = Kp / 4, tn = N / 4;
  float* out = partial + (size_t)blockIdx.x * N * Kp;
  for (int t = tid; t < tn * tk; t += WS_THREADS) {
    const int n0 = (t / tk) * 4, k0 = (t % tk) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int r = 0; r < nr; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(&sdy[(size_t)r * N + n0]);
      const float4 b = *reinterpret_cast<const float4*>(&sx[(size_t)r * Kp + k0]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(&out[(size_t)(n0 + i) * Kp + k0]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
}

__global__ void __launch_bounds__(WS_THREADS)
wgrad_slab_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, float* __restrict__ db, int N, int K,
                         int slabs, int accumulate) {
  const int Kp = (K + 4) & ~3;
  const int o = blockIdx.x * WS_THREADS + threadIdx.x;
  if (o >= N * (K + 1)) return;
  const int n = o / (K + 1), k = o - n * (K + 1);
  const float* src = partial + (size_t)n * Kp + k;
  const size_t stride = (size_t)N * Kp;
  float sum = 0.f;
  int z = 0;
  for (; z + 32 <= slabs; z += 32) {   // 32 loads in flight, added in slab order
    float v[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) v[q] = __ldcg(src + (size_t)(z + q) * stride);
#pragma unroll
    for (int q = 0; q < 32; ++q) sum += v[q];
  }
  for (; z < slabs; ++z) sum += __ldcg(src + (size_t)z * stride);
  if (k < K) {
    const size_t i = (size_t)n * K + k;
    dw[i] = accumulate ? dw[i] + sum : sum;
  } else if (db) {
    db[n] = accumulate ? db[n] + sum : sum;
  }
}

static inline bool wgrad_slab_ok(int M, int N, int K) {
  return M >= 1024 && M <= WS_MAX_SLABS * 64 && N % 4 == 0 && N <= 128 && K <= 128 && (size_t)N * (K + 1) <= 8192;
}
static inline void wgrad_slab_shape(int M, int* slabs, int* rows) {
  int r = (M + WS_MAX_SLABS - 1) / WS_MAX_SLABS;
  if (r < 8) r = 8;
  *rows = r;
  *slabs = (M + r - 1) / r;
}
template <typename T>
static int linear_wgrad_slab_t(const void* dy, const void* x, float* dw, float* db, int M, int N, int K, int accumulate,
                               void* ws, cudaStream_t st) {
  int slabs, rows;
  wgrad_slab_shape(M, &slabs, &rows);
  const int Kp = (K + 4) & ~3;
  const size_t smem = (size_t)rows * (N + Kp) * sizeof(float);
  float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + 4096);
  static size_t max_set = 0;
  if (smem > 48 * 1024 && smem > max_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_slab_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return vb_cuda_error(e);
    max_set = smem;
  }
  wgrad_slab_kernel<T><<<slabs, WS_THREADS, smem, st>>>((const T*)dy, (const T*)x, partial, M, N, K, rows);
  VB_CHECK_LAUNCH();
  const int outs = N * (K + 1);
  wgrad_slab_reduce_kernel<<<(outs + WS_THREADS - 1) / WS_THREADS, WS_THREADS, 0, st>>>(partial, dw, db, N, K, slabs, accumulate);
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

}  // namespace vb

using namespace vb;

// tensor-core (tcgen05 / TMA) implementations, gemm_tc.cu
extern "C" int vitb200_tc_supported(int M, int N, int K);
extern "C" int vitb200_tc_linear_fwd(const void*, const void*, const float*, void*, void*, int, int, int, int, void*);
extern "C" int vitb200_tc_linear_dgrad(const void*, const void*, const void*, void*, int, int, int, void*);
extern "C" size_t vitb200_tc_linear_wgrad_ws_bytes(int, int, int);
extern "C" int vitb200_tc_linear_wgrad(const void*, const void*, float*, float*, int, int, int, int, void*, void*);

// 0 = automatic (bf16 GEMMs on tcgen05 whenever the shape allows), 1 = SIMT only (A/B tests, debugging)
static int g_gemm_mode = 0;
extern "C" int vitb200_set_gemm_mode(int mode) {
  int old = g_gemm_mode;
  if (mode == 0 || mode == 1) g_gemm_mode = mode;
  return old;
}
static inline bool use_tc(int dtype, int M, int N, int K, const void* a, const void* b, const void* c) {
  return dtype == VITB200_BF16 && g_gemm_mode == 0 && vitb200_tc_supported(M, N, K) &&
         ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0;
}

extern "C" int vitb200_linear_fwd(const void* x, const void* w, const float* bias, void* y, void* y_act, int M, int N,
                                  int K, int act, int dtype, void* stream) {
  if (!x || !w || !y || M < 0 || N <= 0 || K <= 0) return VITB200_ERR_ARG;
  if (act == VITB200_ACT_GELU && !y_act) return VITB200_ERR_ARG;
  if (M == 0) return VITB200_OK;
  if (use_tc(dtype, M, N, K, x, w, y) && (!y_act || (reinterpret_cast<uintptr_t>(y_act) & 15) == 0))
    return vitb200_tc_linear_fwd(x, w, bias, y, y_act, M, N, K, act, stream);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VITB200_F32) return linear_fwd_t<float>(x, w, bias, y, y_act, M, N, K, act, st);
  if (dtype == VITB200_BF16) return linear_fwd_t<bf16>(x, w, bias, y, y_act, M, N, K, act, st);
  return VITB200_ERR_ARG;
}

extern "C" int vitb200_linear_dgrad(const void* dy, const void* w, const void* pre_act, void* dx, int M, int N, int K,
                                    int dtype, void* stream) {
  if (!dy || !w || !dx || M < 0 || N <= 0 || K <= 0) return VITB200_ERR_ARG;
  if (M == 0) return VITB200_OK;
  if (use_tc(dtype, M, N, K, dy, w, dx) && (!pre_act || (reinterpret_cast<uintptr_t>(pre_act) & 15) == 0))
    return vitb200_tc_linear_dgrad(dy, w, pre_act, dx, M, N, K, stream);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VITB200_F32) return linear_dgrad_t<float>(dy, w, pre_act, dx, M, N, K, st);
  if (dtype == VITB200_BF16) return linear_dgrad_t<bf16>(dy, w, pre_act, dx, M, N, K, st);
  return VITB200_ERR_ARG;
}

extern "C" size_t vitb200_linear_wgrad_ws_bytes(int M, int N, int K) {
  int splits = gemm_splits(N, K + 1, M, wgrad_bn(K));
  // launch_gemm may lower the split count, never raise it
  size_t simt = 4096 + (splits > 1 ? (size_t)splits * N * (K + 1) * sizeof(float) : 0);
  if (wgrad_slab_ok(M, N, K)) {
    int slabs, rows;
    wgrad_slab_shape(M, &slabs, &rows);
    const size_t slab = 4096 + (size_t)slabs * N * ((K + 4) & ~3) * sizeof(float);
    if (slab > simt) simt = slab;
  }
  size_t tcb = (M > 0 && vitb200_tc_supported(M, N, K)) ? vitb200_tc_linear_wgrad_ws_bytes(M, N, K) : 0;
  return simt > tcb ? simt : tcb;
}

extern "C" int vitb200_linear_wgrad(const void* dy, const void* x, float* dw, float* dbias, int M, int N, int K,
                                    int accumulate, int dtype, void* ws, void* stream) {
  if (!dy || !x || !dw || !ws || M < 0 || N <= 0 || K <= 0) return VITB200_ERR_ARG;
  if (M > 0 && use_tc(dtype, M, N, K, dy, x, dw))
    return vitb200_tc_linear_wgrad(dy, x, dw, dbias, M, N, K, accumulate, ws, stream);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VITB200_F32 && wgrad_slab_ok(M, N, K)) return linear_wgrad_slab_t<float>(dy, x, dw, dbias, M, N, K, accumulate, ws, st);
  if (dtype == VITB200_F32) return linear_wgrad_t<float>(dy, x, dw, dbias, M, N, K, accumulate, ws, st);
  if (dtype == VITB200_BF16) return linear_wgrad_t<bf16>(dy, x, dw, dbias, M, N, K, accumulate, ws, st);
  return VITB200_ERR_ARG;
}
