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
  size_t tcb = (M > 0 && vitb200_tc_supported(M, N, K)) ? vitb200_tc_linear_wgrad_ws_bytes(M, N, K) : 0;
  return simt > tcb ? simt : tcb;
}

extern "C" int vitb200_linear_wgrad(const void* dy, const void* x, float* dw, float* dbias, int M, int N, int K,
                                    int accumulate, int dtype, void* ws, void* stream) {
  if (!dy || !x || !dw || !ws || M < 0 || N <= 0 || K <= 0) return VITB200_ERR_ARG;
  if (M > 0 && use_tc(dtype, M, N, K, dy, x, dw))
    return vitb200_tc_linear_wgrad(dy, x, dw, dbias, M, N, K, accumulate, ws, stream);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VITB200_F32) return linear_wgrad_t<float>(dy, x, dw, dbias, M, N, K, accumulate, ws, st);
  if (dtype == VITB200_BF16) return linear_wgrad_t<bf16>(dy, x, dw, dbias, M, N, K, accumulate, ws, st);
  return VITB200_ERR_ARG;
}
