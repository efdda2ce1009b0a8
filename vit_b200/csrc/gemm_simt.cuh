// Generic fp32-accumulate SIMT GEMM used (a) for FP32 mode, where the 1e-4 parity bar rules out
// TF32/bf16 tensor cores, and (b) as the trusted comparator for the tcgen05 kernels.
//   C[m,n] = sum_k A(m,k) * B(n,k)       A, B given by accessor objects, C consumed by an epilogue
// 64 x BN x 16 tiles, 256 threads, 4 x (BN/16) register micro-tile, optional split-K with a
// deterministic (fixed-order) second stage done by the last CTA of each output tile.
#pragma once
#include "common.cuh"

namespace vb {

constexpr int GBM = 64;
constexpr int GBK = 16;
constexpr int GNT = 256;
constexpr int GPAD = 4;

// ---- operand accessors ------------------------------------------------------------------------
// kContigK == true : ld4(r,k) returns elements (r, k..k+3)   (4 consecutive k are contiguous)
// kContigK == false: ld4(r,k) returns elements (r..r+3, k)   (4 consecutive rows are contiguous)
// Out-of-range elements read as 0.
template <typename T>
struct AccKMajor {  // element (r,k) = p[r*ld + k]
  static constexpr bool kContigK = true;
  const T* p; int ld; int R; int K;
  __device__ __forceinline__ float4 ld4(int r, int k) const {
    if (r >= R) return make_float4(0.f, 0.f, 0.f, 0.f);
    const T* q = p + (size_t)r * ld + k;
    if (k + 3 < K && aligned_vec4(q)) return Vec4<T>::ld(q);
    float4 v;
    v.x = k + 0 < K ? to_f<T>(q[0]) : 0.f;
    v.y = k + 1 < K ? to_f<T>(q[1]) : 0.f;
    v.z = k + 2 < K ? to_f<T>(q[2]) : 0.f;
    v.w = k + 3 < K ? to_f<T>(q[3]) : 0.f;
    return v;
  }
};
template <typename T, bool ONES>
struct AccMNMajor {  // element (r,k) = p[k*ld + r] ; with ONES, row r == R-1 is a virtual column of ones
  static constexpr bool kContigK = false;
  const T* p; int ld; int R; int K;
  __device__ __forceinline__ float at(int r, int k) const {
    if (r >= R || k >= K) return 0.f;
    if (ONES && r == R - 1) return 1.f;
    return to_f<T>(p[(size_t)k * ld + r]);
  }
  __device__ __forceinline__ float4 ld4(int r, int k) const {
    if (k >= K) return make_float4(0.f, 0.f, 0.f, 0.f);
    const int Rv = ONES ? R - 1 : R;
    const T* q = p + (size_t)k * ld + r;
    if (r + 3 < Rv && aligned_vec4(q)) return Vec4<T>::ld(q);
    return make_float4(at(r, k), at(r + 1, k), at(r + 2, k), at(r + 3, k));
  }
};
// Sliding windows of a [B, L] fp32 signal (tokenization.py:45): row m = b*Np + p, element k = x[b, p*S + k];
// windows p >= n_valid are all-zero (tokenization.py:46-48).  ROUND: round to bf16 first (autocast).
template <bool ROUND>
struct AccUnfoldK {
  static constexpr bool kContigK = true;
  const float* x; int L, S, Np, n_valid, R, K;
  __device__ __forceinline__ float4 ld4(int r, int k) const {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= R) return v;
    int b = r / Np, pp = r - b * Np;
    if (pp >= n_valid) return v;
    const float* q = x + (size_t)b * L + (size_t)pp * S + k;
    if (k + 3 < K && aligned16(q)) {
      v = *reinterpret_cast<const float4*>(q);
    } else {
      v.x = k + 0 < K ? q[0] : 0.f; v.y = k + 1 < K ? q[1] : 0.f;
      v.z = k + 2 < K ? q[2] : 0.f; v.w = k + 3 < K ? q[3] : 0.f;
    }
    if (ROUND) { v.x = bf16_round(v.x); v.y = bf16_round(v.y); v.z = bf16_round(v.z); v.w = bf16_round(v.w); }
    return v;
  }
};
// The same windows as the B operand of the patch-projection wgrad: element (r = j in [0,P], k = row m);
// r == P is the virtual ones column that yields the bias gradient.
template <bool ROUND>
struct AccUnfoldMN {
  static constexpr bool kContigK = false;
  const float* x; int L, S, Np, n_valid, R /* = P+1 */, K /* = B*Np */;
  __device__ __forceinline__ float at(int r, int k) const {
    if (r >= R || k >= K) return 0.f;
    if (r == R - 1) return 1.f;
    int b = k / Np, pp = k - b * Np;
    if (pp >= n_valid) return 0.f;
    float v = x[(size_t)b * L + (size_t)pp * S + r];
    return ROUND ? bf16_round(v) : v;
  }
  __device__ __forceinline__ float4 ld4(int r, int k) const {
    return make_float4(at(r, k), at(r + 1, k), at(r + 2, k), at(r + 3, k));
  }
};

// A tile moves global -> registers -> shared memory in two steps, so that the loads of tile k + 1 fly while tile k is
// multiplied (the GEMMs of FP32 mode are small and latency-bound: without this every k-step paid a full L2 round trip).
template <class Acc, int ROWS>
struct TileRegs { float4 v[(ROWS * GBK / 4 + GNT - 1) / GNT]; };

template <class Acc, int ROWS>
__device__ __forceinline__ void load_tile(TileRegs<Acc, ROWS>& t, const Acc& acc, int r0, int k0, int tid) {
  constexpr int N4 = ROWS * GBK / 4;
#pragma unroll
  for (int j = 0; j < (N4 + GNT - 1) / GNT; ++j) {
    const int i = tid + j * GNT;
    if (i < N4) {
      if (Acc::kContigK) {
        const int r = i / (GBK / 4), kq = (i % (GBK / 4)) * 4;
        t.v[j] = acc.ld4(r0 + r, k0 + kq);
      } else {
        const int kk = i / (ROWS / 4), rq = (i % (ROWS / 4)) * 4;
        t.v[j] = acc.ld4(r0 + rq, k0 + kk);
      }
    }
  }
}
template <class Acc, int ROWS>
__device__ __forceinline__ void store_tile(float (*S)[ROWS + GPAD], const TileRegs<Acc, ROWS>& t, int tid) {
  constexpr int N4 = ROWS * GBK / 4;
#pragma unroll
  for (int j = 0; j < (N4 + GNT - 1) / GNT; ++j) {
    const int i = tid + j * GNT;
    if (i < N4) {
      const float4 v = t.v[j];
      if (Acc::kContigK) {
        const int r = i / (GBK / 4), kq = (i % (GBK / 4)) * 4;
        S[kq + 0][r] = v.x; S[kq + 1][r] = v.y; S[kq + 2][r] = v.z; S[kq + 3][r] = v.w;
      } else {
        const int kk = i / (ROWS / 4), rq = (i % (ROWS / 4)) * 4;
        *reinterpret_cast<float4*>(&S[kk][rq]) = v;
      }
    }
  }
}

// Epilogue concept:  template<int TN> __device__ void apply(int m, int n0, const float (&v)[TN]) const;
// (m < M guaranteed; the functor bounds-checks n0 + j < N itself)
template <int BN, class AAcc, class BAcc, class Epi>
__global__ void __launch_bounds__(GNT) gemm_simt_kernel(AAcc A, BAcc B, Epi epi, int M, int N, int K, int k_chunk,
                                                        float* __restrict__ partial, unsigned int* counters) {
  constexpr int TM = GBM / 16, TN = BN / 16;
  __shared__ __align__(16) float As[GBK][GBM + GPAD];
  __shared__ __align__(16) float Bs[GBK][BN + GPAD];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * BN;
  const int kb = blockIdx.z * k_chunk, ke = min(K, kb + k_chunk);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // accessors zero-fill beyond K; a split's tail beyond ke is excluded because k_chunk % GBK == 0
  TileRegs<AAcc, GBM> ta;
  TileRegs<BAcc, BN> tb;
  if (kb < ke) { load_tile<AAcc, GBM>(ta, A, m0, kb, tid); load_tile<BAcc, BN>(tb, B, n0, kb, tid); }
  for (int k0 = kb; k0 < ke; k0 += GBK) {
    store_tile<AAcc, GBM>(As, ta, tid);
    store_tile<BAcc, BN>(Bs, tb, tid);
    __syncthreads();
    if (k0 + GBK < ke) { load_tile<AAcc, GBM>(ta, A, m0, k0 + GBK, tid); load_tile<BAcc, BN>(tb, B, n0, k0 + GBK, tid); }
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      float a[TM], b[TN];
      float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
      a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
      if (TN == 4) {
        float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN]);
        b[0] = bv.x; b[1] = bv.y; b[TN - 2] = bv.z; b[TN - 1] = bv.w;
      } else {
        float2 bv = *reinterpret_cast<const float2*>(&Bs[kk][tx * TN]);
        b[0] = bv.x; b[1] = bv.y;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  if constexpr (Epi::kSplit) {
    if (gridDim.z > 1) {
      // stage 1: every split writes its partial tile; stage 2: the last CTA of this tile sums all
      // splits in split order (fixed order => bitwise deterministic) and runs the epilogue.  The tile's
      // valid elements are spread over all 256 threads and the loads of 8 splits are issued together.
      const size_t MN = (size_t)M * N;
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * TM + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          int n = n0 + tx * TN + j;
          if (n < N) partial[blockIdx.z * MN + (size_t)m * N + n] = acc[i][j];
        }
      }
      if (!last_block_ticket(&counters[blockIdx.y * gridDim.x + blockIdx.x], gridDim.z)) return;
      const int tm = min(GBM, M - m0), tn = min(BN, N - n0);
      const unsigned int Z = gridDim.z;
      for (int e = tid; e < tm * tn; e += GNT) {
        const int m = m0 + e / tn, n = n0 + e % tn;
        const float* src = partial + (size_t)m * N + n;
        float sum = 0.f;
        unsigned int z = 0;
        for (; z + 32 <= Z; z += 32) {   // 32 loads in flight, added in split order
          float v[32];
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = __ldcg(src + (size_t)(z + q) * MN);
#pragma unroll
          for (int q = 0; q < 32; ++q) sum += v[q];
        }
        for (; z + 8 <= Z; z += 8) {
          float v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = __ldcg(src + (size_t)(z + q) * MN);
#pragma unroll
          for (int q = 0; q < 8; ++q) sum += v[q];
        }
        for (; z < Z; ++z) sum += __ldcg(src + (size_t)z * MN);
        epi.apply1(m, n, sum);
      }
      return;
    }
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m < M) epi.template apply<TN>(m, n0 + tx * TN, acc[i]);
  }
}

// number of K splits for a reduction-heavy (wgrad) GEMM with a small output
static inline int gemm_splits(int M, int N, int K, int BN) {
  long long tiles = (long long)ceil_div(M, GBM) * ceil_div(N, BN);
  if (tiles >= 296 || tiles > 1024) return 1;
  int want = (int)((296 + tiles - 1) / tiles);
  int max_by_k = ceil_div(K, 8 * GBK);
  int s = want < max_by_k ? want : max_by_k;
  if (s > 64) s = 64;  // the second stage is one CTA per tile: keep its serial part short
  if (s < 1) s = 1;
  return s;
}

template <int BN, class AAcc, class BAcc, class Epi>
static inline int launch_gemm(const AAcc& A, const BAcc& B, const Epi& epi, int M, int N, int K, int splits,
                              unsigned int* counters, float* partial, cudaStream_t st) {
  if (M <= 0 || N <= 0) return VITB200_OK;
  int k_chunk = ceil_div(ceil_div(K, splits), GBK) * GBK;
  if (k_chunk <= 0) k_chunk = GBK;
  splits = ceil_div(K, k_chunk);
  if (splits < 1) splits = 1;
  dim3 grid(ceil_div(N, BN), ceil_div(M, GBM), splits);
  // `counters`: >= 1024 zero-initialised uints (self-resetting tickets, shared by all kernels that run
  // one after another on a stream); `partial`: splits * M * N floats of scratch.
  if (splits > 1 && (counters == nullptr || partial == nullptr)) return VITB200_ERR_ARG;
  gemm_simt_kernel<BN, AAcc, BAcc, Epi><<<grid, GNT, 0, st>>>(A, B, epi, M, N, K, k_chunk, partial, counters);
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

}  // namespace vb
