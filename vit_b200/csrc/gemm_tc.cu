// bf16 GEMMs on the 5th-generation tensor cores: TMA-staged 128B-swizzled tiles in shared memory,
// tcgen05.mma issued by one thread, fp32 accumulator in TMEM, epilogue warps read it back with
// tcgen05.ld and fuse bias / erf-GELU / gelu' / split-K partial output.
//
//   D[m, n] = sum_k A(m, k) * B(n, k)            128 x BN output tile per CTA, BLOCK_K = 64
//
// Operands come straight from the row-major activations / nn.Linear weights, no transposed copies:
//   K-major  operand: matrix [rows, k] with k contiguous   -> TMA box {64 k, rows}, UMMA major K
//   MN-major operand: matrix [k, rows] with rows contiguous -> TMA boxes {64 rows, 64 k}, UMMA major MN
//   forward  y  = x . w^T         A = x  (K-major)   B = w  (K-major)
//   dgrad    dx = dy . w          A = dy (K-major)   B = w  (MN-major: w[n][k], k contiguous)
//   wgrad    dw = dy^T . x        A = dy (MN-major)  B = x  (MN-major)
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (warp w owns TMEM lanes 32*(w%4) .. +31).
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_host.cuh"

namespace vb {
using namespace vb::tc;

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_THREADS = 192;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;  // 16 KB per stage

template <int BN> struct TcCfg {
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  static constexpr int STAGES = 4;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;  // power of two for BN in {64,128,256}
};

// ---- epilogues: called by one thread per output row with 32 consecutive columns ----
struct TcEpiBiasAct {
  static constexpr bool kSplit = false;
  bf16* y; bf16* y_act; const float* bias; int act;
  __device__ __forceinline__ void row(int m, int n0, const float (&v)[32], int M, int N) const {
    if (m >= M || n0 >= N) return;
    const size_t o = (size_t)m * N + n0;
    const bool vec = (n0 + 32 <= N) && ((N & 7) == 0);
    float a[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) a[j] = v[j] + ((bias && n0 + j < N) ? bias[n0 + j] : 0.f);
    if (vec) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 pk;
        __nv_bfloat162 t0 = __floats2bfloat162_rn(a[j], a[j + 1]), t1 = __floats2bfloat162_rn(a[j + 2], a[j + 3]);
        __nv_bfloat162 t2 = __floats2bfloat162_rn(a[j + 4], a[j + 5]), t3 = __floats2bfloat162_rn(a[j + 6], a[j + 7]);
        pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
        pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
        *reinterpret_cast<uint4*>(y + o + j) = pk;
        if (act == VITB200_ACT_GELU) {
          float g[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) g[q] = gelu_f(bf16_round(a[j + q]));
          t0 = __floats2bfloat162_rn(g[0], g[1]); t1 = __floats2bfloat162_rn(g[2], g[3]);
          t2 = __floats2bfloat162_rn(g[4], g[5]); t3 = __floats2bfloat162_rn(g[6], g[7]);
          pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
          pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
          *reinterpret_cast<uint4*>(y_act + o + j) = pk;
        }
      }
    } else {
      for (int j = 0; j < 32 && n0 + j < N; ++j) {
        y[o + j] = __float2bfloat16_rn(a[j]);
        if (act == VITB200_ACT_GELU) y_act[o + j] = __float2bfloat16_rn(gelu_f(bf16_round(a[j])));
      }
    }
  }
  __device__ __forceinline__ void apply1(int, int, float) const {}
};

struct TcEpiDgrad {
  static constexpr bool kSplit = false;
  bf16* dx; const bf16* pre;
  __device__ __forceinline__ void row(int m, int n0, const float (&v)[32], int M, int N) const {
    if (m >= M || n0 >= N) return;
    const size_t o = (size_t)m * N + n0;
    const bool vec = (n0 + 32 <= N) && ((N & 7) == 0);
    if (vec) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        float g[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) g[q] = v[j + q];
        if (pre) {
          uint4 pv = *reinterpret_cast<const uint4*>(pre + o + j);
          const __nv_bfloat162* pp = reinterpret_cast<const __nv_bfloat162*>(&pv);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float2 f = __bfloat1622float2(pp[q]);
            g[2 * q] = bf16_round(g[2 * q]) * gelu_grad_f(f.x);
            g[2 * q + 1] = bf16_round(g[2 * q + 1]) * gelu_grad_f(f.y);
          }
        }
        uint4 pk;
        __nv_bfloat162 t0 = __floats2bfloat162_rn(g[0], g[1]), t1 = __floats2bfloat162_rn(g[2], g[3]);
        __nv_bfloat162 t2 = __floats2bfloat162_rn(g[4], g[5]), t3 = __floats2bfloat162_rn(g[6], g[7]);
        pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
        pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
        *reinterpret_cast<uint4*>(dx + o + j) = pk;
      }
    } else {
      for (int j = 0; j < 32 && n0 + j < N; ++j) {
        float g = v[j];
        if (pre) g = bf16_round(g) * gelu_grad_f(__bfloat162float(pre[o + j]));
        dx[o + j] = __float2bfloat16_rn(g);
      }
    }
  }
  __device__ __forceinline__ void apply1(int, int, float) const {}
};

struct TcEpiWgrad {  // fp32 output [M_out = N_w, N_out = K_w]; split over the contraction (rows of dy / x)
  static constexpr bool kSplit = true;
  float* dw; int ldw; int accumulate;
  __device__ __forceinline__ void apply1(int m, int n, float v) const {
    float* o = dw + (size_t)m * ldw + n;
    *o = accumulate ? *o + v : v;
  }
  __device__ __forceinline__ void row(int m, int n0, const float (&v)[32], int M, int N) const {
    if (m >= M) return;
    for (int j = 0; j < 32 && n0 + j < N; ++j) apply1(m, n0 + j, v[j]);
  }
};

// fp32 output y[m, n] = bf16_round(acc + bias[n]): the input-preprocessor GEMM (few rows, a large weight matrix).  It is
// weight-streaming bound, so the contraction is split across CTAs until the grid covers the SMs (kSplit); the fp32 result
// carries bf16-rounded values, exactly what autocast's bf16 Linear output widened to fp32 holds.
struct TcEpiBiasF32 {
  static constexpr bool kSplit = true;
  float* y; const float* bias; int ldy;
  __device__ __forceinline__ void apply1(int m, int n, float v) const {
    y[(size_t)m * ldy + n] = bf16_round(v + (bias ? bias[n] : 0.f));
  }
  __device__ __forceinline__ void row(int m, int n0, const float (&v)[32], int M, int N) const {
    if (m >= M) return;
    for (int j = 0; j < 32 && n0 + j < N; ++j) apply1(m, n0 + j, v[j]);
  }
};

template <int BN, bool A_MN, bool B_MN, class Epi>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Epi epi, int M, int N,
               int K, int kb_per_split, float* __restrict__ partial, unsigned int* counters) {
  using Cfg = TcCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
  uint8_t* tiles = smem_raw + pad;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* tmem_full = bars + 2 * Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * BN;
  const int total_kb = (K + TC_BK - 1) / TC_BK;
  const int kb0 = blockIdx.z * kb_per_split;
  const int nkb = min(kb_per_split, total_kb - kb0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % Cfg::STAGES, ph = (i / Cfg::STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
        uint8_t* sA = tiles + s * Cfg::STAGE_BYTES;
        uint8_t* sB = sA + TC_A_BYTES;
        const int k0 = (kb0 + i) * TC_BK;
        if (!A_MN) {
          tma_load_2d(sA, &tmA, &full[s], k0, m0);
        } else {
#pragma unroll
          for (int j = 0; j < TC_BM / 64; ++j) tma_load_2d(sA + j * 8192, &tmA, &full[s], m0 + 64 * j, k0);
        }
        if (!B_MN) {
          tma_load_2d(sB, &tmB, &full[s], k0, n0);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(sB + j * 8192, &tmB, &full[s], n0 + 64 * j, k0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(TC_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % Cfg::STAGES, ph = (i / Cfg::STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t sA = smem_u32(tiles + s * Cfg::STAGE_BYTES);
        const uint32_t sB = sA + TC_A_BYTES;
        const int k0 = (kb0 + i) * TC_BK;
        const int ksteps = min(TC_BK / 16, (K - k0 + 15) / 16);
        for (int k = 0; k < ksteps; ++k) {
          // K-major: 16 k-elements = 32 bytes inside the 128B swizzle row.  MN-major: 16 k-rows of 128 B.
          const uint64_t da = A_MN ? make_sdesc_sw128(sA + k * 2048, 8192, 1024) : make_sdesc_sw128(sA + k * 32, 16, 1024);
          const uint64_t db = B_MN ? make_sdesc_sw128(sB + k * 2048, 8192, 1024) : make_sdesc_sw128(sB + k * 32, 16, 1024);
          umma_bf16(tmem_base, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty[s]);  // frees the smem stage once these MMAs have read it
      }
      umma_commit(tmem_full);    // accumulator complete
    }
  } else {
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    const bool split = Epi::kSplit && gridDim.z > 1;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      if (!split) {
        epi.row(m, n0 + c0, v, M, N);
      } else if (m < M) {
        float* dst = partial + (size_t)blockIdx.z * M * N + (size_t)m * N + n0 + c0;
        if (n0 + c0 + 32 <= N && (N & 3) == 0) {   // 16-byte stores: a quarter of the store instructions
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          for (int j = 0; j < 32 && n0 + c0 + j < N; ++j) dst[j] = v[j];
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);

  if constexpr (Epi::kSplit) {
    if (gridDim.z > 1) {
      // deterministic second stage: the last CTA of this output tile sums the splits in split order
      if (!last_block_ticket(&counters[blockIdx.y * gridDim.x + blockIdx.x], gridDim.z)) return;
      const size_t MN = (size_t)M * N;
      const int tm = min(TC_BM, M - m0), tn = min(BN, N - n0);
      const unsigned int Z = gridDim.z;
      if ((tn & 3) == 0 && (N & 3) == 0) {
        // four columns per thread, the splits' loads batched (the scalar loop below is a chain of dependent L2 latencies
        // when Z is small); same z-ascending summation order, so the result is bit-identical to it
        const int tn4 = tn >> 2;
        for (int e = threadIdx.x; e < tm * tn4; e += TC_THREADS) {
          const int mm = m0 + e / tn4, nn = n0 + (e % tn4) * 4;
          const float* src = partial + (size_t)mm * N + nn;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          unsigned int z = 0;
          for (; z + 4 <= Z; z += 4) {
            float4 t[4];
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) t[qq] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(z + qq) * MN));
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) { acc.x += t[qq].x; acc.y += t[qq].y; acc.z += t[qq].z; acc.w += t[qq].w; }
          }
          for (; z < Z; ++z) {
            const float4 t = __ldcg(reinterpret_cast<const float4*>(src + (size_t)z * MN));
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
          }
          epi.apply1(mm, nn, acc.x); epi.apply1(mm, nn + 1, acc.y);
          epi.apply1(mm, nn + 2, acc.z); epi.apply1(mm, nn + 3, acc.w);
        }
        return;
      }
      for (int e = threadIdx.x; e < tm * tn; e += TC_THREADS) {
        const int mm = m0 + e / tn, nn = n0 + e % tn;
        const float* src = partial + (size_t)mm * N + nn;
        float sum = 0.f;
        unsigned int z = 0;
        for (; z + 8 <= Z; z += 8) {
          float t[8];
#pragma unroll
          for (int qq = 0; qq < 8; ++qq) t[qq] = __ldcg(src + (size_t)(z + qq) * MN);
#pragma unroll
          for (int qq = 0; qq < 8; ++qq) sum += t[qq];
        }
        for (; z < Z; ++z) sum += __ldcg(src + (size_t)z * MN);
        epi.apply1(mm, nn, sum);
      }
    }
  }
}

// ---- deterministic column sums (bias gradients): db[n] = sum_m dy[m, n] ----
constexpr int CS_THREADS = 256;
__global__ void __launch_bounds__(CS_THREADS)
colsum_kernel(const bf16* __restrict__ dy, float* __restrict__ db, int M, int N, int rows_per_block, int accumulate,
              float* __restrict__ partial, unsigned int* counter) {
  // thread (slice, col): 256 / min(N,256) row slices
  __shared__ float red[CS_THREADS];
  const int NB = min(N, CS_THREADS), NS = CS_THREADS / NB;
  const int col_l = threadIdx.x % NB, sl = threadIdx.x / NB;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  for (int c0 = 0; c0 < N; c0 += NB) {
    const int c = c0 + col_l;
    float acc = 0.f;
    if (sl < NS && c < N)
      for (int r = r0 + sl; r < r1; r += NS) acc += __bfloat162float(dy[(size_t)r * N + c]);
    red[threadIdx.x] = acc;
    __syncthreads();
    if (sl == 0 && c < N) {
      float t = 0.f;
      for (int k = 0; k < NS; ++k) t += red[k * NB + col_l];
      partial[(size_t)blockIdx.x * N + c] = t;
    }
    __syncthreads();
  }
  if (!last_block_ticket(counter, gridDim.x)) return;
  for (int c = threadIdx.x; c < N; c += CS_THREADS) {
    float s = 0.f;
    unsigned int b = 0;
    const unsigned int nb = gridDim.x;
    for (; b + 8 <= nb; b += 8) {
      float t[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) t[q] = __ldcg(&partial[(size_t)(b + q) * N + c]);
#pragma unroll
      for (int q = 0; q < 8; ++q) s += t[q];
    }
    for (; b < nb; ++b) s += __ldcg(&partial[(size_t)b * N + c]);
    db[c] = accumulate ? db[c] + s : s;
  }
}

template <int BN, bool A_MN, bool B_MN, class Epi>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const Epi& epi, int M, int N, int K, int splits,
                     float* partial, unsigned int* counters, cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, Epi>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) return vb_cuda_error(e);
    attr_done = true;
  }
  const int total_kb = (K + TC_BK - 1) / TC_BK;
  if (splits < 1) splits = 1;
  int kbps = (total_kb + splits - 1) / splits;
  splits = (total_kb + kbps - 1) / kbps;
  dim3 grid((M + TC_BM - 1) / TC_BM, (N + BN - 1) / BN, splits);
  kern<<<grid, TC_THREADS, Cfg::SMEM, st>>>(tmA, tmB, epi, M, N, K, kbps, partial, counters);
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

static inline bool tc_ok_ptr(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int tc_wgrad_splits(int Mout, int Nout, int Kc, int BN) {
  long long tiles = (long long)((Mout + TC_BM - 1) / TC_BM) * ((Nout + BN - 1) / BN);
  if (tiles >= 148 || tiles > 1024) return 1;
  int want = (int)((148 + tiles - 1) / tiles);
  int total_kb = (Kc + TC_BK - 1) / TC_BK;
  int max_by_k = (total_kb + 1) / 2;  // at least two k-blocks per split
  int s = want < max_by_k ? want : max_by_k;
  if (s > 64) s = 64;
  return s < 1 ? 1 : s;
}

}  // namespace vb

using namespace vb;

// Shapes the tensor-core path takes: row pitches must be multiples of 16 bytes, pointers 16B aligned.
extern "C" int vitb200_tc_supported(int M, int N, int K) { return (M > 0 && N % 8 == 0 && K % 8 == 0) ? 1 : 0; }

extern "C" int vitb200_tc_linear_fwd(const void* x, const void* w, const float* bias, void* y, void* y_act, int M, int N,
                                     int K, int act, void* stream) {
  if (!x || !w || !y || M <= 0 || N <= 0 || K <= 0) return VITB200_ERR_ARG;
  if (act == VITB200_ACT_GELU && !y_act) return VITB200_ERR_ARG;
  if (N % 8 || K % 8) return VITB200_ERR_SHAPE;
  if (!tc_ok_ptr(x) || !tc_ok_ptr(w) || !tc_ok_ptr(y) || (y_act && !tc_ok_ptr(y_act))) return VITB200_ERR_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap ta, tb;
  int rc = get_tmap(x, K, M, 64, TC_BM, &ta);
  if (rc) return rc;
  TcEpiBiasAct epi{(bf16*)y, (bf16*)y_act, bias, act};
  if (N <= 64) {
    if ((rc = get_tmap(w, K, N, 64, 64, &tb))) return rc;
    return launch_tc<64, false, false>(ta, tb, epi, M, N, K, 1, nullptr, nullptr, st);
  }
  if ((rc = get_tmap(w, K, N, 64, 128, &tb))) return rc;
  return launch_tc<128, false, false>(ta, tb, epi, M, N, K, 1, nullptr, nullptr, st);
}

// Skinny-M Linear with fp32 output (the input preprocessor, src/models/layers.py:62-63): splits the contraction so that
// tiles x splits <= 148 CTAs stream the weight matrix in one wave.
static int tc_prelinear_splits(int M, int N, int K) {
  const long long tiles = (long long)((M + TC_BM - 1) / TC_BM) * ((N + 127) / 128);
  const int total_kb = (K + TC_BK - 1) / TC_BK;
  long long s = tiles >= 148 ? 1 : 148 / tiles;
  const int max_by_k = (total_kb + 3) / 4;   // at least four k-blocks (one pipeline depth) per split
  if (s > max_by_k) s = max_by_k;
  if (s > 32) s = 32;
  if (tiles > 1024) s = 1;                   // one ticket counter per tile lives in the first 4096 bytes of ws
  return s < 1 ? 1 : (int)s;
}

extern "C" size_t vitb200_tc_prelinear_ws_bytes(int M, int N, int K) {
  const int splits = tc_prelinear_splits(M, N, K);
  return 4096 + (splits > 1 ? (size_t)splits * M * N * sizeof(float) : 0);
}

extern "C" int vitb200_tc_prelinear_fwd(const void* x, const void* w, const float* bias, float* y, int M, int N, int K,
                                        void* ws, void* stream) {
  if (!x || !w || !y || !ws || M <= 0 || N <= 0 || K <= 0) return VITB200_ERR_ARG;
  if (N % 8 || K % 8) return VITB200_ERR_SHAPE;
  if (!tc_ok_ptr(x) || !tc_ok_ptr(w) || !tc_ok_ptr(y)) return VITB200_ERR_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap ta, tb;
  int rc = get_tmap(x, K, M, 64, TC_BM, &ta);
  if (rc) return rc;
  if ((rc = get_tmap(w, K, N, 64, 128, &tb))) return rc;
  TcEpiBiasF32 epi{y, bias, N};
  unsigned int* counters = reinterpret_cast<unsigned int*>(ws);
  float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + 4096);
  return launch_tc<128, false, false>(ta, tb, epi, M, N, K, tc_prelinear_splits(M, N, K), partial, counters, st);
}

// dx[M, K] = dy[M, N] . w[N, K]  (optionally * gelu'(pre_act))
extern "C" int vitb200_tc_linear_dgrad(const void* dy, const void* w, const void* pre_act, void* dx, int M, int N, int K,
                                       void* stream) {
  if (!dy || !w || !dx || M <= 0 || N <= 0 || K <= 0) return VITB200_ERR_ARG;
  if (N % 8 || K % 8) return VITB200_ERR_SHAPE;
  if (!tc_ok_ptr(dy) || !tc_ok_ptr(w) || !tc_ok_ptr(dx) || (pre_act && !tc_ok_ptr(pre_act))) return VITB200_ERR_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap ta, tb;
  int rc = get_tmap(dy, N, M, 64, TC_BM, &ta);  // A = dy, contraction over its columns (K-major)
  if (rc) return rc;
  if ((rc = get_tmap(w, K, N, 64, 64, &tb))) return rc;  // B: rows = contraction index n, cols = output index k
  TcEpiDgrad epi{(bf16*)dx, (const bf16*)pre_act};
  if (K <= 64) return launch_tc<64, false, true>(ta, tb, epi, M, K, N, 1, nullptr, nullptr, st);
  return launch_tc<128, false, true>(ta, tb, epi, M, K, N, 1, nullptr, nullptr, st);
}

extern "C" size_t vitb200_tc_linear_wgrad_ws_bytes(int M, int N, int K) {
  int bn = K <= 64 ? 64 : 128;
  int splits = tc_wgrad_splits(N, K, M, bn);
  size_t gemm = splits > 1 ? (size_t)splits * N * K * sizeof(float) : 0;
  size_t cs = (size_t)148 * N * sizeof(float);
  return 4096 + (gemm > cs ? gemm : cs);
}

// dw[N, K] (+)= dy[M, N]^T . x[M, K] ; dbias[N] (+)= column sums of dy
extern "C" int vitb200_tc_linear_wgrad(const void* dy, const void* x, float* dw, float* dbias, int M, int N, int K,
                                       int accumulate, void* ws, void* stream) {
  if (!dy || !x || !dw || !ws || M <= 0 || N <= 0 || K <= 0) return VITB200_ERR_ARG;
  if (N % 8 || K % 8) return VITB200_ERR_SHAPE;
  if (!tc_ok_ptr(dy) || !tc_ok_ptr(x)) return VITB200_ERR_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned int* counters = reinterpret_cast<unsigned int*>(ws);
  float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + 4096);
  if (dbias) {
    int blocks = (M + 63) / 64;
    if (blocks > 148) blocks = 148;
    int rpb = (M + blocks - 1) / blocks;
    blocks = (M + rpb - 1) / rpb;
    colsum_kernel<<<blocks, CS_THREADS, 0, st>>>((const bf16*)dy, dbias, M, N, rpb, accumulate, partial, counters);
    VB_CHECK_LAUNCH();
  }
  CUtensorMap ta, tb;
  int rc = get_tmap(dy, N, M, 64, 64, &ta);  // A(n, m): rows of dy are the contraction index
  if (rc) return rc;
  if ((rc = get_tmap(x, K, M, 64, 64, &tb))) return rc;
  TcEpiWgrad epi{dw, K, accumulate};
  const int bn = K <= 64 ? 64 : 128;
  const int splits = tc_wgrad_splits(N, K, M, bn);
  if (bn == 64) return launch_tc<64, true, true>(ta, tb, epi, N, K, M, splits, partial, counters, st);
  return launch_tc<128, true, true>(ta, tb, epi, N, K, M, splits, partial, counters, st);
}
