// Low-rank ZCA preprocessor as two skinny products instead of a dense [D, D] matrix.
//
// The reference builds  P = Vr diag(1/sqrt(lam_r + eps)) Vr^T + s_perp (I - Vr Vr^T)   (src/models/preprocessor.py:40-72)
// and applies it as a dense Linear  y = x P^T + b  (src/models/layers.py:62-63): a 4096 x 4096 matrix is 32 MB of bf16
// streamed per step for 64 spectra.  While the matrix is frozen it is exactly
//     y = s_perp x + ((x Vr) o g) Vr^T + b ,      g = 1/sqrt(lam_r + eps) - s_perp      (P is symmetric)
// i.e. 2 r D values of Vr (256 KB of bf16 at r = 32) and 4 r D flops per spectrum.
//
// One thread-block cluster per spectrum (cluster size 1 or 2 splits the D pixels): thread = a few pixels d.
//   phase 1: acc[k] += x_d Vr[d, k] over the thread's pixels (each lane reads whole 64 / 128-byte rows of Vr), reduced over
//            the warp by a 31-shuffle reduce-scatter, over the warps through shared memory, over the cluster through
//            distributed shared memory -- every sum in a fixed order (bitwise reproducible);
//   phase 2: y_d = s_perp x_d + sum_k t_k g_k Vr[d, k] + b_d  (Vr rows are re-read from L1 / L2).
// bf16 mode: x and Vr are bf16 operands, fp32 accumulation, the output is rounded to bf16 (stored as fp32 in the engine's
// pixel buffer, like vitb200_tc_prelinear_fwd); fp32 mode: everything fp32.
#include "common.cuh"

namespace vb {

constexpr int ZL_THREADS = 512;

__device__ __forceinline__ uint32_t zl_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void zl_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void zl_st_peer(float* p, uint32_t peer, float v) {
  uint32_t a;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(peer));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}

// R values of row d of Vr as floats
template <int R>
__device__ __forceinline__ void zl_load_row(const bf16* __restrict__ vr, size_t d, float (&v)[R]) {
  const uint4* src = reinterpret_cast<const uint4*>(vr + d * R);
#pragma unroll
  for (int c = 0; c < R / 8; ++c) {
    const uint4 pk = __ldg(src + c);
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
    for (int q = 0; q < 4; ++q) { const float2 f = __bfloat1622float2(p[q]); v[c * 8 + 2 * q] = f.x; v[c * 8 + 2 * q + 1] = f.y; }
  }
}
template <int R>
__device__ __forceinline__ void zl_load_row(const float* __restrict__ vr, size_t d, float (&v)[R]) {
  const float4* src = reinterpret_cast<const float4*>(vr + d * R);
#pragma unroll
  for (int c = 0; c < R / 4; ++c) {
    const float4 a = __ldg(src + c);
    v[4 * c] = a.x; v[4 * c + 1] = a.y; v[4 * c + 2] = a.z; v[4 * c + 3] = a.w;
  }
}

// warp reduce-scatter of 32 values per lane: afterwards lane l holds the warp total of v[l] (fixed order, 31 shuffles)
__device__ __forceinline__ float zl_reduce_scatter32(float (&v)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = up ? v[i] : v[i + o];
      const float keep = up ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

template <int R, typename TV, bool BF>
__global__ void __launch_bounds__(ZL_THREADS, 1)
zca_lowrank_kernel(const float* __restrict__ x, const TV* __restrict__ vr, const float* __restrict__ g, float s_perp,
                   const float* __restrict__ bias, float* __restrict__ y, int B, int D, int csz) {
  __shared__ float s_part[ZL_THREADS / 32][R];   // warp partials of t
  __shared__ float s_t[2][R];                     // cluster exchange: t of rank 0 / rank 1
  __shared__ __align__(16) float s_tg[R];         // t o g
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t crank = csz == 2 ? zl_cluster_rank() : 0u;
  const int b = (int)blockIdx.x / csz;
  const int Dh = D / csz, d0 = (int)crank * Dh;     // this CTA's pixels: [d0, d0 + Dh)
  // bf16, r <= 32, <= 4 pixels per thread (D <= 4096 with a CTA pair): the thread's Vr rows (4 x 64 B) are fetched BEFORE the
  // dependency wait -- the matrix is frozen, only x comes from the previous kernel -- and stay in registers for phase 2
  constexpr bool CAN_KEEP = BF && R == 32;
  const bool keep = CAN_KEEP && Dh <= 4 * ZL_THREADS;
  uint4 rows[CAN_KEEP ? 4 : 1][CAN_KEEP ? 4 : 1];
  if (CAN_KEEP && keep) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = d0 + j * ZL_THREADS + tid;
      if (d < d0 + Dh) {
        const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(vr) + (size_t)d * R);
#pragma unroll
        for (int c = 0; c < 4; ++c) rows[j][c] = __ldg(src + c);
      }
    }
  }
  pdl_wait();
  pdl_trigger();
  const float* xr = x + (size_t)b * D;
  // ---- phase 1: t = x Vr over this CTA's pixels ----
  float acc[R];
#pragma unroll
  for (int k = 0; k < R; ++k) acc[k] = 0.f;
  float xk[4] = {0.f, 0.f, 0.f, 0.f};
  if (CAN_KEEP && keep) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = d0 + j * ZL_THREADS + tid;
      if (d < d0 + Dh) {
        xk[j] = bf16_round(xr[d]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&rows[j][c]);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = __bfloat1622float2(p2[q]);
            acc[c * 8 + 2 * q] = fmaf(xk[j], f.x, acc[c * 8 + 2 * q]);
            acc[c * 8 + 2 * q + 1] = fmaf(xk[j], f.y, acc[c * 8 + 2 * q + 1]);
          }
        }
      }
    }
  } else {
    for (int d = d0 + tid; d < d0 + Dh; d += ZL_THREADS) {
      float xd = xr[d];
      if (BF) xd = bf16_round(xd);
      float v[R];
      zl_load_row<R>(vr, (size_t)d, v);
#pragma unroll
      for (int k = 0; k < R; ++k) acc[k] = fmaf(xd, v[k], acc[k]);
    }
  }
#pragma unroll
  for (int h = 0; h < R / 32; ++h) {
    float part[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) part[k] = acc[h * 32 + k];
    s_part[warp][h * 32 + lane] = zl_reduce_scatter32(part, lane);
  }
  __syncthreads();
  if (tid < R) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < ZL_THREADS / 32; ++w) t += s_part[w][tid];
    if (csz == 2) {
      s_t[crank][tid] = t;
      zl_st_peer(&s_t[crank][tid], crank ^ 1u, t);
    } else {
      s_tg[tid] = t * g[tid];
    }
  }
  if (csz == 2) {
    zl_cluster_sync();
    if (tid < R) s_tg[tid] = (s_t[0][tid] + s_t[1][tid]) * g[tid];
  }
  __syncthreads();
  // ---- phase 2: y = s_perp x + (t o g) Vr^T + b ----
  if (CAN_KEEP && keep) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = d0 + j * ZL_THREADS + tid;
      if (d < d0 + Dh) {
        float o = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&rows[j][c]);
          const float4 ta = *reinterpret_cast<const float4*>(&s_tg[c * 8]), tb = *reinterpret_cast<const float4*>(&s_tg[c * 8 + 4]);
          const float2 f0 = __bfloat1622float2(p2[0]), f1 = __bfloat1622float2(p2[1]), f2 = __bfloat1622float2(p2[2]),
                       f3 = __bfloat1622float2(p2[3]);
          o = fmaf(ta.x, f0.x, o); o = fmaf(ta.y, f0.y, o); o = fmaf(ta.z, f1.x, o); o = fmaf(ta.w, f1.y, o);
          o = fmaf(tb.x, f2.x, o); o = fmaf(tb.y, f2.y, o); o = fmaf(tb.z, f3.x, o); o = fmaf(tb.w, f3.y, o);
        }
        o = fmaf(s_perp, xk[j], o) + (bias ? bias[d] : 0.f);
        y[(size_t)b * D + d] = bf16_round(o);
      }
    }
  } else {
    for (int d = d0 + tid; d < d0 + Dh; d += ZL_THREADS) {
      float xd = xr[d];
      if (BF) xd = bf16_round(xd);
      float v[R];
      zl_load_row<R>(vr, (size_t)d, v);
      float o = 0.f;
#pragma unroll
      for (int k = 0; k < R; k += 4) {
        const float4 tg = *reinterpret_cast<const float4*>(&s_tg[k]);
        o = fmaf(tg.x, v[k], o); o = fmaf(tg.y, v[k + 1], o); o = fmaf(tg.z, v[k + 2], o); o = fmaf(tg.w, v[k + 3], o);
      }
      o = fmaf(s_perp, xd, o) + (bias ? bias[d] : 0.f);
      y[(size_t)b * D + d] = BF ? bf16_round(o) : o;
    }
  }
  if (csz == 2) zl_cluster_sync();   // no CTA exits while its peer may still write its shared memory
}

}  // namespace vb

using namespace vb;

extern "C" int vitb200_zca_lowrank_supported(int D, int R) { return (R == 32 || R == 64) && D > 0 && D % 2 == 0; }

// x [B, D] f32, vr [D, R] (bf16 if dtype == BF16 else f32; columns beyond the true rank are zero), g [R] f32,
// bias [D] f32 or NULL, y [B, D] f32
extern "C" int vitb200_zca_lowrank_fwd(const float* x, const void* vr, const float* g, float s_perp, const float* bias,
                                       float* y, int B, int D, int R, int dtype, void* stream) {
  if (!x || !vr || !g || !y || B <= 0) return VITB200_ERR_ARG;
  if (!vitb200_zca_lowrank_supported(D, R)) return VITB200_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(vr) & 15) != 0) return VITB200_ERR_ALIGN;
  if (dtype != VITB200_BF16 && dtype != VITB200_F32) return VITB200_ERR_ARG;
  const int csz = B <= 74 ? 2 : 1;   // few spectra: two CTAs share one, so that more than B SMs pull Vr
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)(B * csz)), block(ZL_THREADS);
#define LAUNCH_ZL(RR)                                                                                                     \
  if (dtype == VITB200_BF16)                                                                                              \
    vb_launch_pdl_cluster(zca_lowrank_kernel<RR, bf16, true>, grid, block, 0, st, csz, x, (const bf16*)vr, g, s_perp, bias, y, B, D, csz); \
  else                                                                                                                    \
    vb_launch_pdl_cluster(zca_lowrank_kernel<RR, float, false>, grid, block, 0, st, csz, x, (const float*)vr, g, s_perp, bias, y, B, D, csz);
  if (R == 32) { LAUNCH_ZL(32) } else { LAUNCH_ZL(64) }
#undef LAUNCH_ZL
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}
