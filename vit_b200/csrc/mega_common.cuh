// Shared pieces of the whole-network kernels (mega_fwd.cu, mega_bwd.cu): N-dimensional TMA, swizzled-tile accessors,
// the single-row ("side row") SIMT helpers and the shared-memory plan.
//
// One CTA owns whole SAMPLES: 128 token rows run on tcgen05 (TMEM lane = row), the 129th token of the configured
// sequence (T = 128 patches + CLS = 129) runs beside them on a 17th warp with plain FMAs -- the same split the
// stand-alone attention kernel uses (attention_tc.cu).  Everything a sample needs between the input spectrum and the
// logits stays in shared memory / TMEM / registers; HBM only sees the tensors backward needs.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_host.cuh"

namespace vb {
using namespace vb::tc;

constexpr int MG_H = 32, MG_I = 128, MG_D = 16, MG_NH = 2, MG_Q = 96;
constexpr int MG_CG = 4;                       // column groups = threads per row
constexpr int MG_HC = MG_H / MG_CG;            // 8 residual-stream columns per thread
constexpr int MG_MAIN = 128 * MG_CG;           // 512 MMA-path threads
constexpr int MG_THREADS = MG_MAIN + 32;       // + the side-row warp
constexpr int MG_SIDE_KEYS = 5;                // keys per lane of the side-row warp (covers 160 >= 144)
constexpr int MG_PRM_LAYER = 416;              // staged fp32 vectors per layer (see MP_* below)
constexpr int MP_LN1G = 0, MP_LN1B = 32, MP_BQ = 64, MP_BO = 160, MP_G2 = 192, MP_B2LN = 224, MP_B1 = 256, MP_B2 = 384;
constexpr float MG_LOG2E = 1.4426950408889634f;

// ---- TMA, 3-D / 4-D (tensor maps over [cols, rows-of-a-sample, samples, layers]: a 128-row box never crosses into the
// next sample: rows past T are clipped on store and zero-filled on load) --------------------------------------------
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// bf16 tensor map of rank `rank` (<= 4): dims / strides innermost first (strides in BYTES for dims 1..rank-1), 128B swizzle
static int get_tmap_nd(const void* p, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                       CUtensorMap* out) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return VITB200_ERR_DEVICE;
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t bx[4], es[4];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(p), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? VITB200_OK : VITB200_ERR_ARG;
}
// activation [layers, B, T, cols] (bf16 elements; fp32 rows are passed as 2 x cols bf16), box = 64 cols x 128 rows
static int get_tmap_act(const void* p, int cols, int T, int B, int layers, CUtensorMap* out) {
  const uint64_t dims[4] = {(uint64_t)cols, (uint64_t)T, (uint64_t)B, (uint64_t)layers};
  const uint64_t str[3] = {(uint64_t)cols * 2, (uint64_t)cols * 2 * T, (uint64_t)cols * 2 * T * B};
  const uint32_t box[4] = {64, 128, 1, 1};
  return get_tmap_nd(p, 4, dims, str, box, out);
}
// weight [layers][rows, cols] inside the bf16 shadow arena (layer stride in elements), box = 64 cols x box_rows
static int get_tmap_weight(const void* p, int cols, int rows, int layers, size_t layer_stride, int box_rows, CUtensorMap* out) {
  const uint64_t dims[3] = {(uint64_t)cols, (uint64_t)rows, (uint64_t)(layers > 0 ? layers : 1)};
  const uint64_t str[2] = {(uint64_t)cols * 2, (uint64_t)(layer_stride > 0 ? layer_stride : (size_t)rows * cols) * 2};
  const uint32_t box[3] = {64, (uint32_t)box_rows, 1};
  return get_tmap_nd(p, 3, dims, str, box, out);
}

// ---- swizzled tiles (128-byte rows, 8-row groups of 1 KB: 16-byte chunk c of row r sits at chunk c ^ (r & 7)) -------
__device__ __forceinline__ uint4 mg_pack8(const float* v) {
  __nv_bfloat162 t0 = __floats2bfloat162_rn(v[0], v[1]), t1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 t2 = __floats2bfloat162_rn(v[4], v[5]), t3 = __floats2bfloat162_rn(v[6], v[7]);
  uint4 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
  pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
  return pk;
}
__device__ __forceinline__ void mg_unpack8(uint4 pk, float* v) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
  for (int q = 0; q < 4; ++q) { float2 f = __bfloat1622float2(p[q]); v[2 * q] = f.x; v[2 * q + 1] = f.y; }
}
// chunk pointer inside one 64-column block (rows of 128 B)
__device__ __forceinline__ uint8_t* mg_chunk(uint8_t* blk, int r, int chunk) {
  return blk + r * 128 + ((chunk ^ (r & 7)) << 4);
}
__device__ __forceinline__ const uint8_t* mg_chunk(const uint8_t* blk, int r, int chunk) {
  return blk + r * 128 + ((chunk ^ (r & 7)) << 4);
}
// multi-block tile of 128-row blocks (16 KB each): chunk index runs over the whole row
__device__ __forceinline__ uint8_t* mg_swz(uint8_t* tile, int r, int chunk) {
  return tile + (chunk >> 3) * 16384 + r * 128 + (((chunk & 7) ^ (r & 7)) << 4);
}
__device__ __forceinline__ const uint8_t* mg_swz(const uint8_t* tile, int r, int chunk) {
  return tile + (chunk >> 3) * 16384 + r * 128 + (((chunk & 7) ^ (r & 7)) << 4);
}
// one bf16 element (row r, column c < 64) of a 64-column block
__device__ __forceinline__ bf16* mg_elem(uint8_t* blk, int r, int c) {
  return reinterpret_cast<bf16*>(blk + r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2);
}
__device__ __forceinline__ const bf16* mg_elem(const uint8_t* blk, int r, int c) {
  return reinterpret_cast<const bf16*>(blk + r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2);
}
// 16-byte piece (floats [c, c+4)) of an fp32 [128, 32] tile staged as 128-byte rows
__device__ __forceinline__ float4* mg_f32(uint8_t* tile, int r, int c) {
  return reinterpret_cast<float4*>(tile + r * 128 + (((c >> 2) ^ (r & 7)) << 4));
}

// ---- UMMA operand views ---------------------------------------------------------------------------------------------
struct MgOp { uint32_t addr, lbo, kblk; int mn; };
__device__ __forceinline__ uint64_t mg_desc(const MgOp& o, int k) {
  if (o.mn) return make_sdesc_sw128(o.addr + k * 2048, o.lbo, 1024);
  return make_sdesc_sw128(o.addr + (k >> 2) * o.kblk + (k & 3) * 32, 16, 1024);
}
// D[128, N] (+)= A * B over `ksteps` k-steps of 16 (issued by one thread).  The two descriptors are built once and
// advanced by adding to their start-address field (16-byte units): a k-step is +32 B inside a K-major 64-column block
// (the next block after 4 steps) or +2048 B (16 rows) of an MN-major tile.  The issue loop is on the critical path of
// every phase: rebuilding both descriptors from scratch cost ~12 uniform-datapath instructions per MMA.
__device__ __forceinline__ void mg_issue(uint32_t tmem_d, const MgOp& A, const MgOp& B, int N, int ksteps, bool acc) {
  const uint32_t idesc = make_idesc_bf16(128, N, A.mn, B.mn);
  uint64_t da = mg_desc(A, 0), db = mg_desc(B, 0);
  const uint32_t a_in = A.mn ? 128u : 2u, b_in = B.mn ? 128u : 2u;
  const uint32_t a_blk = A.mn ? 128u : ((A.kblk - 96u) >> 4), b_blk = B.mn ? 128u : ((B.kblk - 96u) >> 4);
  for (int k = 0; k < ksteps; ++k) {
    umma_bf16(tmem_d, da, db, idesc, (acc || k > 0) ? 1u : 0u);
    const bool wrap = (k & 3) == 3;
    da += wrap ? a_blk : a_in;
    db += wrap ? b_blk : b_in;
  }
}

__device__ __forceinline__ void mg_bar_main() { asm volatile("bar.sync 1, 512;" ::: "memory"); }   // the 16 MMA-path warps

// ---- side row (one warp, lane = column) -----------------------------------------------------------------------------
// fixed-order sum of the per-CTA loss partials by one warp: lane l adds partials l, l + 32, ... (<= 160, all loads in
// flight at once), then a shuffle tree; every lane returns the total
__device__ __forceinline__ float mg_loss_sum(const float* __restrict__ part, unsigned int n, int lane) {
  float v[5];
#pragma unroll
  for (int q = 0; q < 5; ++q) v[q] = (unsigned)(lane + 32 * q) < n ? __ldcg(&part[lane + 32 * q]) : 0.f;
  float t = ((v[0] + v[1]) + (v[2] + v[3])) + v[4];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}
__device__ __forceinline__ float mg_wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float mg_wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// y[i] = sum_{k<32} a[k] * W[row0 + 32 i][k]   (i < NR): W = K-major swizzled tile (rows of 128 B, 32 valid columns),
// a is spread over the warp (lane k holds a[k]); every lane computes its own NR rows.
template <int NR>
__device__ __forceinline__ void mg_side_gemv32(const uint8_t* wt, int row0, float a_lane, float (&y)[NR]) {
#pragma unroll
  for (int i = 0; i < NR; ++i) y[i] = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {   // 8 contraction columns at a time: NR x 8 weights live in registers
    float w[NR][8];
#pragma unroll
    for (int i = 0; i < NR; ++i) mg_unpack8(*reinterpret_cast<const uint4*>(mg_chunk(wt, row0 + 32 * i, c)), w[i]);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float ak = __shfl_sync(0xffffffffu, a_lane, 8 * c + q);
#pragma unroll
      for (int i = 0; i < NR; ++i) y[i] = fmaf(ak, w[i][q], y[i]);
    }
  }
}
// LayerNorm of the side row (lane = column, H = 32): two-pass mean / variance
__device__ __forceinline__ void mg_side_ln(float x, float eps, float& mu, float& rs) {
  mu = mg_wsum(x) * (1.f / 32.f);
  const float d = x - mu;
  rs = rsqrtf(mg_wsum(d * d) * (1.f / 32.f) + eps);
}

}  // namespace vb
