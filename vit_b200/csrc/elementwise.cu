// Small elementwise kernels used only by the modular (hook-friendly, eval) path of the module tree.
#include "common.cuh"

namespace vb {
constexpr int EW_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(EW_THREADS) gelu_kernel(const T* __restrict__ x, T* __restrict__ y, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < n4; i += (size_t)gridDim.x * EW_THREADS) {
    float4 v = Vec4<T>::ld(x + i * 4);
    v.x = gelu_f(v.x); v.y = gelu_f(v.y); v.z = gelu_f(v.z); v.w = gelu_f(v.w);
    Vec4<T>::st(y + i * 4, v);
  }
}
template <typename T>
__global__ void __launch_bounds__(EW_THREADS)
residual_add_kernel(const float* __restrict__ z, const T* __restrict__ d, float* __restrict__ out, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < n4; i += (size_t)gridDim.x * EW_THREADS) {
    float4 a = *reinterpret_cast<const float4*>(z + i * 4);
    float4 b = Vec4<T>::ld(d + i * 4);
    *reinterpret_cast<float4*>(out + i * 4) = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}
static inline int ew_grid(size_t n4) {
  size_t g = (n4 + EW_THREADS - 1) / EW_THREADS;
  return (int)(g > 1184 ? 1184 : (g < 1 ? 1 : g));
}
}  // namespace vb

using namespace vb;

extern "C" int vitb200_gelu_fwd(const void* x, void* y, size_t n, int dtype, void* stream) {
  if (!x || !y) return VITB200_ERR_ARG;
  if (n % 4 != 0) return VITB200_ERR_SHAPE;
  if (n == 0) return VITB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VITB200_F32) gelu_kernel<float><<<ew_grid(n / 4), EW_THREADS, 0, st>>>((const float*)x, (float*)y, n / 4);
  else if (dtype == VITB200_BF16) gelu_kernel<bf16><<<ew_grid(n / 4), EW_THREADS, 0, st>>>((const bf16*)x, (bf16*)y, n / 4);
  else return VITB200_ERR_ARG;
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}

extern "C" int vitb200_residual_add(const float* z, const void* delta, float* out, size_t n, int dtype, void* stream) {
  if (!z || !delta || !out) return VITB200_ERR_ARG;
  if (n % 4 != 0) return VITB200_ERR_SHAPE;
  if (n == 0) return VITB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VITB200_F32) residual_add_kernel<float><<<ew_grid(n / 4), EW_THREADS, 0, st>>>(z, (const float*)delta, out, n / 4);
  else if (dtype == VITB200_BF16) residual_add_kernel<bf16><<<ew_grid(n / 4), EW_THREADS, 0, st>>>(z, (const bf16*)delta, out, n / 4);
  else return VITB200_ERR_ARG;
  VB_CHECK_LAUNCH();
  return VITB200_OK;
}
