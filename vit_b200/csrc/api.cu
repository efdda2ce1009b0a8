// Library-level entry points: version, error strings, device check.
#include <string.h>
#include <stdio.h>

#include "common.cuh"

static thread_local char g_cuda_err[256] = "";

int vb_cuda_error(cudaError_t e) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
  return VITB200_ERR_CUDA;
}

extern "C" const char* vitb200_last_cuda_error(void) { return g_cuda_err; }

extern "C" const char* vitb200_strerror(int rc) {
  switch (rc) {
    case VITB200_OK: return "ok";
    case VITB200_ERR_CUDA: return "CUDA error (see vitb200_last_cuda_error)";
    case VITB200_ERR_SHAPE: return "unsupported shape (need H % 4 == 0, H <= 1024 for LayerNorm, head_dim in {8,16,32,64,128}, arena length % 4 == 0)";
    case VITB200_ERR_ARG: return "invalid argument (null pointer or inconsistent sizes)";
    case VITB200_ERR_ALIGN: return "pointer not sufficiently aligned (16 bytes for fp32 buffers)";
    case VITB200_ERR_DEVICE: return "device is not compute capability 10.x (this library is sm_100a only)";
    default: return "unknown error";
  }
}

extern "C" int vitb200_version(void) { return VITB200_VERSION; }

extern "C" int vitb200_init(int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return vb_cuda_error(e);
  if (prop.major != 10) return VITB200_ERR_DEVICE;
  return VITB200_OK;
}
