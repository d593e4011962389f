// Small host helpers shared by the translation units of libhdsm.
#pragma once
#include <cuda_runtime.h>

namespace hdsm {

// cudaFuncAttributeMaxDynamicSharedMemorySize is an attribute of the KERNEL on the current device, shared by every
// handle of the process.  Each handle therefore raises it to the device's opt-in maximum - the same value for
// all of them - instead of its own need: a second, smaller handle can then never lower the limit under a live
// larger one (whose launches would start failing with cudaErrorInvalidValue).
template <class Kernel>
inline cudaError_t raise_smem_limit(Kernel kernel, int device) {
  int optin = 0;
  cudaError_t e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  if (e != cudaSuccess) return e;
  cudaFuncAttributes fa{};
  if ((e = cudaFuncGetAttributes(&fa, kernel)) != cudaSuccess) return e;
  // the opt-in maximum covers static + dynamic shared memory of a block
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
}

}  // namespace hdsm
