// shim/cuda/CudaBondedUtilities.h — included by the reference's CudaDrudeTGNHKernels.cpp, no member is used.  TEST / BUILD INFRASTRUCTURE.
#ifndef SHIM_CUDA_BONDED_UTILITIES_H_
#define SHIM_CUDA_BONDED_UTILITIES_H_
namespace OpenMM {
class CudaBondedUtilities {};
}
#endif
