// shim/cuda/CudaKernelPrelude.h — what OpenMM's CudaContext puts in front of every kernel source before compiling it at run time,
// and the run-time compilation itself (NVRTC, sm_100a).  No CUDA driver dependency: usable on a machine without a GPU.
// TEST / BUILD INFRASTRUCTURE (shim/README.md).
#ifndef SHIM_CUDA_KERNEL_PRELUDE_H_
#define SHIM_CUDA_KERNEL_PRELUDE_H_
#include <map>
#include <string>
#include <vector>
namespace OpenMM {
/** the `compilationDefines` OpenMM 7.3/7.4's CudaContext constructor sets for a precision mode */
std::map<std::string, std::string> shimCompilationDefines(bool useDoublePrecision, bool useMixedPrecision);
/** CudaContext::createModule's source assembly: options comment, compilationDefines not overridden by `defines`, the real/mixed
 *  typedefs, tileflags, `defines`, then the source */
std::string shimBuildKernelSource(bool useDoublePrecision, bool useMixedPrecision, const std::map<std::string, std::string>& compilationDefines,
                                  const std::string& source, const std::map<std::string, std::string>& defines, const std::string& options);
/** NVRTC, --gpu-architecture=sm_100a (+ --use_fast_math when `options` asks for it); false + log on failure */
bool shimNvrtcCompile(const std::string& source, const std::string& options, std::vector<char>& cubin, std::string& log);
}  // namespace OpenMM
#endif
