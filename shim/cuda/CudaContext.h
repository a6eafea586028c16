// shim/cuda/CudaContext.h — stand-in for OpenMM 7.x's CudaContext (platforms/cuda/include/CudaContext.h), written from the
// members that the reference's platforms/cuda sources (CudaDrudeTGNHKernels.cpp, CudaDrudeTGNHKernelFactory.cpp) and this
// repo's OpenMM-facing glue (plugin/src/B200DrudeTGNHKernelFactory.cpp, -DTGNH_WITH_OPENMM) call.
// TEST / BUILD INFRASTRUCTURE (shim/README.md).  What is real: device arrays in OpenMM's layouts (velm mixed4, posq real4,
// posqCorrection real4 in mixed mode, force long long[3 * paddedNumAtoms] fixed point 2^32), run-time compilation of
// kernel source strings with NVRTC behind OpenMM's prelude (createModule), launches through the driver API with OpenMM's
// grid rule (executeKernel), the atom-reordering bookkeeping.  What is not: neighbour lists, force kernels (a pluggable
// device force model instead), constraints.
#ifndef SHIM_CUDA_CONTEXT_H_
#define SHIM_CUDA_CONTEXT_H_
#include <cuda.h>
#include <vector_functions.h>
#include <vector_types.h>

#include <map>
#include <string>
#include <vector>

#include "CudaArray.h"
#include "CudaForceInfo.h"
#include "CudaIntegrationUtilities.h"
#include "CudaPlatform.h"

namespace OpenMM {

class CudaContext {
public:
    static const int ThreadBlockSize = 64;      // OpenMM 7.x
    static const int TileSize = 32;
    CudaContext(const System& system, const std::string& precision, CudaPlatform::PlatformData& platformData);
    ~CudaContext();
    CudaPlatform::PlatformData& getPlatformData() { return platformData; }
    void setAsCurrent();
    CUcontext getContext() { return cuContext; }
    CUstream getCurrentStream() { return 0; }
    int getDeviceIndex() const { return deviceIndex; }
    int getNumAtoms() const { return numAtoms; }
    int getPaddedNumAtoms() const { return paddedNumAtoms; }
    int getNumThreadBlocks() const { return numThreadBlocks; }
    bool getUseDoublePrecision() const { return useDoublePrecision; }
    bool getUseMixedPrecision() const { return useMixedPrecision; }
    CudaArray& getPosq() { return *posq; }
    CudaArray& getPosqCorrection() { return *posqCorrection; }
    CudaArray& getVelm() { return *velm; }
    CudaArray& getForce() { return *force; }
    CudaArray& getAtomIndexArray() { return *atomIndexDevice; }
    /** slot -> original particle index (OpenMM: cu.getAtomIndex()) */
    const std::vector<int>& getAtomIndex() const { return atomIndex; }
    CudaIntegrationUtilities& getIntegrationUtilities() { return *integration; }
    /** OpenMM: prepends the precision typedefs / macros and `defines`, compiles at run time (here: NVRTC, sm_100a cubin) */
    CUmodule createModule(const std::string source, const char* optimizationFlags = NULL);
    CUmodule createModule(const std::string source, const std::map<std::string, std::string>& defines, const char* optimizationFlags = NULL);
    CUfunction getKernel(CUmodule& module, const std::string& name);
    /** OpenMM: grid = min(ceil(threads / blockSize), numThreadBlocks), blockSize defaults to ThreadBlockSize */
    void executeKernel(CUfunction kernel, void** arguments, int threads, int blockSize = -1, unsigned int sharedSize = 0);
    void clearBuffer(CudaArray& array);
    std::string intToString(int value) const;
    std::string doubleToString(double value) const;
    double getTime() { return time; }
    void setTime(double t) { time = t; }
    int getStepCount() { return stepCount; }
    void setStepCount(int steps) { stepCount = steps; }
    void addForce(CudaForceInfo* force) { forces.push_back(force); }
    std::vector<CudaForceInfo*>& getForceInfos() { return forces; }
    /** OpenMM sorts atoms along a space-filling curve every few hundred steps, molecule-wise, only swapping molecules all
     *  registered CudaForceInfos call identical.  The shim swaps pairs of neighbouring interchangeable molecules every
     *  `shimReorderInterval` steps (0 = never, the default): enough to exercise everything that depends on the order. */
    void reorderAtoms();
    bool getAtomsWereReordered() const { return atomsWereReordered; }
    void setAtomsWereReordered(bool wereReordered) { atomsWereReordered = wereReordered; }
    int shimReorderInterval;
    int shimReorderCount;      // reorders that moved atoms
    int shimReorderAttempts;
    /** the source handed to NVRTC by the most recent createModule (tests look at the prelude) */
    std::string shimLastSource;
    long long shimKernelLaunches;
    const System& getSystem() const { return system; }
private:
    const System& system;
    CudaPlatform::PlatformData& platformData;
    CUcontext cuContext;
    CUdevice device;
    int deviceIndex, numAtoms, paddedNumAtoms, numThreadBlocks, stepCount;
    bool useDoublePrecision, useMixedPrecision, atomsWereReordered;
    double time;
    std::map<std::string, std::string> compilationDefines;
    CudaArray *posq, *posqCorrection, *velm, *force, *atomIndexDevice;
    std::vector<int> atomIndex;
    std::vector<CudaForceInfo*> forces;
    std::vector<CUmodule> modules;
    CudaIntegrationUtilities* integration;
};

}  // namespace OpenMM
#endif
