// shim/cuda/CudaForceInfo.h — stand-in for OpenMM 7.x's CudaForceInfo (the interface cu.addForce takes; it tells the atom
// reordering which particles / groups are interchangeable).  TEST / BUILD INFRASTRUCTURE.
#ifndef SHIM_CUDA_FORCE_INFO_H_
#define SHIM_CUDA_FORCE_INFO_H_
#include <vector>
namespace OpenMM {
class CudaForceInfo {
public:
    explicit CudaForceInfo(int requiredForceBuffers) : requiredForceBuffers(requiredForceBuffers) {}
    virtual ~CudaForceInfo() {}
    int getRequiredForceBuffers() { return requiredForceBuffers; }
    virtual bool areParticlesIdentical(int particle1, int particle2) { return true; }
    virtual int getNumParticleGroups() { return 0; }
    virtual void getParticlesInGroup(int index, std::vector<int>& particles) {}
    virtual bool areGroupsIdentical(int group1, int group2) { return true; }
private:
    int requiredForceBuffers;
};
}  // namespace OpenMM
#endif
