// Plugin entry points OpenMM's plugin loader looks for, with the names and behaviour of the reference's
// /root/reference/platforms/cuda/src/CudaDrudeTGNHKernelFactory.cpp:37-66.
#include "B200DrudeTGNHKernelFactory.h"

#include <exception>

#include "B200DrudeTGNHKernels.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/internal/windowsExport.h"

#ifdef TGNH_WITH_OPENMM
// Real OpenMM: device arrays come from the CUDA platform's CudaContext.  OpenMM is not installed in this repo's container; the
// block is compiled and run against the CUDA-platform stand-in of shim/cuda (oracle/Makefile, target _refcuda/libb200cuda.so).
#include "CudaContext.h"
#include "CudaPlatform.h"
namespace OpenMM {
class CudaContextAccess : public TgnhDeviceAccess {
public:
    explicit CudaContextAccess(CudaContext& cu) : cu(cu) {}
    TgnhDeviceView view() {
        cu.setAsCurrent();
        TgnhDeviceView v;
        v.precision = cu.getUseDoublePrecision() ? TGNH_PRECISION_DOUBLE : cu.getUseMixedPrecision() ? TGNH_PRECISION_MIXED : TGNH_PRECISION_SINGLE;
        v.posqCorrection = cu.getUseMixedPrecision() ? (void*)cu.getPosqCorrection().getDevicePointer() : NULL;
        v.velm = (void*)cu.getVelm().getDevicePointer();
        v.posq = (void*)cu.getPosq().getDevicePointer();
        v.force = (const void*)cu.getForce().getDevicePointer();
        v.posDelta = (void*)cu.getIntegrationUtilities().getPosDelta().getDevicePointer();
        v.paddedNumAtoms = cu.getPaddedNumAtoms();
        v.forceFormat = TGNH_FORCE_I64_SOA;
        v.stream = (void*)cu.getCurrentStream();
        v.device = cu.getDeviceIndex();
        return v;
    }
    /** Keeps cu.reorderAtoms() from swapping molecules whose per-particle tables differ (temperature group, role, offsets):
     *  the kernels' tables are indexed by atom slot and are not permuted (nor are the reference's, which registers nothing). */
    void registerForceInfo(const std::vector<unsigned int>& descriptors, const std::vector<int>& residueOf) {
        class Info : public CudaForceInfo {
        public:
            Info(const std::vector<unsigned int>& d, const std::vector<int>& r) : CudaForceInfo(0), desc(d), res(r) {}
            bool areParticlesIdentical(int i, int j) { return desc[i] == desc[j]; }
            int getNumParticleGroups() { return 0; }          // molecules are already held together by bonds / the DrudeForce
            void getParticlesInGroup(int, std::vector<int>&) {}
            bool areGroupsIdentical(int, int) { return true; }
        private:
            std::vector<unsigned int> desc;
            std::vector<int> res;
        };
        cu.addForce(new Info(descriptors, residueOf));
    }
    void initializeContexts(const System& system) { cu.getPlatformData().initializeContexts(system); }
    bool atomsWereReordered() { return cu.getAtomsWereReordered(); }
    void applyConstraints(double tol) { cu.getIntegrationUtilities().applyConstraints(tol); }
    void computeVirtualSites() { cu.getIntegrationUtilities().computeVirtualSites(); }
    void applyVelocityConstraints(double tol) { cu.getIntegrationUtilities().applyVelocityConstraints(tol); }
    void advanceTime(double dt) {
        cu.setTime(cu.getTime() + dt);
        cu.setStepCount(cu.getStepCount() + 1);
        cu.reorderAtoms();
    }
private:
    CudaContext& cu;
};
}  // namespace OpenMM
#endif

using namespace OpenMM;

extern "C" OPENMM_EXPORT void registerPlatforms() {}

extern "C" OPENMM_EXPORT void registerKernelFactories() {
    try {
        Platform& platform = Platform::getPlatformByName("CUDA");
        platform.registerKernelFactory(IntegrateDrudeTGNHStepKernel::Name(), new B200DrudeTGNHKernelFactory());
    } catch (const std::exception&) {
        // no CUDA platform in this process: nothing to register (the reference swallows this too, :44-48)
    }
}

extern "C" OPENMM_EXPORT void registerDrudeTGNHCudaKernelFactories() { registerKernelFactories(); }

KernelImpl* B200DrudeTGNHKernelFactory::createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const {
    if (name != IntegrateDrudeTGNHStepKernel::Name())
        throw OpenMMException((std::string("Tried to create kernel with illegal kernel name '") + name + "'").c_str());
#ifdef TGNH_WITH_OPENMM
    CudaContext& cu = *static_cast<CudaPlatform::PlatformData*>(context.getPlatformData())->contexts[0];
    return new B200IntegrateDrudeTGNHStepKernel(name, platform, *new CudaContextAccess(cu), true);
#else
    // shim build: the platform data IS the device access object (plugin/tests/ShimCudaPlatform.h)
    TgnhDeviceAccess* access = static_cast<TgnhDeviceAccess*>(context.getPlatformData());
    if (access == NULL) throw OpenMMException("DrudeTGNH: the CUDA platform holds no device arrays for this context");
    return new B200IntegrateDrudeTGNHStepKernel(name, platform, *access);
#endif
}
