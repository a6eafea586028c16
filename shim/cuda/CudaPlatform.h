// shim/cuda/CudaPlatform.h — stand-in for OpenMM 7.x's CudaPlatform / CudaPlatform::PlatformData.  TEST / BUILD INFRASTRUCTURE.
// Context property "Precision" (or OpenMM 7.x's "CudaPrecision") = single | mixed | double, default single as in OpenMM.
#ifndef SHIM_CUDA_PLATFORM_H_
#define SHIM_CUDA_PLATFORM_H_
#include <map>
#include <string>
#include <vector>

#include "openmm/Platform.h"
#include "openmm/internal/ContextImpl.h"

namespace OpenMM {
class CudaContext;
class CudaPlatform : public Platform {
public:
    class PlatformData {
    public:
        PlatformData(ContextImpl* context, const System& system, const std::string& precision);
        ~PlatformData();
        /** OpenMM: finishes setting up the contexts once all kernels are created; here: first upload of the host state */
        void initializeContexts(const System& system);
        ContextImpl* context;
        std::vector<CudaContext*> contexts;
        bool contextsInitialized;
        int initializeCalls;
    };
    CudaPlatform() {}
    const std::string& getName() const { static const std::string n = "CUDA"; return n; }
    double getSpeed() const { return 100; }
    void contextCreated(ContextImpl& context, const std::map<std::string, std::string>& properties) const;
    void contextDestroyed(ContextImpl& context) const;
};
}  // namespace OpenMM
#endif
