#ifndef TGNH_B200_DRUDETGNH_KERNELS_H_
#define TGNH_B200_DRUDETGNH_KERNELS_H_
/*
 * The kernel interface DrudeTGNHIntegrator drives: the same name, constructor and exactly the three pure virtuals of the
 * reference (/root/reference/openmmapi/include/openmm/DrudeTGNHKernels.h:48-74), in the same order — the vtable of a
 * KernelImpl built against this header is the one a reference-built libOpenMMDrudeTGNH.so expects, and the other way round.
 * What this repo's kernel offers beyond it lives in a separate interface (DrudeTGNHKernelExtensions.h) that an integrator may
 * discover with dynamic_cast; the reference's integrator never looks for it and gets the reference's semantics.
 */
#include <string>

#include "openmm/DrudeForce.h"
#include "openmm/DrudeTGNHIntegrator.h"
#include "openmm/KernelImpl.h"
#include "openmm/Platform.h"
#include "openmm/System.h"

namespace OpenMM {

class IntegrateDrudeTGNHStepKernel : public KernelImpl {
public:
    static std::string Name() { return "IntegrateDrudeTGNHStep"; }
    IntegrateDrudeTGNHStepKernel(std::string name, const Platform& platform) : KernelImpl(name, platform) {}
    /** Build index tables and thermostat parameters for `system`; `force` names the Drude pairs. */
    virtual void initialize(const System& system, const DrudeTGNHIntegrator& integrator, const DrudeForce& force) = 0;
    /** Advance one step. */
    virtual void execute(ContextImpl& context, const DrudeTGNHIntegrator& integrator) = 0;
    /** Kinetic energy for State::getKineticEnergy; isKESumValid = the cached sum of the last step may be used. */
    virtual double computeKineticEnergy(ContextImpl& context, const DrudeTGNHIntegrator& integrator, bool isKESumValid) = 0;
};

}  // namespace OpenMM

#endif
