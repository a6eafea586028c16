#ifndef TGNH_B200_DRUDETGNH_KERNELS_H_
#define TGNH_B200_DRUDETGNH_KERNELS_H_
/*
 * The kernel interface DrudeTGNHIntegrator drives; same name, constructor and three pure virtuals as the reference
 * (/root/reference/openmmapi/include/openmm/DrudeTGNHKernels.h:48-74), so a platform plugin written for the
 * reference still satisfies it.  The two hooks at the end are additions with default bodies.
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
    /** The state was modified outside the integrator (Context::setVelocities ...): drop cached kinetic energies. */
    virtual void stateChanged() {}
    /** step(n) is about to return: make the device state what a reader of the Context expects. */
    virtual void finishSteps(ContextImpl& context) {}
};

}  // namespace OpenMM

#endif
