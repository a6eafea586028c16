#ifndef TGNH_B200_DRUDETGNH_KERNEL_EXTENSIONS_H_
#define TGNH_B200_DRUDETGNH_KERNEL_EXTENSIONS_H_
/*
 * Optional second interface of an IntegrateDrudeTGNHStepKernel implementation (not part of the reference).  An integrator that
 * knows about it finds it with dynamic_cast<DrudeTGNHKernelExtensions*>(&kernel.getImpl()); one that does not (the reference's
 * own openmmapi/src/DrudeTGNHIntegrator.cpp) gets a kernel that behaves like the reference's: kinetic energies reduced from
 * the velocities at the start of every step, velocities final at the end of every step.
 */
#include <vector>

namespace OpenMM {

class ContextImpl;

class DrudeTGNHKernelExtensions {
public:
    virtual ~DrudeTGNHKernelExtensions() {}
    /** The caller promises to announce every change of the velocities that happens outside execute() with velocitiesChanged().
     *  The kernel may then carry the kinetic energies over from the end of one step to the thermostat half-step that begins the
     *  next (one pass over the velocities less per step).  Off by default. */
    virtual void setKineticEnergyCarryOver(bool on) = 0;
    virtual bool getKineticEnergyCarryOver() const = 0;
    /** Context::setVelocities, a barostat / thermostat / CMMotionRemover in updateContextState, ...: cached energies are dropped. */
    virtual void velocitiesChanged() = 0;
    /** Leave the second thermostat half-step's velocity scaling pending between the steps of one step(n) call (it is folded into
     *  the next step's first pass).  Only legal when nothing reads or writes velocities between those steps; finishSteps()
     *  applies what is pending.  Off by default; needs the carry-over. */
    virtual void setDeferScaling(bool on) = 0;
    /** step(n) is about to return: make the device state what a reader of the Context expects. */
    virtual void finishSteps(ContextImpl& context) = 0;
    /** thermostat state for checkpointing (the reference keeps it in host vectors and never saves it): eta [T*M],
     *  etaDot [T*(M+1)], etaDotDot [T*M], T = numTempGroups + 2 */
    virtual void getChainState(std::vector<double>& eta, std::vector<double>& etaDot, std::vector<double>& etaDotDot) = 0;
    virtual void setChainState(const std::vector<double>& eta, const std::vector<double>& etaDot, const std::vector<double>& etaDotDot) = 0;
};

}  // namespace OpenMM

#endif
