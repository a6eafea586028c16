#ifndef TGNH_B200_DRUDETGNHINTEGRATOR_H_
#define TGNH_B200_DRUDETGNHINTEGRATOR_H_
/*
 * DrudeTGNHIntegrator — public API of the temperature-grouped dual Nose-Hoover integrator for Drude-polarizable
 * systems, kept call-for-call compatible with the reference plugin
 * (/root/reference/openmmapi/include/openmm/DrudeTGNHIntegrator.h:56-311) so that user code and the SWIG module
 * (python/drudetgnhplugin.i) bind to it unchanged.  The step itself runs in libtgnh (include/tgnh.h) behind
 * IntegrateDrudeTGNHStepKernel.
 *
 * Thermostats: one per temperature group (motion relative to the molecular centre of mass), one for the molecular
 * centres of mass (useCOMTempGroup), one for the Drude-pair internal motion at drudeTemperature.
 */
#include <string>
#include <vector>

#include "openmm/Integrator.h"
#include "openmm/Kernel.h"
#include "openmm/State.h"
#include "openmm/internal/windowsExportDrude.h"

namespace OpenMM {

class OPENMM_EXPORT_DRUDE DrudeTGNHIntegrator : public Integrator {
public:
    /**
     * @param temperature            heat-bath temperature of the real degrees of freedom (K)
     * @param couplingTime           thermostat time constant of the real degrees of freedom (ps)
     * @param drudeTemperature       heat-bath temperature of the Drude internal motion (K)
     * @param drudeCouplingTime      thermostat time constant of the Drude internal motion (ps)
     * @param stepSize               integration step (ps)
     * @param drudeStepsPerRealStep  thermostat sub-steps per step
     * @param numNHChains            Nose-Hoover chain length
     * @param useDrudeNHChains       chain the Drude thermostat too
     * @param useCOMTempGroup        thermostat molecular centres of mass separately
     */
    DrudeTGNHIntegrator(double temperature, double couplingTime, double drudeTemperature, double drudeCouplingTime, double stepSize,
                        int drudeStepsPerRealStep = 20, int numNHChains = 1, bool useDrudeNHChains = false, bool useCOMTempGroup = true);

    double getTemperature() const { return temperature; }
    void setTemperature(double temp) { temperature = temp; }
    double getCouplingTime() const { return couplingTime; }
    void setCouplingTime(double tau) { couplingTime = tau; }
    double getDrudeTemperature() const { return drudeTemperature; }
    void setDrudeTemperature(double temp) { drudeTemperature = temp; }
    double getDrudeCouplingTime() const { return drudeCouplingTime; }
    void setDrudeCouplingTime(double tau) { drudeCouplingTime = tau; }
    /** Hard-wall limit on the Drude-parent distance (nm); 0 switches the wall off. */
    double getMaxDrudeDistance() const;
    void setMaxDrudeDistance(double distance);
    void step(int steps);
    int getDrudeStepsPerRealStep() const { return drudeStepsPerRealStep; }
    void setDrudeStepsPerRealStep(int drudeSteps) { drudeStepsPerRealStep = drudeSteps; }
    int getNumNHChains() const { return numNHChains; }
    void setNumNHChains(int numChains) { numNHChains = numChains; }
    int getUseDrudeNHChains() const { return useDrudeNHChains; }
    void setUseDrudeNHChains(int useChains) { useDrudeNHChains = useChains; }
    bool getUseCOMTempGroup() const { return useCOMTempGroup; }
    void setUseCOMTempGroup(int useCOMGroup) { useCOMTempGroup = useCOMGroup; }
    /** Number of temperature groups of the real degrees of freedom. */
    int getNumTempGroups() const { return (int)tempGroups.size(); }
    /** (not in the reference) number of particles that have been given a temperature group with addParticleTempGroup */
    int getNumParticleTempGroups() const { return (int)particleTempGroup.size(); }
    /** Creates a temperature group; returns its index (the first call returns 0). */
    int addTempGroup();
    /** Assigns the next particle (in System order) to a group; returns that particle's index. */
    int addParticleTempGroup(int tempGroup);
    void setParticleTempGroup(int particle, int tempGroup);
    void getParticleTempGroup(int particle, int& tempGroup) const;
    /** Residue (molecule) tables, valid once the integrator is bound to a Context. */
    int getNumResidues() const { return (int)residueMasses.size(); }
    double getResInvMass(int resid) const;
    int getParticleResId(int particle) const;

    // ---- additions of this implementation (no counterpart in the reference; defaults reproduce its behaviour) ----
    /**
     * Let the kernel carry the kinetic energies over from the end of one step to the thermostat half-step that begins the next,
     * instead of reducing them from the velocities again (what the reference does at the start of every step).  This integrator
     * then tells the kernel about every velocity change it can see: Context::setVelocities (stateChanged) and any
     * updateContextState() that returns true.  Forces that rewrite velocities in updateContextState and return false
     * (CMMotionRemover, AndersenThermostat) are invisible to it — switch this on only for Systems without them, or where their
     * correction is known to be at rounding level.  Set before the Context is created.  Default: off.
     */
    void setKineticEnergyCarryOver(bool on) { carryKineticEnergies = on; }
    bool getKineticEnergyCarryOver() const { return carryKineticEnergies; }
    /**
     * Within one step(n) call, fold the velocity scaling that ends a step into the first pass of the next step (velocities are
     * made consistent before step(n) returns).  Needs the carry-over and a System in which nothing reads or writes velocities
     * between the steps of a step(n) call.  Set before the Context is created.  Default: off.
     */
    void setDeferScaling(bool on) { deferScaling = on; }
    bool getDeferScaling() const { return deferScaling; }

protected:
    void initialize(ContextImpl& context);
    void cleanup();
    void stateChanged(State::DataType changed);
    std::vector<std::string> getKernelNames();
    double computeKineticEnergy();

private:
    double temperature, couplingTime, drudeTemperature, drudeCouplingTime, maxDrudeDistance;
    int drudeStepsPerRealStep, numNHChains;
    bool useDrudeNHChains, useCOMTempGroup, isKESumValid;
    std::vector<int> particleTempGroup, tempGroups, particleResId;
    std::vector<double> residueMasses, residueInvMasses;
    Kernel kernel;
    bool carryKineticEnergies, deferScaling;
    class DrudeTGNHKernelExtensions* extensions;
};

}  // namespace OpenMM

#endif
