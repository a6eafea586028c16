// Host side of the API layer; behaviour follows /root/reference/openmmapi/src/DrudeTGNHIntegrator.cpp:47-194
// (validation, default temperature group, residue tables from ContextImpl::getMolecules, step loop) without its
// unconditional std::cout chatter.
#include "openmm/DrudeTGNHIntegrator.h"

#include "openmm/Context.h"
#include "openmm/DrudeForce.h"
#include "openmm/DrudeTGNHKernelExtensions.h"
#include "openmm/DrudeTGNHKernels.h"
#include "openmm/OpenMMException.h"
#include "openmm/System.h"
#include "openmm/internal/AssertionUtilities.h"
#include "openmm/internal/ContextImpl.h"

using namespace OpenMM;

DrudeTGNHIntegrator::DrudeTGNHIntegrator(double temperature, double couplingTime, double drudeTemperature, double drudeCouplingTime, double stepSize,
                                         int drudeStepsPerRealStep, int numNHChains, bool useDrudeNHChains, bool useCOMTempGroup)
    : temperature(temperature), couplingTime(couplingTime), drudeTemperature(drudeTemperature), drudeCouplingTime(drudeCouplingTime),
      maxDrudeDistance(0.0), drudeStepsPerRealStep(drudeStepsPerRealStep), numNHChains(numNHChains), useDrudeNHChains(useDrudeNHChains),
      useCOMTempGroup(useCOMTempGroup), isKESumValid(false), carryKineticEnergies(false), deferScaling(false), extensions(NULL) {
    setStepSize(stepSize);
    setConstraintTolerance(1e-5);
}

int DrudeTGNHIntegrator::addTempGroup() {
    tempGroups.push_back((int)tempGroups.size());
    return (int)tempGroups.size() - 1;
}

int DrudeTGNHIntegrator::addParticleTempGroup(int tempGroup) {
    ASSERT_VALID_INDEX(tempGroup, tempGroups);
    particleTempGroup.push_back(tempGroup);
    return (int)particleTempGroup.size() - 1;
}

void DrudeTGNHIntegrator::setParticleTempGroup(int particle, int tempGroup) {
    ASSERT_VALID_INDEX(particle, particleTempGroup);
    ASSERT_VALID_INDEX(tempGroup, tempGroups);
    particleTempGroup[particle] = tempGroup;
}

void DrudeTGNHIntegrator::getParticleTempGroup(int particle, int& tempGroup) const {
    ASSERT_VALID_INDEX(particle, particleTempGroup);
    tempGroup = particleTempGroup[particle];
}

double DrudeTGNHIntegrator::getResInvMass(int resid) const {
    ASSERT_VALID_INDEX(resid, residueInvMasses);
    return residueInvMasses[resid];
}

int DrudeTGNHIntegrator::getParticleResId(int particle) const {
    ASSERT_VALID_INDEX(particle, particleResId);
    return particleResId[particle];
}

double DrudeTGNHIntegrator::getMaxDrudeDistance() const { return maxDrudeDistance; }

void DrudeTGNHIntegrator::setMaxDrudeDistance(double distance) {
    if (distance < 0) throw OpenMMException("setMaxDrudeDistance: Distance cannot be negative");
    maxDrudeDistance = distance;
}

void DrudeTGNHIntegrator::initialize(ContextImpl& contextRef) {
    if (owner != NULL && &contextRef.getOwner() != owner) throw OpenMMException("This Integrator is already bound to a context");
    const System& system = contextRef.getSystem();
    const DrudeForce* drude = NULL;
    for (int i = 0; i < system.getNumForces(); i++) {
        const DrudeForce* f = dynamic_cast<const DrudeForce*>(&system.getForce(i));
        if (f == NULL) continue;
        if (drude != NULL) throw OpenMMException("The System contains multiple DrudeForces");
        drude = f;
    }
    if (drude == NULL) throw OpenMMException("The System does not contain a DrudeForce");
    isKESumValid = false;

    const int numParticles = system.getNumParticles();
    if (particleTempGroup.empty()) {
        // nothing assigned: everything in one group (created if the user made none)
        if (tempGroups.empty()) tempGroups.push_back(0);
        particleTempGroup.assign(numParticles, 0);
    } else if ((int)particleTempGroup.size() != numParticles)
        throw OpenMMException("Number of particles assigned with temperature groups does not match the number of system particles");

    // residues = molecules of the Context
    const std::vector<std::vector<int> > molecules = contextRef.getMolecules();
    particleResId.assign(numParticles, -1);
    residueMasses.assign(molecules.size(), 0.0);
    for (size_t r = 0; r < molecules.size(); r++)
        for (size_t j = 0; j < molecules[r].size(); j++) {
            particleResId[molecules[r][j]] = (int)r;
            residueMasses[r] += system.getParticleMass(molecules[r][j]);
        }
    residueInvMasses.resize(molecules.size());
    for (size_t r = 0; r < molecules.size(); r++) residueInvMasses[r] = 1.0 / residueMasses[r];

    context = &contextRef;
    owner = &contextRef.getOwner();
    kernel = context->getPlatform().createKernel(IntegrateDrudeTGNHStepKernel::Name(), contextRef);
    kernel.getAs<IntegrateDrudeTGNHStepKernel>().initialize(system, *this, *drude);
    // a kernel that offers the extension interface is told what this integrator guarantees (the reference's integrator never asks)
    extensions = dynamic_cast<DrudeTGNHKernelExtensions*>(&kernel.getImpl());
    if (extensions != NULL) {
        extensions->setKineticEnergyCarryOver(carryKineticEnergies);
        extensions->setDeferScaling(carryKineticEnergies && deferScaling);
    }
}

void DrudeTGNHIntegrator::cleanup() {
    extensions = NULL;
    kernel = Kernel();
}

void DrudeTGNHIntegrator::stateChanged(State::DataType changed) {
    // the step assumes valid forces on entry, so they are refreshed whenever the user touches the state
    isKESumValid = false;
    if (context != NULL) {
        if (extensions != NULL) extensions->velocitiesChanged();
        context->calcForcesAndEnergy(true, false);
    }
}

std::vector<std::string> DrudeTGNHIntegrator::getKernelNames() {
    return std::vector<std::string>(1, IntegrateDrudeTGNHStepKernel::Name());
}

double DrudeTGNHIntegrator::computeKineticEnergy() {
    return kernel.getAs<IntegrateDrudeTGNHStepKernel>().computeKineticEnergy(*context, *this, isKESumValid);
}

void DrudeTGNHIntegrator::step(int steps) {
    if (context == NULL) throw OpenMMException("This Integrator is not bound to a context!");
    IntegrateDrudeTGNHStepKernel& k = kernel.getAs<IntegrateDrudeTGNHStepKernel>();
    for (int i = 0; i < steps; ++i) {
        // openmmapi/src/DrudeTGNHIntegrator.cpp:186-189.  Forces that act in updateContextState may rewrite velocities whatever
        // they return (CMMotionRemover and AndersenThermostat return false): with such forces in the System the carry-over is
        // only as good as the user's promise (setKineticEnergyCarryOver), and a `true` always invalidates.
        if (context->updateContextState()) {
            if (extensions != NULL) extensions->velocitiesChanged();
            context->calcForcesAndEnergy(true, false);
        } else if (context->getLastForceGroups() >= 0)
            context->calcForcesAndEnergy(true, false);
        k.execute(*context, *this);
        isKESumValid = true;
    }
    if (extensions != NULL) extensions->finishSteps(*context);
}
