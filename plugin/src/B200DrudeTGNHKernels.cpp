#include "B200DrudeTGNHKernels.h"

#include <typeinfo>

#include "openmm/CMMotionRemover.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"

using namespace OpenMM;

void B200IntegrateDrudeTGNHStepKernel::check(int rc) const {
    if (rc != TGNH_OK) throw OpenMMException(std::string("DrudeTGNH (libtgnh): ") + tgnh_last_error());
}

B200IntegrateDrudeTGNHStepKernel::~B200IntegrateDrudeTGNHStepKernel() {
    tgnh_destroy(handle);
    if (ownsDevice) delete &device;
}

// CudaIntegrateDrudeTGNHStepKernel::initialize (CudaDrudeTGNHKernels.cpp:75-282): gather the tables, hand them to tgnh_create
void B200IntegrateDrudeTGNHStepKernel::initialize(const System& system, const DrudeTGNHIntegrator& integrator, const DrudeForce& force) {
    const int n = system.getNumParticles();
    std::vector<double> masses(n);
    std::vector<int32_t> tempGroup(n), resId(n), pairDrude, pairParent, consA, consB;
    for (int i = 0; i < n; i++) {
        masses[i] = system.getParticleMass(i);
        int tg;
        integrator.getParticleTempGroup(i, tg);
        tempGroup[i] = tg;
        resId[i] = integrator.getParticleResId(i);
    }
    for (int i = 0; i < force.getNumParticles(); i++) {
        int p, p1, p2, p3, p4;
        double charge, polarizability, aniso12, aniso34;
        force.getParticleParameters(i, p, p1, p2, p3, p4, charge, polarizability, aniso12, aniso34);
        pairDrude.push_back(p);
        pairParent.push_back(p1);
    }
    for (int i = 0; i < system.getNumConstraints(); i++) {
        int a, b;
        double d;
        system.getConstraintParameters(i, a, b, d);
        consA.push_back(a);
        consB.push_back(b);
    }
    bool hasCMMotionRemover = false;
    for (int i = 0; i < system.getNumForces(); i++)
        if (dynamic_cast<const CMMotionRemover*>(&system.getForce(i)) != NULL) hasCMMotionRemover = true;

    const TgnhDeviceView dv = device.view();
    tgnh_params p = {};
    p.num_particles = n;
    p.padded_num_particles = dv.paddedNumAtoms;
    p.num_pairs = (int32_t)pairDrude.size();
    p.num_residues = integrator.getNumResidues();
    p.num_temp_groups = integrator.getNumTempGroups();
    p.num_constraints = (int32_t)consA.size();
    p.num_nh_chains = integrator.getNumNHChains();
    p.drude_steps_per_real_step = integrator.getDrudeStepsPerRealStep();
    p.use_drude_nh_chains = integrator.getUseDrudeNHChains();
    p.use_com_temp_group = integrator.getUseCOMTempGroup();
    p.has_cm_motion_remover = hasCMMotionRemover;
    p.force_format = dv.forceFormat;
    p.device = dv.device;
    p.precision = dv.precision;
    p.temperature = integrator.getTemperature();
    p.coupling_time = integrator.getCouplingTime();
    p.drude_temperature = integrator.getDrudeTemperature();
    p.drude_coupling_time = integrator.getDrudeCouplingTime();
    p.step_size = integrator.getStepSize();
    p.max_drude_distance = integrator.getMaxDrudeDistance();
    p.masses = masses.data();
    p.pair_drude = pairDrude.data();
    p.pair_parent = pairParent.data();
    p.particle_temp_group = tempGroup.data();
    p.particle_res_id = resId.data();
    p.constraint_p = consA.data();
    p.constraint_p1 = consB.data();
    constrained = system.getNumConstraints() > 0;
    // atom reordering (cu.reorderAtoms): only molecules whose per-particle tables agree may trade places.  The CudaForceInfo
    // has to be registered before OpenMM finishes setting up its contexts (cu.addForce, then initializeContexts, :76); the
    // descriptor words come from the host-side plan, which needs no device.
    std::vector<unsigned int> descriptors(n);
    check(tgnh_plan_descriptors(&p, descriptors.data()));
    device.registerForceInfo(descriptors, std::vector<int>(resId.begin(), resId.end()));
    device.initializeContexts(system);                                                                  // :76
    tgnh_destroy(handle);
    handle = NULL;
    check(tgnh_create(&p, &handle));
}

// CudaIntegrateDrudeTGNHStepKernel::execute (CudaDrudeTGNHKernels.cpp:284-408)
void B200IntegrateDrudeTGNHStepKernel::execute(ContextImpl& context, const DrudeTGNHIntegrator& integrator) {
    const TgnhDeviceView dv = device.view();
    const double tol = integrator.getConstraintTolerance();
    if (dv.precision == TGNH_PRECISION_MIXED) check(tgnh_set_posq_correction(handle, dv.posqCorrection));
    // Reference semantics unless the integrator has announced that it reports every foreign velocity change: the kinetic
    // energies are reduced from velm at the start of the step (:336, :471-490).  ContextImpl::updateContextState (barostat,
    // AndersenThermostat, CMMotionRemover, custom forces) rewrites velocities without telling anybody.
    if (!carryKE) check(tgnh_invalidate(handle));
    // the previous step's cu.reorderAtoms() moved atoms: the force buffer belongs to the old order (:344-347).  The thermostat
    // half-step that the reference runs before this test reads velocities only, so refreshing the forces first is equivalent.
    if (device.atomsWereReordered()) context.calcForcesAndEnergy(true, false);
    if (!constrained) {
        // thermostat half-step, half kick, drift, hard wall in one launch (:336-376)
        check(tgnh_half1(handle, dv.stream, dv.velm, dv.posq, dv.force));
        device.computeVirtualSites();                                                                   // :377
        context.calcForcesAndEnergy(true, false);                                                       // :380
        const TgnhDeviceView dv2 = device.view();                                                       // the force buffer may have moved
        check(tgnh_half2(handle, dv2.stream, dv2.velm, dv2.force, deferScale ? TGNH_HALF2_DEFER_SCALE : TGNH_HALF2_DEFAULT));   // :384-402
    } else {
        // same step with OpenMM's SETTLE / SHAKE / CCMA kernels where the reference calls them
        if (dv.posDelta == NULL) throw OpenMMException("DrudeTGNH: the platform provides no posDelta buffer for a constrained system");
        check(tgnh_half1_kick(handle, dv.stream, dv.velm, dv.force, dv.posDelta));                      // :336-360
        device.applyConstraints(tol);                                                                   // :363
        check(tgnh_half1_drift(handle, dv.stream, dv.velm, dv.posq, dv.posDelta));                      // :366-376
        device.computeVirtualSites();                                                                   // :377
        context.calcForcesAndEnergy(true, false);                                                       // :380
        const TgnhDeviceView dv2 = device.view();
        check(tgnh_half2(handle, dv2.stream, dv2.velm, dv2.force, TGNH_HALF2_KICK_ONLY));               // :384-388
        device.applyVelocityConstraints(tol);                                                           // :391
        check(tgnh_thermostat(handle, dv2.stream, dv2.velm, deferScale ? TGNH_HALF2_DEFER_SCALE : TGNH_HALF2_DEFAULT));   // :394-402
    }
    device.advanceTime(integrator.getStepSize());                                                       // :405-406
}

double B200IntegrateDrudeTGNHStepKernel::computeKineticEnergy(ContextImpl& context, const DrudeTGNHIntegrator& integrator, bool isKESumValid) {
    const TgnhDeviceView dv = device.view();
    double ke = 0.0;
    if (isKESumValid) {                                   // KESum cached by the last chain update (:493-497, :654-658)
        check(tgnh_kinetic_energy(handle, dv.stream, &ke));
        return ke;
    }
    check(tgnh_flush(handle, dv.stream, dv.velm));
    std::vector<double> ke2(tgnh_num_thermostats(handle));
    check(tgnh_compute_kinetic_energies(handle, dv.stream, dv.velm, ke2.data()));
    for (size_t i = 0; i < ke2.size(); i++) ke += ke2[i];
    return 0.5 * ke;
}

void B200IntegrateDrudeTGNHStepKernel::velocitiesChanged() {
    if (handle != NULL) check(tgnh_invalidate(handle));
}

std::vector<double> B200IntegrateDrudeTGNHStepKernel::getScaleFactors() {
    std::vector<double> v(tgnh_num_thermostats(handle));
    check(tgnh_get_vscale(handle, device.view().stream, v.data()));
    return v;
}

void B200IntegrateDrudeTGNHStepKernel::getThermostatParams(std::vector<double>& dof, std::vector<double>& nkbt, std::vector<double>& etaMass) {
    const int T = tgnh_num_thermostats(handle), M = tgnh_num_nh_chains(handle);
    dof.assign(T, 0.0); nkbt.assign(T, 0.0); etaMass.assign((size_t)T * M, 0.0);
    check(tgnh_get_thermostat_params(handle, dof.data(), nkbt.data(), etaMass.data()));
}

void B200IntegrateDrudeTGNHStepKernel::finishSteps(ContextImpl& context) {
    if (!deferScale) return;
    const TgnhDeviceView dv = device.view();
    check(tgnh_flush(handle, dv.stream, dv.velm));
}

void B200IntegrateDrudeTGNHStepKernel::getChainState(std::vector<double>& eta, std::vector<double>& etaDot, std::vector<double>& etaDotDot) {
    const int T = tgnh_num_thermostats(handle), M = tgnh_num_nh_chains(handle);
    eta.assign((size_t)T * M, 0.0); etaDot.assign((size_t)T * (M + 1), 0.0); etaDotDot.assign((size_t)T * M, 0.0);
    check(tgnh_get_chain_state(handle, device.view().stream, eta.data(), etaDot.data(), etaDotDot.data()));
}

void B200IntegrateDrudeTGNHStepKernel::setChainState(const std::vector<double>& eta, const std::vector<double>& etaDot, const std::vector<double>& etaDotDot) {
    check(tgnh_set_chain_state(handle, device.view().stream, eta.data(), etaDot.data(), etaDotDot.data()));
}
