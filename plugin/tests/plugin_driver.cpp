// TEST INFRASTRUCTURE — C entry points (same shape as oracle/ref_driver.cpp) that run THIS repo's plugin stack:
// plugin/src DrudeTGNHIntegrator -> B200IntegrateDrudeTGNHStepKernel -> libtgnh.so C-ABI -> sm_100a kernels, on the shim's
// "CUDA" platform.  tests/test_plugin.py compares it with the oracle.
#include <cstring>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "OpenMMDrudeTGNH.h"
#include "ShimCudaPlatform.h"
#include "../src/B200DrudeTGNHKernelFactory.h"
#include "openmm/DrudeTGNHKernels.h"
#include "openmm/CMMotionRemover.h"
#include "openmm/Context.h"
#include "openmm/System.h"
#include "openmm/serialization/XmlSerializer.h"

using namespace OpenMM;

extern "C" void registerDrudeTGNHCudaKernelFactories();

namespace {
std::string g_error;
struct Sim {
    System system;
    DrudeTGNHIntegrator* integrator;
    Context* context;
    std::vector<int> pairD, pairP;
    std::vector<double> kSpring;
    std::vector<Vec3> extForce;
    int forceModel;
    Sim() : integrator(NULL), context(NULL), forceModel(0) {}
    ~Sim() { delete context; delete integrator; }
};
std::vector<std::pair<int, int> > g_nextConstraints;   // consumed by the next ref_create
std::string g_nextPrecision = "single";                // "Precision" property of the next Context
Platform* g_platform[2] = {NULL, NULL};
Platform& platform_for(int forceFormat) {
    // one "CUDA" platform per force format; the kernel factory is registered on whichever is current
    static bool registered = false;
    if (!g_platform[forceFormat]) g_platform[forceFormat] = new ShimCudaPlatform(forceFormat);
    if (!registered) { Platform::registerPlatform(g_platform[forceFormat]); registerDrudeTGNHCudaKernelFactories(); registered = true; }
    g_platform[forceFormat]->registerKernelFactory(IntegrateDrudeTGNHStepKernel::Name(), new B200DrudeTGNHKernelFactory());
    return *g_platform[forceFormat];
}
}  // namespace

extern "C" {

const char* ref_last_error() { return g_error.c_str(); }

void* ref_create(int n, const double* masses, int npairs, const int* pairDrude, const int* pairParent, const int* resId,
                 const int* tempGroup, int numTempGroups, double temperature, double couplingTime, double drudeTemperature,
                 double drudeCouplingTime, double stepSize, int drudeSteps, int numNHChains, int useDrudeNHChains, int useCOMTempGroup,
                 double maxDrudeDistance, int hasCMMotionRemover, int forceModel, const double* kSpring) {
    try {
        Sim* s = new Sim();
        for (int i = 0; i < n; i++) s->system.addParticle(masses[i]);
        DrudeForce* drude = new DrudeForce();
        for (int i = 0; i < npairs; i++) {
            drude->addParticle(pairDrude[i], pairParent[i], -1, -1, -1, -1.0, 1.0, 1, 1);
            s->pairD.push_back(pairDrude[i]); s->pairP.push_back(pairParent[i]);
            s->kSpring.push_back(kSpring ? kSpring[i] : 0.0);
        }
        s->system.addForce(drude);
        ShimBondForce* bonds = new ShimBondForce();
        for (int i = 1; i < n; i++) if (resId[i] == resId[i - 1]) bonds->addBond(i - 1, i);
        s->system.addForce(bonds);
        if (hasCMMotionRemover) s->system.addForce(new CMMotionRemover());
        for (size_t i = 0; i < g_nextConstraints.size(); i++) s->system.addConstraint(g_nextConstraints[i].first, g_nextConstraints[i].second, 0.1);
        g_nextConstraints.clear();
        s->integrator = new DrudeTGNHIntegrator(temperature, couplingTime, drudeTemperature, drudeCouplingTime, stepSize, drudeSteps, numNHChains,
                                                useDrudeNHChains != 0, useCOMTempGroup != 0);
        s->integrator->setMaxDrudeDistance(maxDrudeDistance);
        for (int g = 0; g < numTempGroups; g++) s->integrator->addTempGroup();
        if (tempGroup) for (int i = 0; i < n; i++) s->integrator->addParticleTempGroup(tempGroup[i]);
        s->forceModel = forceModel;
        s->extForce.assign(n, Vec3());
        std::map<std::string, std::string> properties;
        properties["Precision"] = g_nextPrecision;
        g_nextPrecision = "single";
        s->context = new Context(s->system, *s->integrator, platform_for(1), properties);    // OpenMM's int64 fixed-point force buffer
        Sim* sp = s;
        ShimCudaPlatform::installForceModel(s->context->getImpl(), [sp](const std::vector<Vec3>& pos, std::vector<Vec3>& f) {
            for (size_t i = 0; i < f.size(); i++) f[i] = sp->extForce[i];
            if (sp->forceModel == 0) return;
            for (size_t i = 0; i < sp->pairD.size(); i++) {
                const int d = sp->pairD[i], p = sp->pairP[i];
                for (int c = 0; c < 3; c++) {
                    const double fc = -sp->kSpring[i] * (pos[d][c] - pos[p][c]);
                    f[d][c] += fc;
                    f[p][c] -= fc;
                }
            }
        });
        return s;
    } catch (const std::exception& e) {
        g_error = e.what();
        return NULL;
    }
}

void ref_destroy(void* h) { delete (Sim*)h; }
int ref_num_residues(void* h) { return ((Sim*)h)->integrator->getNumResidues(); }

// fixed model: `force` (or extForce when given) is THE force; harmonic: extForce + springs.  State goes through
// Context::setPositions / setVelocities, i.e. through DrudeTGNHIntegrator::stateChanged like a user script.
int ref_step(void* h, double* pos, double* vel, double* force, int nsteps, const double* extForce) {
    Sim* s = (Sim*)h;
    try {
        const int n = s->system.getNumParticles();
        const double* fixed = extForce ? extForce : (s->forceModel == 0 ? force : NULL);
        for (int i = 0; i < n; i++) s->extForce[i] = fixed ? Vec3(fixed[3 * i], fixed[3 * i + 1], fixed[3 * i + 2]) : Vec3();
        std::vector<Vec3> p(n), v(n);
        for (int i = 0; i < n; i++) { p[i] = Vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]); v[i] = Vec3(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]); }
        s->context->setPositions(p);
        s->context->setVelocities(v);
        s->integrator->step(nsteps);
        State st = s->context->getState(State::Positions | State::Velocities | State::Forces | State::Energy);
        for (int i = 0; i < n; i++)
            for (int c = 0; c < 3; c++) {
                pos[3 * i + c] = st.getPositions()[i][c];
                vel[3 * i + c] = st.getVelocities()[i][c];
                force[3 * i + c] = st.getForces()[i][c];
            }
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

// constraints of the System the next ref_create builds (the shim never applies them: they only switch the kernel to
// the split call sequence and enter the DOF bookkeeping)
void plugin_set_next_constraints(int n, const int* a, const int* b) {
    g_nextConstraints.clear();
    for (int i = 0; i < n; i++) g_nextConstraints.push_back(std::make_pair(a[i], b[i]));
}

void plugin_set_next_precision(const char* precision) { g_nextPrecision = precision; }

int plugin_constraint_calls(void* h) {
    Sim* s = (Sim*)h;
    return static_cast<ShimCudaPlatform::Data*>(static_cast<TgnhDeviceAccess*>(s->context->getImpl().getPlatformData()))->constraintCalls;
}

double plugin_kinetic_energy(void* h) {
    try { return ((Sim*)h)->context->getState(State::Energy).getKineticEnergy(); } catch (const std::exception& e) { g_error = e.what(); return -1.0; }
}

}  // extern "C"
