// TEST INFRASTRUCTURE — driver around the reference's OWN sources (compiled unmodified from /root/reference,
// see oracle/Makefile target _ref) built against the OpenMM API shim (shim/).  Exposes a small C API so that
// tests can run the real ReferenceIntegrateDrudeTGNHStepKernel / DrudeTGNHIntegrator / serialization proxy and
// compare the oracle restatement (tgnh_oracle.c, layer TGNH_ORACLE_REF) against them.
//
// What is real here: openmmapi/src/DrudeTGNHIntegrator.cpp, platforms/reference/src/ReferenceDrudeTGNHKernels.cpp,
// platforms/reference/src/ReferenceDrudeTGNHKernelFactory.cpp, serialization/src/*.cpp.
// What is the shim: Context / System / Platform plumbing, constraints and virtual sites (no-ops), and the forces
// (fixed, or the isotropic Drude spring the reference tests use as synthetic force).
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "OpenMMDrudeTGNH.h"
#include "ReferencePlatform.h"
#include "openmm/CMMotionRemover.h"
#include "openmm/Context.h"
#include "openmm/System.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/serialization/XmlSerializer.h"

using namespace OpenMM;

extern "C" void registerDrudeTGNHReferenceKernelFactories();
extern "C" void registerDrudeTGNHSerializationProxies();

namespace {
struct Quiet {      // the reference prints unconditionally during initialize
    std::streambuf* old;
    std::ostringstream sink;
    Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~Quiet() { std::cout.rdbuf(old); }
};
std::string g_error;

struct RefSim {
    System system;
    DrudeTGNHIntegrator* integrator;
    Context* context;
    std::vector<int> pairD, pairP;
    std::vector<double> kSpring;
    std::vector<Vec3> extForce;
    int forceModel;
    RefSim() : integrator(NULL), context(NULL), forceModel(0) {}
    ~RefSim() { delete context; delete integrator; }
};

void ensure_platform() {
    static bool done = false;
    if (done) return;
    Platform::registerPlatform(new ReferencePlatform());
    registerDrudeTGNHReferenceKernelFactories();
    registerDrudeTGNHSerializationProxies();
    done = true;
}
}  // namespace

extern "C" {

const char* ref_last_error() { return g_error.c_str(); }

// residues are declared through bonds between consecutive particles of equal res_id
void* ref_create(int n, const double* masses, int npairs, const int* pairDrude, const int* pairParent, const int* resId,
                 const int* tempGroup, int numTempGroups, double temperature, double couplingTime, double drudeTemperature,
                 double drudeCouplingTime, double stepSize, int drudeSteps, int numNHChains, int useDrudeNHChains, int useCOMTempGroup,
                 double maxDrudeDistance, int hasCMMotionRemover, int forceModel, const double* kSpring) {
    try {
        Quiet q;
        ensure_platform();
        RefSim* s = new RefSim();
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
        s->integrator = new DrudeTGNHIntegrator(temperature, couplingTime, drudeTemperature, drudeCouplingTime, stepSize, drudeSteps, numNHChains,
                                                useDrudeNHChains != 0, useCOMTempGroup != 0);
        s->integrator->setMaxDrudeDistance(maxDrudeDistance);
        for (int g = 0; g < numTempGroups; g++) s->integrator->addTempGroup();
        if (tempGroup) for (int i = 0; i < n; i++) s->integrator->addParticleTempGroup(tempGroup[i]);
        s->forceModel = forceModel;
        s->extForce.assign(n, Vec3());
        s->context = new Context(s->system, *s->integrator, Platform::getPlatformByName("Reference"));
        RefSim* sp = s;
        s->context->getImpl().shimSetForceModel([sp](const std::vector<Vec3>& pos, std::vector<Vec3>& f) {
            if (sp->forceModel == 0) return;                       // fixed forces: leave what the caller installed
            for (size_t i = 0; i < f.size(); i++) f[i] = sp->extForce[i];
            for (size_t i = 0; i < sp->pairD.size(); i++) {        // isotropic Drude spring, as in the oracle's harmonic model
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

void ref_destroy(void* h) { Quiet q; delete (RefSim*)h; }

int ref_num_residues(void* h) { return ((RefSim*)h)->integrator->getNumResidues(); }

// pos / vel / force: [n][3] doubles, updated in place.  `force` must be valid for `pos` on entry.
int ref_step(void* h, double* pos, double* vel, double* force, int nsteps, const double* extForce) {
    RefSim* s = (RefSim*)h;
    try {
        Quiet q;
        const int n = s->system.getNumParticles();
        ContextImpl& impl = s->context->getImpl();
        if (extForce) for (int i = 0; i < n; i++) s->extForce[i] = Vec3(extForce[3 * i], extForce[3 * i + 1], extForce[3 * i + 2]);
        for (int i = 0; i < n; i++) {
            impl.shimPositions()[i] = Vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
            impl.shimVelocities()[i] = Vec3(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
            impl.shimForces()[i] = Vec3(force[3 * i], force[3 * i + 1], force[3 * i + 2]);
        }
        s->integrator->step(nsteps);
        for (int i = 0; i < n; i++)
            for (int c = 0; c < 3; c++) {
                pos[3 * i + c] = impl.shimPositions()[i][c];
                vel[3 * i + c] = impl.shimVelocities()[i][c];
                force[3 * i + c] = impl.shimForces()[i][c];
            }
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

// XML round trip through the reference's own proxy (serialization/src/DrudeTGNHIntegratorProxy.cpp); out: 8 getters of the copy
int ref_serialization_roundtrip(double temperature, double couplingTime, double drudeTemperature, double drudeCouplingTime, double stepSize,
                                int drudeSteps, int numNHChains, int useDrudeNHChains, double constraintTol, double* out, char* xml, int xmlCap) {
    try {
        ensure_platform();
        DrudeTGNHIntegrator a(temperature, couplingTime, drudeTemperature, drudeCouplingTime, stepSize, drudeSteps, numNHChains, useDrudeNHChains != 0);
        a.setConstraintTolerance(constraintTol);
        std::stringstream buffer;
        XmlSerializer::serialize<DrudeTGNHIntegrator>(&a, "Integrator", buffer);
        if (xml) { strncpy(xml, buffer.str().c_str(), xmlCap - 1); xml[xmlCap - 1] = 0; }
        DrudeTGNHIntegrator* b = XmlSerializer::deserialize<DrudeTGNHIntegrator>(buffer);
        out[0] = b->getTemperature(); out[1] = b->getCouplingTime(); out[2] = b->getDrudeTemperature(); out[3] = b->getDrudeCouplingTime();
        out[4] = b->getStepSize(); out[5] = b->getDrudeStepsPerRealStep(); out[6] = b->getNumNHChains(); out[7] = b->getUseDrudeNHChains();
        out[8] = b->getConstraintTolerance();
        delete b;
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

}  // extern "C"
