// TEST INFRASTRUCTURE — one driver, two builds (oracle/Makefile):
//
//   -DDRIVER_REFERENCE  oracle/_refcuda/librefcuda.so: the reference's OWN CUDA platform — openmmapi/src/DrudeTGNHIntegrator.cpp,
//        platforms/cuda/src/CudaDrudeTGNHKernels.cpp, CudaDrudeTGNHKernelFactory.cpp and the kernel strings of
//        platforms/cuda/src/kernels/{vectorOps,drudeTGNH}.cu, all compiled / JIT-compiled UNMODIFIED from /root/reference —
//        running on the GPU behind the CUDA-platform stand-in of shim/cuda (real device arrays in OpenMM's layouts, NVRTC
//        behind OpenMM's kernel prelude, OpenMM's launch rule).  This is the execution-level pin of the temperature-group /
//        COM-thermostat arithmetic, which exists nowhere else in the reference, and the same-box GPU baseline.
//   -DDRIVER_B200       oracle/_refcuda/libb200cuda.so: the SAME unmodified reference DrudeTGNHIntegrator and the SAME stand-in
//        CUDA platform, but "IntegrateDrudeTGNHStep" is served by this repo's plugin (plugin/src/B200DrudeTGNHKernels.cpp +
//        B200DrudeTGNHKernelFactory.cpp compiled with -DTGNH_WITH_OPENMM, i.e. the CudaContext-facing code path) over
//        libtgnh.so: the drop-in claim, executed.
//
// Both expose the same C API, so a test drives the two side by side on the same inputs.
#include <cstring>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "CudaContext.h"
#include "CudaPlatform.h"
#include "openmm/CMMotionRemover.h"
#include "openmm/Context.h"
#include "openmm/System.h"
#include "openmm/internal/ContextImpl.h"

// the thermostat state of both implementations is private: test access only
#define private public
#define protected public
#include "openmm/DrudeTGNHIntegrator.h"
#ifdef DRIVER_REFERENCE
#include "CudaDrudeTGNHKernelFactory.h"
#include "CudaDrudeTGNHKernels.h"
#else
#include "B200DrudeTGNHKernelFactory.h"
#include "B200DrudeTGNHKernels.h"
#endif
#undef private
#undef protected

using namespace OpenMM;

namespace OpenMM {
void shimCudaInstallForceModel(ContextImpl& context, ShimForceModel model, const std::vector<Vec3>* fixedForces);
}

namespace {
struct Quiet {      // the reference prints unconditionally during initialize
    std::streambuf* old;
    std::ostringstream sink;
    Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~Quiet() { std::cout.rdbuf(old); }
};
std::string g_error;

struct Sim {
    System system;
    CudaPlatform platform;              // private to this driver: no global registry involved
    DrudeTGNHIntegrator* integrator;
    Context* context;
    std::vector<int> pairD, pairP;
    std::vector<double> kSpring;
    std::vector<Vec3> extForce;
    int forceModel;
    Sim() : integrator(NULL), context(NULL), forceModel(0) {}
    ~Sim() { delete context; delete integrator; }
    CudaContext& cu() { return *static_cast<CudaPlatform::PlatformData*>(context->getImpl().getPlatformData())->contexts[0]; }
    CudaPlatform::PlatformData& pd() { return *static_cast<CudaPlatform::PlatformData*>(context->getImpl().getPlatformData()); }
};

void install_forces(Sim* s, const double* fixed) {
    ContextImpl& impl = s->context->getImpl();
    const int n = s->system.getNumParticles();
    if (s->forceModel == 0) {
        std::vector<Vec3> f(n);
        for (int i = 0; i < n; i++) f[i] = fixed ? Vec3(fixed[3 * i], fixed[3 * i + 1], fixed[3 * i + 2]) : Vec3();
        shimCudaInstallForceModel(impl, ShimForceModel(), &f);
        return;
    }
    Sim* sp = s;
    shimCudaInstallForceModel(impl, [sp](const std::vector<Vec3>& pos, std::vector<Vec3>& f) {
        for (size_t i = 0; i < f.size(); i++) f[i] = sp->extForce[i];
        for (size_t i = 0; i < sp->pairD.size(); i++) {        // isotropic Drude spring, as in the oracle's harmonic model
            const int d = sp->pairD[i], p = sp->pairP[i];
            for (int c = 0; c < 3; c++) {
                const double fc = -sp->kSpring[i] * (pos[d][c] - pos[p][c]);
                f[d][c] += fc;
                f[p][c] -= fc;
            }
        }
    }, NULL);
}
}  // namespace

extern "C" {

const char* cudadrv_last_error() { return g_error.c_str(); }
const char* cudadrv_flavour() {
#ifdef DRIVER_REFERENCE
    return "reference";
#else
    return "b200";
#endif
}

// precision: 0 single, 1 mixed, 2 double.  residues are declared through bonds between consecutive particles of equal res_id.
void* cudadrv_create(int n, const double* masses, int npairs, const int* pairDrude, const int* pairParent, const int* resId, const int* tempGroup,
                     int numTempGroups, double temperature, double couplingTime, double drudeTemperature, double drudeCouplingTime, double stepSize,
                     int drudeSteps, int numNHChains, int useDrudeNHChains, int useCOMTempGroup, double maxDrudeDistance, int hasCMMotionRemover,
                     int precision, int forceModel, const double* kSpring, int reorderInterval, int numConstraints, const int* consA, const int* consB) {
    Sim* s = NULL;
    try {
        Quiet q;
        s = new Sim();
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
        // constraints enter the DOF bookkeeping (CudaDrudeTGNHKernels.cpp:186-196); the stand-in platform applies none (it counts the calls)
        for (int i = 0; i < numConstraints; i++) s->system.addConstraint(consA[i], consB[i], 0.1);
        s->integrator = new DrudeTGNHIntegrator(temperature, couplingTime, drudeTemperature, drudeCouplingTime, stepSize, drudeSteps, numNHChains,
                                                useDrudeNHChains != 0, useCOMTempGroup != 0);
        s->integrator->setMaxDrudeDistance(maxDrudeDistance);
        for (int g = 0; g < numTempGroups; g++) s->integrator->addTempGroup();
        if (tempGroup) for (int i = 0; i < n; i++) s->integrator->addParticleTempGroup(tempGroup[i]);
        s->forceModel = forceModel;
        s->extForce.assign(n, Vec3());
#ifdef DRIVER_REFERENCE
        s->platform.registerKernelFactory(IntegrateDrudeTGNHStepKernel::Name(), new CudaDrudeTGNHKernelFactory());
#else
        s->platform.registerKernelFactory(IntegrateDrudeTGNHStepKernel::Name(), new B200DrudeTGNHKernelFactory());
#endif
        std::map<std::string, std::string> props;
        props["Precision"] = precision == 2 ? "double" : precision == 1 ? "mixed" : "single";
        s->context = new Context(s->system, *s->integrator, s->platform, props);
        s->cu().shimReorderInterval = reorderInterval;
        install_forces(s, NULL);
        return s;
    } catch (const std::exception& e) {
        g_error = e.what();
        delete s;
        return NULL;
    }
}

void cudadrv_destroy(void* h) { Quiet q; delete (Sim*)h; }
int cudadrv_num_residues(void* h) { return ((Sim*)h)->integrator->getNumResidues(); }

// pos / vel / force: [n][3] doubles in ORIGINAL particle order.  force: the fixed forces (force model 0) or the external part
// added to the Drude springs (force model 1).  Goes through Context::setPositions / setVelocities like a user script.
int cudadrv_set_state(void* h, const double* pos, const double* vel, const double* force) {
    Sim* s = (Sim*)h;
    try {
        Quiet q;
        const int n = s->system.getNumParticles();
        std::vector<Vec3> p(n), v(n);
        for (int i = 0; i < n; i++) {
            p[i] = Vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
            v[i] = Vec3(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
            if (force) s->extForce[i] = Vec3(force[3 * i], force[3 * i + 1], force[3 * i + 2]);
        }
        install_forces(s, force);
        s->context->setPositions(p);
        s->context->setVelocities(v);
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

// only the velocities, through Context::setVelocities (-> Integrator::stateChanged)
int cudadrv_set_velocities(void* h, const double* vel) {
    Sim* s = (Sim*)h;
    try {
        Quiet q;
        const int n = s->system.getNumParticles();
        s->context->getState(State::Positions | State::Velocities);      // host copies current (the shim uploads both)
        std::vector<Vec3> v(n);
        for (int i = 0; i < n; i++) v[i] = Vec3(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
        s->context->setVelocities(v);
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

int cudadrv_step(void* h, int nsteps) {
    Sim* s = (Sim*)h;
    try {
        Quiet q;
        s->integrator->step(nsteps);
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

// nsteps steps bracketed by events on the platform's stream; returns milliseconds (< 0 on error)
double cudadrv_time_steps(void* h, int nsteps) {
    Sim* s = (Sim*)h;
    try {
        Quiet q;
        s->cu().setAsCurrent();
        CUevent e0, e1;
        cuEventCreate(&e0, CU_EVENT_DEFAULT); cuEventCreate(&e1, CU_EVENT_DEFAULT);
        cuEventRecord(e0, s->cu().getCurrentStream());
        s->integrator->step(nsteps);
        cuEventRecord(e1, s->cu().getCurrentStream());
        cuEventSynchronize(e1);
        float ms = 0.f;
        cuEventElapsedTime(&ms, e0, e1);
        cuEventDestroy(e0); cuEventDestroy(e1);
        return ms;
    } catch (const std::exception& e) {
        g_error = e.what();
        return -1.0;
    }
}

int cudadrv_get_state(void* h, double* pos, double* vel, double* force, double* kineticEnergy) {
    Sim* s = (Sim*)h;
    try {
        Quiet q;
        const int n = s->system.getNumParticles();
        State st = s->context->getState(State::Positions | State::Velocities | State::Forces | (kineticEnergy ? State::Energy : 0));
        for (int i = 0; i < n; i++)
            for (int c = 0; c < 3; c++) {
                if (pos) pos[3 * i + c] = st.getPositions()[i][c];
                if (vel) vel[3 * i + c] = st.getVelocities()[i][c];
                if (force) force[3 * i + c] = st.getForces()[i][c];
            }
        if (kineticEnergy) *kineticEnergy = st.getKineticEnergy();
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

// eta [T*M], etaDot [T*(M+1)], etaDotDot [T*M], vscale [T] of the last chain update
int cudadrv_get_thermostat(void* h, double* eta, double* etaDot, double* etaDotDot, double* vscale) {
    Sim* s = (Sim*)h;
    try {
        const int T = s->integrator->getNumTempGroups() + 2, M = s->integrator->getNumNHChains();
#ifdef DRIVER_REFERENCE
        CudaIntegrateDrudeTGNHStepKernel& k = s->integrator->kernel.getAs<CudaIntegrateDrudeTGNHStepKernel>();
        for (int g = 0; g < T; g++) {
            for (int i = 0; i < M; i++) { eta[g * M + i] = k.eta[g][i]; etaDotDot[g * M + i] = k.etaDotDot[g][i]; }
            for (int i = 0; i <= M; i++) etaDot[g * (M + 1) + i] = k.etaDot[g][i];
            if (vscale) vscale[g] = k.vscaleFactorsVec[g];
        }
#else
        B200IntegrateDrudeTGNHStepKernel& k = s->integrator->kernel.getAs<B200IntegrateDrudeTGNHStepKernel>();
        std::vector<double> a, b, c;
        k.getChainState(a, b, c);
        memcpy(eta, a.data(), a.size() * 8); memcpy(etaDot, b.data(), b.size() * 8); memcpy(etaDotDot, c.data(), c.size() * 8);
        if (vscale) { std::vector<double> v = k.getScaleFactors(); memcpy(vscale, v.data(), (size_t)T * 8); }
#endif
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

// dof[T] (tempGroupDof - tempGroupRedMass is NOT kept by the reference: it keeps tempGroupDof and tempGroupNkbT), nkbt[T], etaMass[T*M]
int cudadrv_get_thermostat_params(void* h, double* nkbt, double* etaMass) {
    Sim* s = (Sim*)h;
    try {
        const int T = s->integrator->getNumTempGroups() + 2, M = s->integrator->getNumNHChains();
#ifdef DRIVER_REFERENCE
        CudaIntegrateDrudeTGNHStepKernel& k = s->integrator->kernel.getAs<CudaIntegrateDrudeTGNHStepKernel>();
        for (int g = 0; g < T; g++) {
            nkbt[g] = k.tempGroupNkbT[g];
            for (int i = 0; i < M; i++) etaMass[g * M + i] = k.etaMass[g][i];
        }
#else
        B200IntegrateDrudeTGNHStepKernel& k = s->integrator->kernel.getAs<B200IntegrateDrudeTGNHStepKernel>();
        std::vector<double> dof(T), nk(T), q((size_t)T * M);
        k.getThermostatParams(dof, nk, q);
        memcpy(nkbt, nk.data(), (size_t)T * 8); memcpy(etaMass, q.data(), q.size() * 8);
#endif
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

// out[0] kernel launches through cu.executeKernel (reference) / libtgnh launch count (b200), [1] initializeContexts calls, [2] force evaluations,
// [3] reorders performed, [4] applyConstraints calls, [5] applyVelocityConstraints calls, [6] computeVirtualSites calls, [7] step count
int cudadrv_counters(void* h, long long* out) {
    Sim* s = (Sim*)h;
    try {
        CudaContext& cu = s->cu();
#ifdef DRIVER_REFERENCE
        out[0] = cu.shimKernelLaunches;
#else
        out[0] = s->integrator->kernel.getAs<B200IntegrateDrudeTGNHStepKernel>().getLaunchCount();
#endif
        out[1] = s->pd().initializeCalls;
        out[2] = s->context->getImpl().shimForceCalls();
        out[3] = cu.shimReorderCount;
        out[4] = cu.getIntegrationUtilities().constraintCalls;
        out[5] = cu.getIntegrationUtilities().velocityConstraintCalls;
        out[6] = cu.getIntegrationUtilities().virtualSiteCalls;
        out[7] = cu.getStepCount();
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

// the kernel source the platform compiled at run time, prelude included (reference flavour; tests read the prelude)
int cudadrv_last_kernel_source(void* h, char* out, int cap) {
    Sim* s = (Sim*)h;
    const std::string& src = s->cu().shimLastSource;
    if (out && cap > 0) { strncpy(out, src.c_str(), cap - 1); out[cap - 1] = 0; }
    return (int)src.size();
}

}  // extern "C"

extern "C" {
// KESum cached by the last chain update (CudaDrudeTGNHKernels.cpp:493-497)
double cudadrv_kesum(void* h) {
    Sim* s = (Sim*)h;
    try {
#ifdef DRIVER_REFERENCE
        return s->integrator->kernel.getAs<CudaIntegrateDrudeTGNHStepKernel>().KESum;
#else
        return s->integrator->kernel.getAs<B200IntegrateDrudeTGNHStepKernel>().computeKineticEnergy(s->context->getImpl(), *s->integrator, true);
#endif
    } catch (const std::exception& e) {
        g_error = e.what();
        return -1.0;
    }
}
// b200 flavour: which kernel generation serves the handle (tgnh_kernel_generation); 0 for the reference flavour
int cudadrv_kernel_generation(void* h) {
#ifdef DRIVER_REFERENCE
    return 0;
#else
    return ((Sim*)h)->integrator->kernel.getAs<B200IntegrateDrudeTGNHStepKernel>().getKernelGeneration();
#endif
}
}
