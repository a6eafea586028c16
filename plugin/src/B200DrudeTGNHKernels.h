#ifndef TGNH_B200_KERNELS_H_
#define TGNH_B200_KERNELS_H_
/*
 * IntegrateDrudeTGNHStepKernel on B200: the KernelImpl OpenMM's CUDA platform instantiates for
 * "IntegrateDrudeTGNHStep".  It replaces CudaIntegrateDrudeTGNHStepKernel
 * (/root/reference/platforms/cuda/src/CudaDrudeTGNHKernels.{h,cpp}): no runtime-compiled kernel strings, no host
 * chain, no per-step D2H/H2D — every call below forwards to the C-ABI of libtgnh.so (include/tgnh.h).
 */
#include <vector>

#include "../../include/tgnh.h"
#include "openmm/DrudeTGNHKernels.h"

namespace OpenMM {

/** Device arrays of the platform the kernel runs on, in OpenMM's CUDA layouts. */
struct TgnhDeviceView {
    void* velm;          // float4 (single) / double4 (mixed) [paddedNumAtoms]   cu.getVelm()
    void* posq;          // float4 (single, mixed) / double4 (double) [paddedNumAtoms]   cu.getPosq()
    const void* force;   // SoA [3][paddedNumAtoms]  cu.getForce()
    void* posDelta;      // float4 / double4 [paddedNumAtoms]   integration.getPosDelta() (constrained systems only, may be NULL otherwise)
    int paddedNumAtoms;  // cu.getPaddedNumAtoms()
    int forceFormat;     // TGNH_FORCE_I64_SOA for OpenMM's fixed-point buffer
    void* stream;        // cudaStream_t the platform launches on
    int device;          // CUDA device ordinal
    int precision;       // TGNH_PRECISION_SINGLE / _MIXED / _DOUBLE   cu.getUseMixedPrecision(), cu.getUseDoublePrecision()
    void* posqCorrection;  // float4[paddedNumAtoms], mixed only   cu.getPosqCorrection()
};

/** How the kernel reaches the platform: implemented over CudaContext with a real OpenMM, over the shim platform in tests. */
class TgnhDeviceAccess {
public:
    virtual ~TgnhDeviceAccess() {}
    virtual TgnhDeviceView view() = 0;
    virtual void advanceTime(double dt) = 0;     // cu.setTime / cu.setStepCount (CudaDrudeTGNHKernels.cpp:405-406)
    // OpenMM's own kernels the step calls between ours (CudaDrudeTGNHKernels.cpp:363, 377, 391); no-ops by default
    virtual void applyConstraints(double tol) {}          // integration.applyConstraints: acts on posDelta
    virtual void computeVirtualSites() {}                 // integration.computeVirtualSites
    virtual void applyVelocityConstraints(double tol) {}  // integration.applyVelocityConstraints: acts on velm
    /** Tell the platform which particles are interchangeable for this integrator (equal descriptor words, tgnh_plan_descriptors),
     *  so that its atom reordering leaves the kernels' slot-indexed tables valid.  No-op where atoms are never reordered. */
    virtual void registerForceInfo(const std::vector<unsigned int>& descriptors, const std::vector<int>& residueOf) {}
};

class B200IntegrateDrudeTGNHStepKernel : public IntegrateDrudeTGNHStepKernel {
public:
    B200IntegrateDrudeTGNHStepKernel(std::string name, const Platform& platform, TgnhDeviceAccess& device)
        : IntegrateDrudeTGNHStepKernel(name, platform), device(device), handle(NULL), deferScale(false), recomputeKE(false), constrained(false) {}
    ~B200IntegrateDrudeTGNHStepKernel();
    void initialize(const System& system, const DrudeTGNHIntegrator& integrator, const DrudeForce& force);
    void execute(ContextImpl& context, const DrudeTGNHIntegrator& integrator);
    double computeKineticEnergy(ContextImpl& context, const DrudeTGNHIntegrator& integrator, bool isKESumValid);
    void stateChanged();
    void finishSteps(ContextImpl& context);
    /** thermostat state for checkpointing (the reference keeps it in host vectors and never saves it) */
    void getChainState(std::vector<double>& eta, std::vector<double>& etaDot, std::vector<double>& etaDotDot);
    void setChainState(const std::vector<double>& eta, const std::vector<double>& etaDot, const std::vector<double>& etaDotDot);
    /** Leave the second half-step's scaling pending between the steps of one step(n) call.  Only safe when nothing else
     *  (barostat, CMMotionRemover, reporters) touches velocities between steps; off by default. */
    void setDeferScaling(bool on) { deferScale = on; }
    /** Recompute the kinetic energies from velm at the start of every step, as the reference does, instead of carrying
     *  them over from the end of the previous step.  Needed only when something rewrites velocities between steps without
     *  going through Context::setVelocities (an AndersenThermostat; CMMotionRemover's correction is at rounding level once
     *  the total momentum has been removed).  Costs one extra pass over velm per step; off by default. */
    void setRecomputeKineticEnergies(bool on) { recomputeKE = on; }
private:
    void check(int rc) const;
    TgnhDeviceAccess& device;
    tgnh_handle* handle;
    bool deferScale;
    bool recomputeKE;
    bool constrained;    // the System has constraints: the split call sequence leaves room for OpenMM's constraint kernels
};

}  // namespace OpenMM

#endif
