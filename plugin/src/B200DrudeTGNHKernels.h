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
#include "openmm/DrudeTGNHKernelExtensions.h"
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
    /** cu.getPlatformData().initializeContexts(system) (CudaDrudeTGNHKernels.cpp:76): OpenMM finishes setting up its CudaContexts
     *  (atom order, molecule groups for the reordering) once the last kernel has registered its CudaForceInfo */
    virtual void initializeContexts(const System& system) {}
    /** cu.getAtomsWereReordered() (CudaDrudeTGNHKernels.cpp:344): the previous step's cu.reorderAtoms() moved atoms to other
     *  slots; the force buffer still belongs to the old order */
    virtual bool atomsWereReordered() { return false; }
    /** Tell the platform which particles are interchangeable for this integrator (equal descriptor words, tgnh_plan_descriptors),
     *  so that its atom reordering leaves the kernels' slot-indexed tables valid.  No-op where atoms are never reordered. */
    virtual void registerForceInfo(const std::vector<unsigned int>& descriptors, const std::vector<int>& residueOf) {}
};

class B200IntegrateDrudeTGNHStepKernel : public IntegrateDrudeTGNHStepKernel, public DrudeTGNHKernelExtensions {
public:
    /** ownsDevice: delete `device` with the kernel (the access object the real-OpenMM factory creates per kernel) */
    B200IntegrateDrudeTGNHStepKernel(std::string name, const Platform& platform, TgnhDeviceAccess& device, bool ownsDevice = false)
        : IntegrateDrudeTGNHStepKernel(name, platform), device(device), ownsDevice(ownsDevice), handle(NULL), deferScale(false), carryKE(false), constrained(false) {}
    ~B200IntegrateDrudeTGNHStepKernel();
    // ---- the reference's interface (openmmapi/include/openmm/DrudeTGNHKernels.h:48-74) ----
    void initialize(const System& system, const DrudeTGNHIntegrator& integrator, const DrudeForce& force);
    void execute(ContextImpl& context, const DrudeTGNHIntegrator& integrator);
    double computeKineticEnergy(ContextImpl& context, const DrudeTGNHIntegrator& integrator, bool isKESumValid);
    // ---- DrudeTGNHKernelExtensions ----
    void setKineticEnergyCarryOver(bool on) { carryKE = on; if (!on) deferScale = false; }
    bool getKineticEnergyCarryOver() const { return carryKE; }
    void velocitiesChanged();
    void setDeferScaling(bool on) { deferScale = on && carryKE; }
    void finishSteps(ContextImpl& context);
    void getChainState(std::vector<double>& eta, std::vector<double>& etaDot, std::vector<double>& etaDotDot);
    void setChainState(const std::vector<double>& eta, const std::vector<double>& etaDot, const std::vector<double>& etaDotDot);
    // ---- diagnostics ----
    /** vscaleFactorsVec of the most recent chain update */
    std::vector<double> getScaleFactors();
    /** dof (minus the COM share), N kT, thermostat masses [T*M] */
    void getThermostatParams(std::vector<double>& dof, std::vector<double>& nkbt, std::vector<double>& etaMass);
    long long getLaunchCount() const { return tgnh_launch_count(handle); }
    int getKernelGeneration() const { return tgnh_kernel_generation(handle); }
private:
    void check(int rc) const;
    TgnhDeviceAccess& device;
    bool ownsDevice;
    tgnh_handle* handle;
    bool deferScale;
    bool carryKE;        // the integrator announces velocity changes (DrudeTGNHKernelExtensions): energies are carried over
    bool constrained;    // the System has constraints: the split call sequence leaves room for OpenMM's constraint kernels
};

}  // namespace OpenMM

#endif
