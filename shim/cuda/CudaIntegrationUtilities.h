// shim/cuda/CudaIntegrationUtilities.h — stand-in for OpenMM 7.x's CudaIntegrationUtilities: posDelta / stepSize arrays and the
// constraint / virtual-site entry points the TGNH step calls between its own kernels.  TEST / BUILD INFRASTRUCTURE: the shim
// has no constraint solver (integrator-only path); the calls are counted so that tests can see the sequence.
#ifndef SHIM_CUDA_INTEGRATION_UTILITIES_H_
#define SHIM_CUDA_INTEGRATION_UTILITIES_H_
#include "CudaArray.h"
namespace OpenMM {
class CudaContext;
class System;
class CudaIntegrationUtilities {
public:
    CudaIntegrationUtilities(CudaContext& context, const System& system);
    ~CudaIntegrationUtilities();
    /** mixed4[paddedNumAtoms] */
    CudaArray& getPosDelta() { return *posDelta; }
    /** mixed2[1]; .y = step size */
    CudaArray& getStepSize() { return *stepSize; }
    void applyConstraints(double tol) { constraintCalls++; }
    void applyVelocityConstraints(double tol) { velocityConstraintCalls++; }
    void computeVirtualSites() { virtualSiteCalls++; }
    /** 0.5 sum m v^2 of the velocities as stored (timeShift is ignored: the TGNH path only asks for 0) */
    double computeKineticEnergy(double timeShift);
    int constraintCalls, velocityConstraintCalls, virtualSiteCalls;
private:
    CudaContext& context;
    CudaArray* posDelta;
    CudaArray* stepSize;
};
}  // namespace OpenMM
#endif
