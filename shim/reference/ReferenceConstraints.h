#ifndef SHIM_REFERENCE_CONSTRAINTS_H_
#define SHIM_REFERENCE_CONSTRAINTS_H_
#include <vector>
#include "RealVec.h"
namespace OpenMM {
/** Integrator-only path: no constraint is ever applied (SETTLE/SHAKE/CCMA stay inside OpenMM). */
class ReferenceConstraints {
public:
    void apply(std::vector<RealVec>&, std::vector<RealVec>&, std::vector<double>&, double) {}
    void applyToVelocities(std::vector<RealVec>&, std::vector<RealVec>&, std::vector<double>&, double) {}
};
}
#endif
