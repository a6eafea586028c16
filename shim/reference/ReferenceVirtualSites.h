#ifndef SHIM_REFERENCE_VIRTUAL_SITES_H_
#define SHIM_REFERENCE_VIRTUAL_SITES_H_
#include <vector>
#include "RealVec.h"
#include "openmm/System.h"
namespace OpenMM {
class ReferenceVirtualSites {
public:
    static void computePositions(const System&, std::vector<RealVec>&) {}
};
}
#endif
