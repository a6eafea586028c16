#ifndef SHIM_REFERENCE_PLATFORM_H_
#define SHIM_REFERENCE_PLATFORM_H_
#include "openmm/shim_core.h"
#include "RealVec.h"
#include "ReferenceConstraints.h"
namespace OpenMM {
/** Host platform: its PlatformData points at the ContextImpl's host arrays, like OpenMM's ReferencePlatform. */
class ReferencePlatform : public Platform {
public:
    class PlatformData {
    public:
        PlatformData(ContextImpl& c) : time(0.0), stepCount(0), numParticles(c.getSystem().getNumParticles()),
              positions(&c.shimPositions()), velocities(&c.shimVelocities()), forces(&c.shimForces()), constraints(new ReferenceConstraints()) {}
        ~PlatformData() { delete (ReferenceConstraints*)constraints; }
        double time;
        int stepCount, numParticles;
        void *positions, *velocities, *forces, *constraints;
    };
    const std::string& getName() const { static const std::string n = "Reference"; return n; }
    void contextCreated(ContextImpl& context, const std::map<std::string, std::string>&) const { context.setPlatformData(new PlatformData(context)); }
    void contextDestroyed(ContextImpl& context) const { delete (PlatformData*)context.getPlatformData(); }
};
}
#endif
