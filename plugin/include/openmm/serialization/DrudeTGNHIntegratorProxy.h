#ifndef TGNH_B200_DRUDETGNH_INTEGRATOR_PROXY_H_
#define TGNH_B200_DRUDETGNH_INTEGRATOR_PROXY_H_
#include "openmm/internal/windowsExportDrude.h"
#include "openmm/serialization/SerializationProxy.h"

namespace OpenMM {

/**
 * XML proxy of DrudeTGNHIntegrator.  Type name and the version-1 properties are the reference's
 * (/root/reference/serialization/src/DrudeTGNHIntegratorProxy.cpp:40-67).  Version 2 additionally persists what
 * version 1 silently drops (maxDrudeDistance, useCOMTempGroup, the temperature-group tables); both versions load.
 */
class OPENMM_EXPORT_DRUDE DrudeTGNHIntegratorProxy : public SerializationProxy {
public:
    DrudeTGNHIntegratorProxy();
    void serialize(const void* object, SerializationNode& node) const;
    void* deserialize(const SerializationNode& node) const;
    /** 1 (default) = the reference's format, byte for byte: files interchange with the upstream plugin; 2 = complete (also
     *  maxDrudeDistance, useCOMTempGroup and the temperature groups, which version 1 loses), not readable by the upstream proxy */
    static int writeVersion;
};

}  // namespace OpenMM

#endif
