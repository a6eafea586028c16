// XML (de)serialization; version 1 == /root/reference/serialization/src/DrudeTGNHIntegratorProxy.cpp:43-67.
#include "openmm/serialization/DrudeTGNHIntegratorProxy.h"

#include <typeinfo>

#include "openmm/DrudeTGNHIntegrator.h"
#include "openmm/OpenMMException.h"
#include "openmm/serialization/SerializationNode.h"

using namespace OpenMM;

int DrudeTGNHIntegratorProxy::writeVersion = 1;

DrudeTGNHIntegratorProxy::DrudeTGNHIntegratorProxy() : SerializationProxy("DrudeTGNHIntegrator") {}

void DrudeTGNHIntegratorProxy::serialize(const void* object, SerializationNode& node) const {
    const DrudeTGNHIntegrator& integ = *reinterpret_cast<const DrudeTGNHIntegrator*>(object);
    node.setIntProperty("version", writeVersion);
    node.setDoubleProperty("stepSize", integ.getStepSize());
    node.setDoubleProperty("constraintTolerance", integ.getConstraintTolerance());
    node.setDoubleProperty("temperature", integ.getTemperature());
    node.setDoubleProperty("couplingTime", integ.getCouplingTime());
    node.setDoubleProperty("drudeTemperature", integ.getDrudeTemperature());
    node.setDoubleProperty("drudeCouplingTime", integ.getDrudeCouplingTime());
    node.setIntProperty("drudeStepsPerRealStep", integ.getDrudeStepsPerRealStep());
    node.setIntProperty("numNHChains", integ.getNumNHChains());
    node.setIntProperty("useDrudeNHChains", integ.getUseDrudeNHChains());
    if (writeVersion < 2) return;
    // what version 1 loses
    node.setDoubleProperty("maxDrudeDistance", integ.getMaxDrudeDistance());
    node.setBoolProperty("useCOMTempGroup", integ.getUseCOMTempGroup());
    node.setIntProperty("numTempGroups", integ.getNumTempGroups());
    SerializationNode& groups = node.createChildNode("ParticleTempGroups");
    for (int i = 0; i < integ.getNumParticleTempGroups(); i++) {
        int tg;
        integ.getParticleTempGroup(i, tg);
        groups.createChildNode("Particle").setIntProperty("group", tg);
    }
}

void* DrudeTGNHIntegratorProxy::deserialize(const SerializationNode& node) const {
    const int version = node.getIntProperty("version");
    if (version < 1 || version > 2) throw OpenMMException("Unsupported version number");
    DrudeTGNHIntegrator* integ = new DrudeTGNHIntegrator(
        node.getDoubleProperty("temperature"), node.getDoubleProperty("couplingTime"), node.getDoubleProperty("drudeTemperature"),
        node.getDoubleProperty("drudeCouplingTime"), node.getDoubleProperty("stepSize"), node.getIntProperty("drudeStepsPerRealStep"),
        node.getIntProperty("numNHChains"), node.getBoolProperty("useDrudeNHChains"));
    integ->setConstraintTolerance(node.getDoubleProperty("constraintTolerance"));
    if (version >= 2) {
        integ->setMaxDrudeDistance(node.getDoubleProperty("maxDrudeDistance"));
        integ->setUseCOMTempGroup(node.getBoolProperty("useCOMTempGroup"));
        for (int g = 0; g < node.getIntProperty("numTempGroups"); g++) integ->addTempGroup();
        const SerializationNode& groups = node.getChildNode("ParticleTempGroups");
        for (size_t i = 0; i < groups.getChildren().size(); i++) integ->addParticleTempGroup(groups.getChildren()[i].getIntProperty("group"));
    }
    return integ;
}

// registered when the library is loaded, like the reference (DrudeTGNHSerializationProxyRegistration.cpp:58-65)
extern "C" OPENMM_EXPORT_DRUDE void registerDrudeTGNHSerializationProxies() {
    SerializationProxy::registerProxy(typeid(DrudeTGNHIntegrator), new DrudeTGNHIntegratorProxy());
}
namespace {
struct RegisterAtLoad {
    RegisterAtLoad() { registerDrudeTGNHSerializationProxies(); }
} registerAtLoad;
}  // namespace
