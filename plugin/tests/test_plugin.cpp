// CPU unit tests of the plugin's host side (API surface, validation, XML, plugin registration), shaped after the
// reference's own tests (serialization/tests/TestSerializeDrudeTGNHIntegrator.cpp, platforms/*/tests).  No GPU needed:
// the one place that would touch the device must fail loudly there ("no CPU fallback").
#include <cstring>
#include <iostream>
#include <sstream>

#include "../src/B200DrudeTGNHKernelFactory.h"
#include "../src/B200DrudeTGNHKernels.h"
#include "OpenMMDrudeTGNH.h"
#include "openmm/Context.h"
#include "openmm/DrudeTGNHKernels.h"
#include "openmm/System.h"
#include "openmm/internal/AssertionUtilities.h"
#include "openmm/serialization/DrudeTGNHIntegratorProxy.h"
#include "openmm/serialization/XmlSerializer.h"

using namespace OpenMM;
using namespace std;

extern "C" void registerKernelFactories();

#define EXPECT_THROW(stmt, text)                                                                      \
    {                                                                                                 \
        bool threw = false;                                                                           \
        try { stmt; } catch (const OpenMMException& e) { threw = true; if (!strstr(e.what(), text)) throwException(__FILE__, __LINE__, string("wrong message: ") + e.what()); } \
        if (!threw) throwException(__FILE__, __LINE__, "expected an exception");                      \
    }

static void testConstructorAndSetters() {
    DrudeTGNHIntegrator integ(300.0, 0.1, 1.0, 0.005, 0.001);
    ASSERT_EQUAL(300.0, integ.getTemperature()); ASSERT_EQUAL(0.1, integ.getCouplingTime());
    ASSERT_EQUAL(1.0, integ.getDrudeTemperature()); ASSERT_EQUAL(0.005, integ.getDrudeCouplingTime());
    ASSERT_EQUAL(0.001, integ.getStepSize()); ASSERT_EQUAL(20, integ.getDrudeStepsPerRealStep());
    ASSERT_EQUAL(1, integ.getNumNHChains()); ASSERT_EQUAL(0, integ.getUseDrudeNHChains()); ASSERT(integ.getUseCOMTempGroup());
    ASSERT_EQUAL(0.0, integ.getMaxDrudeDistance()); ASSERT_EQUAL(1e-5, integ.getConstraintTolerance());
    integ.setMaxDrudeDistance(0.02); ASSERT_EQUAL(0.02, integ.getMaxDrudeDistance());
    EXPECT_THROW(integ.setMaxDrudeDistance(-1), "Distance cannot be negative");
    integ.setNumNHChains(3); integ.setUseDrudeNHChains(1); integ.setUseCOMTempGroup(0); integ.setDrudeStepsPerRealStep(10);
    ASSERT_EQUAL(3, integ.getNumNHChains()); ASSERT_EQUAL(1, integ.getUseDrudeNHChains()); ASSERT(!integ.getUseCOMTempGroup());
    ASSERT_EQUAL(10, integ.getDrudeStepsPerRealStep());
    EXPECT_THROW(integ.step(1), "not bound to a context");
}

static void testTempGroups() {
    DrudeTGNHIntegrator integ(300.0, 0.1, 1.0, 0.005, 0.001);
    ASSERT_EQUAL(0, integ.getNumTempGroups());
    EXPECT_THROW(integ.addParticleTempGroup(0), "Index out of range");    // a group must exist first
    ASSERT_EQUAL(0, integ.addTempGroup());                                   // the FIRST group has index 0
    ASSERT_EQUAL(1, integ.addTempGroup());
    ASSERT_EQUAL(0, integ.addParticleTempGroup(1));
    ASSERT_EQUAL(1, integ.addParticleTempGroup(0));
    EXPECT_THROW(integ.addParticleTempGroup(2), "Index out of range");
    int tg = -1;
    integ.getParticleTempGroup(0, tg); ASSERT_EQUAL(1, tg);
    integ.setParticleTempGroup(0, 0); integ.getParticleTempGroup(0, tg); ASSERT_EQUAL(0, tg);
    EXPECT_THROW(integ.setParticleTempGroup(5, 0), "Index out of range");
    EXPECT_THROW(integ.getParticleTempGroup(2, tg), "Index out of range");
}

static void testSerialization() {
    // the reference's testSerialization (TestSerializeDrudeTGNHIntegrator.cpp:45-67)
    DrudeTGNHIntegrator integ1(301.1, 0.1, 10.5, 0.005, 0.001);
    stringstream buffer;
    XmlSerializer::serialize<DrudeTGNHIntegrator>(&integ1, "Integrator", buffer);
    DrudeTGNHIntegrator* copy = XmlSerializer::deserialize<DrudeTGNHIntegrator>(buffer);
    DrudeTGNHIntegrator& integ2 = *copy;
    ASSERT_EQUAL(integ1.getTemperature(), integ2.getTemperature());
    ASSERT_EQUAL(integ1.getCouplingTime(), integ2.getCouplingTime());
    ASSERT_EQUAL(integ1.getDrudeTemperature(), integ2.getDrudeTemperature());
    ASSERT_EQUAL(integ1.getDrudeCouplingTime(), integ2.getDrudeCouplingTime());
    ASSERT_EQUAL(integ1.getDrudeStepsPerRealStep(), integ2.getDrudeStepsPerRealStep());
    ASSERT_EQUAL(integ1.getNumNHChains(), integ2.getNumNHChains());
    ASSERT_EQUAL(integ1.getUseDrudeNHChains(), integ2.getUseDrudeNHChains());
    ASSERT_EQUAL(integ1.getConstraintTolerance(), integ2.getConstraintTolerance());
    delete copy;

    // version 2 keeps what version 1 drops
    DrudeTGNHIntegrator a(300.0, 0.1, 1.0, 0.005, 0.001, 10, 3, true, false);
    a.setMaxDrudeDistance(0.02);
    a.addTempGroup(); a.addTempGroup();
    a.addParticleTempGroup(0); a.addParticleTempGroup(1); a.addParticleTempGroup(1);
    stringstream b2;
    DrudeTGNHIntegratorProxy::writeVersion = 2;
    XmlSerializer::serialize<DrudeTGNHIntegrator>(&a, "Integrator", b2);
    DrudeTGNHIntegratorProxy::writeVersion = 1;
    DrudeTGNHIntegrator* c2 = XmlSerializer::deserialize<DrudeTGNHIntegrator>(b2);
    ASSERT_EQUAL(0.02, c2->getMaxDrudeDistance()); ASSERT(!c2->getUseCOMTempGroup()); ASSERT_EQUAL(2, c2->getNumTempGroups());
    int tg; c2->getParticleTempGroup(2, tg); ASSERT_EQUAL(1, tg); c2->getParticleTempGroup(0, tg); ASSERT_EQUAL(0, tg);
    ASSERT_EQUAL(3, c2->getNumNHChains()); ASSERT_EQUAL(1, c2->getUseDrudeNHChains());
    delete c2;

    // a file written by the REFERENCE's proxy (version 1; text captured from oracle/_ref) loads
    const char* refXml = "<?xml version=\"1.0\" ?>\n<Integrator constraintTolerance=\"1.0000000000000001e-05\" couplingTime=\"0.10000000000000001\" "
        "drudeCouplingTime=\"0.0050000000000000001\" drudeStepsPerRealStep=\"20\" drudeTemperature=\"10.5\" numNHChains=\"1\" stepSize=\"0.001\" "
        "temperature=\"301.10000000000002\" type=\"DrudeTGNHIntegrator\" useDrudeNHChains=\"0\" version=\"1\"/>\n";
    stringstream b3(refXml);
    DrudeTGNHIntegrator* c3 = XmlSerializer::deserialize<DrudeTGNHIntegrator>(b3);
    ASSERT_EQUAL(301.1, c3->getTemperature()); ASSERT_EQUAL(10.5, c3->getDrudeTemperature()); ASSERT_EQUAL(20, c3->getDrudeStepsPerRealStep());
    delete c3;

    // the default writing mode is version 1: exactly the reference's nine properties, nothing else
    ASSERT_EQUAL(1, DrudeTGNHIntegratorProxy::writeVersion);
    stringstream b4;
    XmlSerializer::serialize<DrudeTGNHIntegrator>(&integ1, "Integrator", b4);
    ASSERT(b4.str() == string(refXml));
    stringstream b5("<?xml version=\"1.0\" ?>\n<Integrator type=\"DrudeTGNHIntegrator\" version=\"3\"/>\n");
    EXPECT_THROW(XmlSerializer::deserialize<DrudeTGNHIntegrator>(b5), "Unsupported version number");
}

/** a "CUDA" platform without any device behind it: enough to test registration and the loud failure */
class FakeCudaPlatform : public Platform, public TgnhDeviceAccess {
public:
    const string& getName() const { static const string n = "CUDA"; return n; }
    void contextCreated(ContextImpl& c, const map<string, string>&) const { c.setPlatformData(static_cast<TgnhDeviceAccess*>(const_cast<FakeCudaPlatform*>(this))); }
    TgnhDeviceView view() { TgnhDeviceView v = {NULL, NULL, NULL, NULL, 32, TGNH_FORCE_I64_SOA, NULL, 0}; return v; }
    void advanceTime(double) {}
};

static void testIntegratorRequiresOneDrudeForce() {
    FakeCudaPlatform* platform = new FakeCudaPlatform();
    Platform::registerPlatform(platform);
    registerKernelFactories();                                              // the plugin entry point finds "CUDA" and registers the factory
    vector<string> names(1, IntegrateDrudeTGNHStepKernel::Name());
    ASSERT(platform->supportsKernels(names));
    {
        System system; system.addParticle(1.0);
        DrudeTGNHIntegrator integ(300.0, 0.1, 1.0, 0.005, 0.001);
        EXPECT_THROW(Context context(system, integ, *platform), "does not contain a DrudeForce");
    }
    {
        System system; system.addParticle(1.0); system.addForce(new DrudeForce()); system.addForce(new DrudeForce());
        DrudeTGNHIntegrator integ(300.0, 0.1, 1.0, 0.005, 0.001);
        EXPECT_THROW(Context context(system, integ, *platform), "multiple DrudeForces");
    }
    {
        System system; system.addParticle(1.0); system.addParticle(0.1); system.addForce(new DrudeForce());
        DrudeTGNHIntegrator integ(300.0, 0.1, 1.0, 0.005, 0.001);
        integ.addTempGroup(); integ.addParticleTempGroup(0);             // 1 of 2 particles assigned
        EXPECT_THROW(Context context(system, integ, *platform), "does not match the number of system particles");
    }
    {
        // a valid system: on a machine without a B200 the kernel must refuse loudly; there is no CPU path
        System system; system.addParticle(1.0); system.addParticle(0.1);
        DrudeForce* drude = new DrudeForce(); drude->addParticle(1, 0, -1, -1, -1, 0.1, 0.001, 1, 1); system.addForce(drude);
        DrudeTGNHIntegrator integ(300.0, 0.1, 10.0, 0.005, 0.003, 20, 2, false);
        try {
            Context context(system, integ, *platform);
            ASSERT_EQUAL(1, integ.getNumResidues());                        // a GPU is present: initialisation succeeded
            ASSERT_EQUAL_TOL(1.0 / 1.1, integ.getResInvMass(0), 1e-12);
        } catch (const OpenMMException& e) {
            if (!strstr(e.what(), "no CPU fallback") && !strstr(e.what(), "sm_100a")) throw;
        }
    }
    {
        Kernel k;
        bool threw = false;
        try { B200DrudeTGNHKernelFactory f; ContextImpl* none = NULL; f.createKernelImpl("SomethingElse", *platform, *none); } catch (const OpenMMException& e) { threw = strstr(e.what(), "illegal kernel name") != NULL; }
        ASSERT(threw);
    }
}

int main() {
    try {
        testConstructorAndSetters();
        testTempGroups();
        testSerialization();
        testIntegratorRequiresOneDrudeForce();
    } catch (const exception& e) {
        cout << "exception: " << e.what() << endl;
        cout << "FAIL - ERROR.  Test failed." << endl;
        return 1;
    }
    cout << "Done" << endl;
    return 0;
}
