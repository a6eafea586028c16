// `drudetgnhplugin` for environments without SWIG / OpenMM's Python layer: the same class, method names, argument
// order and Python-side defaults as the SWIG module (drudetgnhplugin.i), bound with pybind11 against whichever OpenMM
// API the C++ side was compiled with (here: the shim).  Getters return plain floats unless a `unit` module
// (openmm.unit / simtk.unit) is importable, in which case they return Quantities like the reference's module.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <map>
#include <sstream>
#include <vector>

#include "OpenMMDrudeTGNH.h"
#include "openmm/serialization/XmlSerializer.h"
#include "openmm/serialization/DrudeTGNHIntegratorProxy.h"
#ifndef TGNH_WITH_OPENMM
// Without OpenMM's own Python layer there is no Context to bind an integrator to.  `drudetgnhplugin.shim` exposes the stand-in
// classes this repo's C++ side is built against (shim/, plugin/tests/ShimCudaPlatform.h) under OpenMM's names, so that a script
// reads like example/nacl_tg.py: System, DrudeForce, Platform.getPlatformByName("CUDA"), Context(system, integrator, platform,
// {"Precision": ...}), integrator.step(n), context.getState(...).  Forces come from the shim's host force model (constant external
// forces + harmonic Drude springs); the integrator step runs on the GPU through the real plugin stack.
#include "../tests/ShimCudaPlatform.h"
#include "../src/B200DrudeTGNHKernelFactory.h"
#include "openmm/CMMotionRemover.h"
#include "openmm/Context.h"
#include "openmm/DrudeTGNHKernels.h"
#include "openmm/System.h"
extern "C" void registerDrudeTGNHCudaKernelFactories();
#endif

namespace py = pybind11;
using namespace OpenMM;

static py::object with_unit(double v, const char* unitName) {
    static py::object unit = []() -> py::object {
        for (const char* mod : {"openmm.unit", "simtk.unit"}) {
            try { return py::module_::import(mod); } catch (py::error_already_set&) {}
        }
        return py::none();
    }();
    if (unit.is_none()) return py::float_(v);
    return unit.attr("Quantity")(v, unit.attr(unitName));
}

#ifndef TGNH_WITH_OPENMM
namespace {
typedef py::array_t<double, py::array::c_style | py::array::forcecast> Arr;
std::vector<Vec3> to_vec3(const Arr& a, int n) {
    if (a.ndim() != 2 || a.shape(0) != n || a.shape(1) != 3) throw OpenMMException("expected an array of shape (numParticles, 3)");
    std::vector<Vec3> v(n);
    auto r = a.unchecked<2>();
    for (int i = 0; i < n; i++) v[i] = Vec3(r(i, 0), r(i, 1), r(i, 2));
    return v;
}
Arr from_vec3(const std::vector<Vec3>& v) {
    Arr a({(py::ssize_t)v.size(), (py::ssize_t)3});
    auto w = a.mutable_unchecked<2>();
    for (size_t i = 0; i < v.size(); i++) for (int c = 0; c < 3; c++) w(i, c) = v[i][c];
    return a;
}
// the shim's host force model: constant external forces plus harmonic springs on the Drude pairs
struct ForceModelData {
    std::vector<Vec3> ext;
    std::vector<int> pairD, pairP;
    std::vector<double> k;
};
Platform& shim_cuda_platform() {
    static Platform* platform = NULL;
    if (!platform) {
        platform = new ShimCudaPlatform(TGNH_FORCE_I64_SOA);          // OpenMM's int64 fixed-point force buffer
        Platform::registerPlatform(platform);
        registerDrudeTGNHCudaKernelFactories();
        platform->registerKernelFactory(IntegrateDrudeTGNHStepKernel::Name(), new B200DrudeTGNHKernelFactory());
    }
    return *platform;
}
void bind_shim(py::module_& m) {
    py::module_ s = m.def_submodule("shim", "stand-in for the OpenMM classes a DrudeTGNHIntegrator script touches (no OpenMM in this environment)");
    // Forces are owned by the System they are added to (OpenMM's rule): Python never deletes them
    py::class_<Force, std::unique_ptr<Force, py::nodelete>>(s, "Force");
    py::class_<DrudeForce, Force, std::unique_ptr<DrudeForce, py::nodelete>>(s, "DrudeForce")
        .def(py::init<>())
        .def("getNumParticles", &DrudeForce::getNumParticles)
        .def("addParticle", &DrudeForce::addParticle, py::arg("particle"), py::arg("particle1"), py::arg("particle2") = -1, py::arg("particle3") = -1,
             py::arg("particle4") = -1, py::arg("charge") = -1.0, py::arg("polarizability") = 1.0, py::arg("aniso12") = 1.0, py::arg("aniso34") = 1.0);
    py::class_<ShimBondForce, Force, std::unique_ptr<ShimBondForce, py::nodelete>>(s, "BondForce", "only the bond list: what System molecules (the integrator's residues) are made of")
        .def(py::init<>())
        .def("addBond", &ShimBondForce::addBond, py::arg("particle1"), py::arg("particle2"));
    py::class_<CMMotionRemover, Force, std::unique_ptr<CMMotionRemover, py::nodelete>>(s, "CMMotionRemover").def(py::init<>());
    py::class_<System>(s, "System")
        .def(py::init<>())
        .def("addParticle", &System::addParticle, py::arg("mass"))
        .def("getNumParticles", &System::getNumParticles)
        .def("getParticleMass", &System::getParticleMass, py::arg("index"))
        .def("addConstraint", &System::addConstraint, py::arg("particle1"), py::arg("particle2"), py::arg("distance"))
        .def("getNumConstraints", &System::getNumConstraints)
        .def("addForce", [](System& sys, Force* f) { return sys.addForce(f); }, py::arg("force"))
        .def("getNumForces", &System::getNumForces);
    py::class_<Platform, std::unique_ptr<Platform, py::nodelete>>(s, "Platform")
        .def("getName", &Platform::getName)
        .def_static("getPlatformByName", [](const std::string& name) -> Platform& {
            if (name == "CUDA") return shim_cuda_platform();
            return Platform::getPlatformByName(name);
        }, py::return_value_policy::reference, py::arg("name"));
    py::class_<State>(s, "State")
        .def("getTime", &State::getTime)
        .def("getPositions", [](const State& st) { return from_vec3(st.getPositions()); })
        .def("getVelocities", [](const State& st) { return from_vec3(st.getVelocities()); })
        .def("getForces", [](const State& st) { return from_vec3(st.getForces()); })
        .def("getKineticEnergy", &State::getKineticEnergy);
    py::class_<Context>(s, "Context")
        .def(py::init([](System& system, DrudeTGNHIntegrator& integrator, Platform& platform, const std::map<std::string, std::string>& properties) {
                 return new Context(system, integrator, platform, properties);
             }),
             py::arg("system"), py::arg("integrator"), py::arg("platform"), py::arg("properties") = std::map<std::string, std::string>(),
             py::keep_alive<1, 2>(), py::keep_alive<1, 3>())
        .def("setPositions", [](Context& c, const Arr& a) { c.setPositions(to_vec3(a, c.getSystem().getNumParticles())); }, py::arg("positions"))
        .def("setVelocities", [](Context& c, const Arr& a) { c.setVelocities(to_vec3(a, c.getSystem().getNumParticles())); }, py::arg("velocities"))
        .def("getState", [](Context& c, bool getPositions, bool getVelocities, bool getForces, bool getEnergy) {
                 return c.getState((getPositions ? State::Positions : 0) | (getVelocities ? State::Velocities : 0) | (getForces ? State::Forces : 0) |
                                   (getEnergy ? State::Energy : 0));
             },
             py::arg("getPositions") = false, py::arg("getVelocities") = false, py::arg("getForces") = false, py::arg("getEnergy") = false)
        .def("setForceModel", [](Context& c, const Arr& external, const std::vector<int>& pairDrude, const std::vector<int>& pairParent, const std::vector<double>& k) {
                 std::shared_ptr<ForceModelData> d(new ForceModelData());
                 d->ext = to_vec3(external, c.getSystem().getNumParticles());
                 d->pairD = pairDrude; d->pairP = pairParent; d->k = k;
                 if (pairDrude.size() != pairParent.size() || k.size() != pairDrude.size()) throw OpenMMException("setForceModel: one spring constant per Drude pair");
                 ShimCudaPlatform::installForceModel(c.getImpl(), [d](const std::vector<Vec3>& pos, std::vector<Vec3>& f) {
                     for (size_t i = 0; i < f.size(); i++) f[i] = d->ext[i];
                     for (size_t i = 0; i < d->pairD.size(); i++)
                         for (int x = 0; x < 3; x++) {
                             const double fc = -d->k[i] * (pos[d->pairD[i]][x] - pos[d->pairP[i]][x]);
                             f[d->pairD[i]][x] += fc;
                             f[d->pairP[i]][x] -= fc;
                         }
                 });
                 c.getImpl().calcForcesAndEnergy(true, false);        // the step assumes valid forces on entry (DrudeTGNHIntegrator::stateChanged)
             },
             py::arg("externalForces"), py::arg("pairDrude") = std::vector<int>(), py::arg("pairParent") = std::vector<int>(), py::arg("springConstants") = std::vector<double>(),
             "shim only: the forces the Context evaluates (constant external forces + harmonic Drude springs), computed on the host from the downloaded positions");
}
}  // namespace
#endif

PYBIND11_MODULE(drudetgnhplugin, m) {
    m.doc() = "DrudeTGNHIntegrator (temperature-grouped dual Nose-Hoover thermostat for Drude-polarizable MD), B200 build";
    py::register_exception<OpenMMException>(m, "OpenMMException");
    py::class_<DrudeTGNHIntegrator>(m, "DrudeTGNHIntegrator")
        .def(py::init([](double temperature, double couplingTime, double drudeTemperature, double drudeCouplingTime, double stepSize,
                         int drudeStepsPerRealStep, int numNHChains, int useDrudeNHChains, int useCOMTempGroup) {
                 return new DrudeTGNHIntegrator(temperature, couplingTime, drudeTemperature, drudeCouplingTime, stepSize, drudeStepsPerRealStep,
                                                numNHChains, useDrudeNHChains != 0, useCOMTempGroup != 0);
             }),
             py::arg("temperature"), py::arg("couplingTime"), py::arg("drudeTemperature"), py::arg("drudeCouplingTime"), py::arg("stepSize"),
             py::arg("drudeStepsPerRealStep") = 20, py::arg("numNHChains") = 1, py::arg("useDrudeNHChains") = 1, py::arg("useCOMTempGroup") = 1)
        .def("getTemperature", [](const DrudeTGNHIntegrator& i) { return with_unit(i.getTemperature(), "kelvin"); })
        .def("setTemperature", &DrudeTGNHIntegrator::setTemperature, py::arg("temp"))
        .def("getCouplingTime", [](const DrudeTGNHIntegrator& i) { return with_unit(i.getCouplingTime(), "picosecond"); })
        .def("setCouplingTime", &DrudeTGNHIntegrator::setCouplingTime, py::arg("tau"))
        .def("getDrudeTemperature", [](const DrudeTGNHIntegrator& i) { return with_unit(i.getDrudeTemperature(), "kelvin"); })
        .def("setDrudeTemperature", &DrudeTGNHIntegrator::setDrudeTemperature, py::arg("temp"))
        .def("getDrudeCouplingTime", [](const DrudeTGNHIntegrator& i) { return with_unit(i.getDrudeCouplingTime(), "picosecond"); })
        .def("setDrudeCouplingTime", &DrudeTGNHIntegrator::setDrudeCouplingTime, py::arg("tau"))
        .def("getMaxDrudeDistance", [](const DrudeTGNHIntegrator& i) { return with_unit(i.getMaxDrudeDistance(), "nanometer"); })
        .def("setMaxDrudeDistance", &DrudeTGNHIntegrator::setMaxDrudeDistance, py::arg("distance"))
        .def("getStepSize", &DrudeTGNHIntegrator::getStepSize)
        .def("setStepSize", &DrudeTGNHIntegrator::setStepSize, py::arg("size"))
        .def("getConstraintTolerance", &DrudeTGNHIntegrator::getConstraintTolerance)
        .def("setConstraintTolerance", &DrudeTGNHIntegrator::setConstraintTolerance, py::arg("tol"))
        .def("step", &DrudeTGNHIntegrator::step, py::arg("steps"))
        .def("getDrudeStepsPerRealStep", &DrudeTGNHIntegrator::getDrudeStepsPerRealStep)
        .def("setDrudeStepsPerRealStep", &DrudeTGNHIntegrator::setDrudeStepsPerRealStep, py::arg("drudeSteps"))
        .def("getNumNHChains", &DrudeTGNHIntegrator::getNumNHChains)
        .def("setNumNHChains", &DrudeTGNHIntegrator::setNumNHChains, py::arg("numChains"))
        .def("getUseDrudeNHChains", &DrudeTGNHIntegrator::getUseDrudeNHChains)
        .def("setUseDrudeNHChains", &DrudeTGNHIntegrator::setUseDrudeNHChains, py::arg("useDrudeNHChains"))
        .def("getUseCOMTempGroup", [](const DrudeTGNHIntegrator& i) { return (int)i.getUseCOMTempGroup(); })
        .def("setUseCOMTempGroup", &DrudeTGNHIntegrator::setUseCOMTempGroup, py::arg("useCOMTempGroup"))
        .def("getNumTempGroups", &DrudeTGNHIntegrator::getNumTempGroups)
        .def("addTempGroup", &DrudeTGNHIntegrator::addTempGroup)
        .def("addParticleTempGroup", &DrudeTGNHIntegrator::addParticleTempGroup, py::arg("tempGroup"))
        .def("setParticleTempGroup", &DrudeTGNHIntegrator::setParticleTempGroup, py::arg("particle"), py::arg("tempGroup"))
        // SWIG's `int& OUTPUT` typemap: the group comes back as the return value
        .def("getParticleTempGroup", [](const DrudeTGNHIntegrator& i, int particle) { int tg; i.getParticleTempGroup(particle, tg); return tg; }, py::arg("particle"));
    // XmlSerializer.serialize / deserialize for the integrator (what openmm.XmlSerializer does for the SWIG class)
    // version 1 (default) is the upstream plugin's format; version 2 also keeps maxDrudeDistance, useCOMTempGroup and the temperature groups
    m.def("serialize", [](const DrudeTGNHIntegrator& i, int version) {
        const int saved = DrudeTGNHIntegratorProxy::writeVersion;
        DrudeTGNHIntegratorProxy::writeVersion = version;
        std::stringstream s;
        try { XmlSerializer::serialize<DrudeTGNHIntegrator>(&i, "Integrator", s); } catch (...) { DrudeTGNHIntegratorProxy::writeVersion = saved; throw; }
        DrudeTGNHIntegratorProxy::writeVersion = saved;
        return s.str();
    }, py::arg("integrator"), py::arg("version") = 1);
    m.def("deserialize", [](const std::string& xml) { std::stringstream s(xml); return XmlSerializer::deserialize<DrudeTGNHIntegrator>(s); }, py::return_value_policy::take_ownership);
#ifndef TGNH_WITH_OPENMM
    bind_shim(m);
#endif
}
