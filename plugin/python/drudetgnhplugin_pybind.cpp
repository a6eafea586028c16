// `drudetgnhplugin` for environments without SWIG / OpenMM's Python layer: the same class, method names, argument
// order and Python-side defaults as the SWIG module (drudetgnhplugin.i), bound with pybind11 against whichever OpenMM
// API the C++ side was compiled with (here: the shim).  Getters return plain floats unless a `unit` module
// (openmm.unit / simtk.unit) is importable, in which case they return Quantities like the reference's module.
#include <pybind11/pybind11.h>

#include <sstream>

#include "OpenMMDrudeTGNH.h"
#include "openmm/serialization/XmlSerializer.h"
#include "openmm/serialization/DrudeTGNHIntegratorProxy.h"

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
}
