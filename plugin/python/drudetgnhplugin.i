/* SWIG interface of the `drudetgnhplugin` Python module for builds against a real OpenMM (needs swig and OpenMM's
 * swig/OpenMMSwigHeaders.i + typemaps.i; neither exists in this repo's container — INTEGRATION.md).  Module name,
 * class, method list, unit-wrapped getters and the Python-side defaults are those of the reference's
 * /root/reference/python/drudetgnhplugin.i:1-94 so that example/nacl_tg.py runs unchanged. */
%module drudetgnhplugin

%import(module="simtk.openmm") "swig/OpenMMSwigHeaders.i"
%include "swig/typemaps.i"
%include "std_vector.i"
namespace std {
  %template(vectord) vector<double>;
  %template(vectori) vector<int>;
};

%{
#include "OpenMM.h"
#include "OpenMMDrude.h"
#include "OpenMMDrudeTGNH.h"
%}

%pythoncode %{
import simtk.openmm as mm
import simtk.unit as unit
%}

/* getters return Quantities */
%define TGNH_UNIT(method, u)
%pythonappend OpenMM::DrudeTGNHIntegrator::method() const %{
   val = unit.Quantity(val, u)
%}
%enddef
TGNH_UNIT(getTemperature, unit.kelvin)
TGNH_UNIT(getCouplingTime, unit.picosecond)
TGNH_UNIT(getDrudeTemperature, unit.kelvin)
TGNH_UNIT(getDrudeCouplingTime, unit.picosecond)
TGNH_UNIT(getMaxDrudeDistance, unit.nanometer)

namespace OpenMM {

class DrudeTGNHIntegrator : public Integrator {
public:
   /* NB: Python-side defaults differ from C++ for useDrudeNHChains (True), as in the reference's interface file */
   DrudeTGNHIntegrator(double temperature, double couplingTime, double drudeTemperature, double drudeCouplingTime, double stepSize,
                       int drudeStepsPerRealStep=20, int numNHChains=1, int useDrudeNHChains=True, int useCOMTempGroup=True);
   double getTemperature() const;
   void setTemperature(double temp);
   double getCouplingTime() const;
   void setCouplingTime(double tau);
   double getDrudeTemperature() const;
   void setDrudeTemperature(double temp);
   double getDrudeCouplingTime() const;
   void setDrudeCouplingTime(double tau);
   double getMaxDrudeDistance() const;
   void setMaxDrudeDistance(double distance);
   virtual void step(int steps);
   int getDrudeStepsPerRealStep() const;
   void setDrudeStepsPerRealStep(int drudeSteps);
   int getNumNHChains() const;
   void setNumNHChains(int numChains);
   int getUseDrudeNHChains() const;
   void setUseDrudeNHChains(int useDrudeNHChains);
   int getUseCOMTempGroup() const;
   void setUseCOMTempGroup(int useCOMTempGroup);
   int getNumTempGroups() const;
   int addTempGroup();
   int addParticleTempGroup(int tempGroup);
   void setParticleTempGroup(int particle, int tempGroup);
   %apply int& OUTPUT {int& tempGroup};
   void getParticleTempGroup(int particle, int& tempGroup) const;
   %clear int& tempGroup;
};

}
