#ifndef SHIM_SIMTK_REALTYPE_H_
#define SHIM_SIMTK_REALTYPE_H_
#include <cmath>   // as the original does (M_PI); the reference's sources rely on it for sqrt / exp / pow
// OpenMM's physical constants (SimTKOpenMMRealType.h, 2019 SI values); BOLTZ is the one the integrator uses
#define ANGSTROM     (1e-10)
#define KILO         (1e3)
#define NANO         (1e-9)
#define PICO         (1e-12)
#define A2NM         (ANGSTROM/NANO)
#define NM2A         (NANO/ANGSTROM)
#define RAD2DEG      (180.0/M_PI)
#define CAL2JOULE    (4.184)
#define E_CHARGE     (1.602176634e-19)
#define AMU          (1.66053906660e-27)
#define BOLTZMANN    (1.380649e-23)
#define AVOGADRO     (6.02214076e23)
#define RGAS         (BOLTZMANN*AVOGADRO)
#define BOLTZ        (RGAS/KILO)
#define FARADAY      (E_CHARGE*AVOGADRO)
#define ELECTRONVOLT (E_CHARGE*AVOGADRO/KILO)
#define EPSILON0     (1e-6*8.8541878128e-12/(E_CHARGE*E_CHARGE*AVOGADRO))
#define ONE_4PI_EPS0 (1/(4*M_PI*EPSILON0))
#endif
