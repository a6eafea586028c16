#ifndef SHIM_SIMTK_UTILITIES_H_
#define SHIM_SIMTK_UTILITIES_H_
#include <cmath>
#include "SimTKOpenMMRealType.h"
#include "RealVec.h"
#endif
