#ifndef SHIM_REALVEC_H_
#define SHIM_REALVEC_H_
#include "openmm/Vec3.h"
namespace OpenMM {
typedef double RealOpenMM;     // OpenMM <= 7.3 reference platform vocabulary
typedef Vec3 RealVec;
}
#endif
