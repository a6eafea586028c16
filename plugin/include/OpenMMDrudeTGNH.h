#ifndef TGNH_B200_OPENMM_DRUDETGNH_H_
#define TGNH_B200_OPENMM_DRUDETGNH_H_
// umbrella header, as /root/reference/openmmapi/include/OpenMMDrudeTGNH.h
#include "openmm/DrudeForce.h"
#include "openmm/DrudeTGNHIntegrator.h"
#endif
