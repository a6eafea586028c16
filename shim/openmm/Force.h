#ifndef SHIM_OPENMM_FWD_Force_H_
#define SHIM_OPENMM_FWD_Force_H_
#include "openmm/shim_core.h"
#endif
