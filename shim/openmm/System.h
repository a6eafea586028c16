#ifndef SHIM_OPENMM_FWD_System_H_
#define SHIM_OPENMM_FWD_System_H_
#include "openmm/shim_core.h"
#endif
