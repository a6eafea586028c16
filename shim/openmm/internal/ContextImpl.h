#ifndef SHIM_OPENMM_CONTEXTIMPL_H_
#define SHIM_OPENMM_CONTEXTIMPL_H_
#include "openmm/shim_core.h"
#endif
