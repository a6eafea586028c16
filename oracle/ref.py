"""TEST INFRASTRUCTURE — ctypes wrapper over oracle/_ref/libref.so: the reference's OWN sources
(openmmapi/src/DrudeTGNHIntegrator.cpp, platforms/reference/src/*.cpp, serialization/src/*.cpp, compiled unmodified
from /root/reference against the OpenMM API shim; see oracle/Makefile and oracle/ref_driver.cpp).
Exists only where /root/reference was mounted at build time (this container); the .so then travels with the repo.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.ref_last_error.restype = C.c_char_p
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int, dp, C.c_int, ip, ip, ip, ip, C.c_int] + [C.c_double] * 5 + [C.c_int] * 4 + [C.c_double, C.c_int, C.c_int, dp]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_num_residues.argtypes = [C.c_void_p]
        L.ref_step.argtypes = [C.c_void_p, dp, dp, dp, C.c_int, dp]
        L.ref_serialization_roundtrip.argtypes = [C.c_double] * 5 + [C.c_int] * 3 + [C.c_double, dp, C.c_char_p, C.c_int]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


class RefError(RuntimeError):
    pass


class ReferenceSim:
    """The real DrudeTGNHIntegrator + ReferenceIntegrateDrudeTGNHStepKernel on a synth.DrudeSystem."""

    def __init__(self, system, force_model=0, has_cm_motion_remover=False, assign_groups=True):
        s = system
        L = lib()
        self._keep = [np.ascontiguousarray(s.masses, np.float64), np.ascontiguousarray(s.pair_drude, np.int32),
                      np.ascontiguousarray(s.pair_parent, np.int32), np.ascontiguousarray(s.res_id, np.int32),
                      np.ascontiguousarray(s.temp_group, np.int32), np.ascontiguousarray(s.k_spring, np.float64)]
        m, pd, pp, res, tg, k = self._keep
        self.h = L.ref_create(s.num_particles, _dp(m), len(pd), _ip(pd), _ip(pp), _ip(res), _ip(tg) if assign_groups else None,
                              s.num_temp_groups, s.temperature, s.coupling_time, s.drude_temperature, s.drude_coupling_time,
                              s.step_size, s.drude_steps, s.num_nh_chains, int(s.use_drude_nh_chains), int(s.use_com_temp_group),
                              s.max_drude_distance, int(has_cm_motion_remover), force_model, _dp(k))
        if not self.h:
            raise RefError(L.ref_last_error().decode())
        self.num_residues = L.ref_num_residues(self.h)

    def step(self, pos, vel, force, nsteps=1, ext_force=None):
        if lib().ref_step(self.h, _dp(pos), _dp(vel), _dp(force), nsteps, _dp(ext_force)):
            raise RefError(lib().ref_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_destroy(self.h)
            self.h = None


def serialization_roundtrip(temperature, coupling, drude_temperature, drude_coupling, step_size, drude_steps=20, chains=1,
                            use_drude_chains=False, constraint_tol=1e-5):
    out = np.zeros(9)
    xml = C.create_string_buffer(4096)
    if lib().ref_serialization_roundtrip(temperature, coupling, drude_temperature, drude_coupling, step_size, drude_steps, chains,
                                         int(use_drude_chains), constraint_tol, _dp(out), xml, 4096):
        raise RefError(lib().ref_last_error().decode())
    return out, xml.value.decode()
