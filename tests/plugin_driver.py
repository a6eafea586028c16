"""Test helper: ctypes wrapper over plugin/libtgnhplugin_test.so — this repo's plugin stack (DrudeTGNHIntegrator ->
B200IntegrateDrudeTGNHStepKernel -> libtgnh.so) on the shim's CUDA platform; same calling convention as oracle/ref.py."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "plugin", "libtgnhplugin_test.so")
_lib = None


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
        L.plugin_set_next_constraints.argtypes = [C.c_int, ip, ip]
        L.plugin_constraint_calls.argtypes = [C.c_void_p]
        L.plugin_set_next_precision.argtypes = [C.c_char_p]
        L.plugin_kinetic_energy.argtypes = [C.c_void_p]
        L.plugin_kinetic_energy.restype = C.c_double
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


class PluginSim:
    def __init__(self, system, force_model=0, has_cm_motion_remover=False, with_constraints=False, precision="single"):
        s = system
        L = lib()
        L.plugin_set_next_precision(precision.encode())      # the Context's "Precision" platform property
        if with_constraints and len(s.constraints):
            c = np.ascontiguousarray(s.constraints, np.int32)
            a, b = np.ascontiguousarray(c[:, 0]), np.ascontiguousarray(c[:, 1])
            L.plugin_set_next_constraints(len(c), _ip(a), _ip(b))
        self._keep = [np.ascontiguousarray(s.masses, np.float64), np.ascontiguousarray(s.pair_drude, np.int32),
                      np.ascontiguousarray(s.pair_parent, np.int32), np.ascontiguousarray(s.res_id, np.int32),
                      np.ascontiguousarray(s.temp_group, np.int32), np.ascontiguousarray(s.k_spring, np.float64)]
        m, pd, pp, res, tg, k = self._keep
        self.h = L.ref_create(s.num_particles, _dp(m), len(pd), _ip(pd), _ip(pp), _ip(res), _ip(tg), s.num_temp_groups, s.temperature,
                              s.coupling_time, s.drude_temperature, s.drude_coupling_time, s.step_size, s.drude_steps, s.num_nh_chains,
                              int(s.use_drude_nh_chains), int(s.use_com_temp_group), s.max_drude_distance, int(has_cm_motion_remover),
                              force_model, _dp(k))
        if not self.h:
            raise RuntimeError(L.ref_last_error().decode())

    def step(self, pos, vel, force, nsteps=1, ext_force=None):
        if lib().ref_step(self.h, _dp(pos), _dp(vel), _dp(force), nsteps, _dp(ext_force)):
            raise RuntimeError(lib().ref_last_error().decode())

    def constraint_calls(self):
        return lib().plugin_constraint_calls(self.h)

    def kinetic_energy(self):
        return lib().plugin_kinetic_energy(self.h)

    def close(self):
        if getattr(self, "h", None):
            lib().ref_destroy(self.h)
            self.h = None
