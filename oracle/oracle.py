"""TEST INFRASTRUCTURE — ctypes wrapper over oracle/liboracle.so (the CPU restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

TG, REF = 0, 1
FORCE_FIXED, FORCE_HARMONIC = 0, 1
BOLTZ = 1.380649e-23 * 6.02214076e23 / 1000.0


class _Params(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "num_particles", "num_pairs", "num_residues", "num_temp_groups", "num_nh_chains",
        "drude_steps_per_real_step", "use_drude_nh_chains", "use_com_temp_group",
        "has_cm_motion_remover", "num_constraints")] + [(n, C.c_double) for n in (
        "temperature", "coupling_time", "drude_temperature", "drude_coupling_time", "step_size",
        "max_drude_distance")] + [
        ("masses", C.POINTER(C.c_double)), ("pair_drude", C.POINTER(C.c_int)),
        ("pair_parent", C.POINTER(C.c_int)), ("particle_temp_group", C.POINTER(C.c_int)),
        ("particle_res_id", C.POINTER(C.c_int)), ("constraint_p", C.POINTER(C.c_int)),
        ("constraint_p1", C.POINTER(C.c_int))]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "tgnh_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.tgnh_oracle_create.argtypes = [C.POINTER(_Params), C.c_int, C.POINTER(C.c_void_p)]
        L.tgnh_oracle_last_error.restype = C.c_char_p
        L.tgnh_oracle_destroy.argtypes = [C.c_void_p]
        dp = C.POINTER(C.c_double)
        for name, args in {
            "tgnh_oracle_propagate_nh_chain": [C.c_void_p, dp],
            "tgnh_oracle_half_kick": [C.c_void_p, dp, dp],
            "tgnh_oracle_drift": [C.c_void_p, dp, dp],
            "tgnh_oracle_hard_wall": [C.c_void_p, dp, dp],
            "tgnh_oracle_step": [C.c_void_p, dp, dp, dp, C.c_int, C.c_int, dp, dp],
            "tgnh_oracle_compute_ke2": [C.c_void_p, dp, dp],
        }.items():
            getattr(L, name).argtypes = args
            getattr(L, name).restype = C.c_int
        L.tgnh_oracle_num_thermostats.argtypes = [C.c_void_p]
        L.tgnh_oracle_get_ke2.argtypes = [C.c_void_p, dp]
        L.tgnh_oracle_get_vscale.argtypes = [C.c_void_p, dp]
        L.tgnh_oracle_get_chain_state.argtypes = [C.c_void_p, dp, dp, dp]
        L.tgnh_oracle_set_chain_state.argtypes = [C.c_void_p, dp, dp, dp]
        L.tgnh_oracle_get_thermostat_params.argtypes = [C.c_void_p, dp, dp, dp]
        L.tgnh_oracle_get_ke_sum.argtypes = [C.c_void_p]
        L.tgnh_oracle_get_ke_sum.restype = C.c_double
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


class OracleError(RuntimeError):
    pass


class Oracle:
    """fp64 CPU oracle for one Drude system (see openmm_drudenose_b200.synth.DrudeSystem).

    pos / vel / force are float64 arrays of shape [N, 3] updated in place.
    """

    def __init__(self, system, which=TG, constraints=None, has_cm_motion_remover=False):
        self.L = lib()
        s = system
        self.N, self.M = s.num_particles, s.num_nh_chains
        self._keep = dict(
            masses=np.ascontiguousarray(s.masses, np.float64),
            pd=np.ascontiguousarray(s.pair_drude, np.int32), pp=np.ascontiguousarray(s.pair_parent, np.int32),
            tg=np.ascontiguousarray(s.temp_group, np.int32), res=np.ascontiguousarray(s.res_id, np.int32))
        cons = np.zeros((0, 2), np.int32) if constraints is None else np.ascontiguousarray(constraints, np.int32)
        self._keep["c0"] = np.ascontiguousarray(cons[:, 0]); self._keep["c1"] = np.ascontiguousarray(cons[:, 1])
        k = self._keep
        p = _Params(
            num_particles=s.num_particles, num_pairs=len(k["pd"]), num_residues=s.num_residues,
            num_temp_groups=s.num_temp_groups, num_nh_chains=s.num_nh_chains,
            drude_steps_per_real_step=s.drude_steps, use_drude_nh_chains=int(s.use_drude_nh_chains),
            use_com_temp_group=int(s.use_com_temp_group), has_cm_motion_remover=int(has_cm_motion_remover),
            num_constraints=len(cons), temperature=s.temperature, coupling_time=s.coupling_time,
            drude_temperature=s.drude_temperature, drude_coupling_time=s.drude_coupling_time,
            step_size=s.step_size, max_drude_distance=s.max_drude_distance,
            masses=_dp(k["masses"]), pair_drude=_ip(k["pd"]), pair_parent=_ip(k["pp"]),
            particle_temp_group=_ip(k["tg"]), particle_res_id=_ip(k["res"]),
            constraint_p=_ip(k["c0"]), constraint_p1=_ip(k["c1"]))
        h = C.c_void_p()
        if self.L.tgnh_oracle_create(C.byref(p), which, C.byref(h)):
            raise OracleError(self.L.tgnh_oracle_last_error().decode())
        self.h = h
        self.T = self.L.tgnh_oracle_num_thermostats(h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.tgnh_oracle_destroy(self.h)
            self.h = None

    def _chk(self, rc):
        if rc:
            raise OracleError(self.L.tgnh_oracle_last_error().decode())

    @staticmethod
    def _a(x):
        assert x.dtype == np.float64 and x.flags.c_contiguous
        return _dp(x)

    def propagate_nh_chain(self, vel):
        self._chk(self.L.tgnh_oracle_propagate_nh_chain(self.h, self._a(vel)))

    def half_kick(self, vel, force):
        self._chk(self.L.tgnh_oracle_half_kick(self.h, self._a(vel), self._a(force)))

    def drift(self, pos, vel):
        self._chk(self.L.tgnh_oracle_drift(self.h, self._a(pos), self._a(vel)))

    def hard_wall(self, pos, vel):
        self._chk(self.L.tgnh_oracle_hard_wall(self.h, self._a(pos), self._a(vel)))

    def step(self, pos, vel, force, nsteps=1, force_model=FORCE_FIXED, ext_force=None, k_spring=None):
        self._chk(self.L.tgnh_oracle_step(self.h, self._a(pos), self._a(vel), self._a(force), nsteps, force_model,
                                          None if ext_force is None else self._a(ext_force),
                                          None if k_spring is None else self._a(k_spring)))

    def compute_ke2(self, vel):
        out = np.zeros(self.T)
        self._chk(self.L.tgnh_oracle_compute_ke2(self.h, self._a(vel), _dp(out)))
        return out

    @property
    def ke2(self):
        out = np.zeros(self.T); self.L.tgnh_oracle_get_ke2(self.h, _dp(out)); return out

    @property
    def vscale(self):
        out = np.zeros(self.T); self.L.tgnh_oracle_get_vscale(self.h, _dp(out)); return out

    @property
    def ke_sum(self):
        return self.L.tgnh_oracle_get_ke_sum(self.h)

    def chain_state(self):
        eta = np.zeros((self.T, self.M)); ed = np.zeros((self.T, self.M + 1)); edd = np.zeros((self.T, self.M))
        self.L.tgnh_oracle_get_chain_state(self.h, _dp(eta), _dp(ed), _dp(edd))
        return eta, ed, edd

    def set_chain_state(self, eta, ed, edd):
        eta = np.ascontiguousarray(eta, np.float64); ed = np.ascontiguousarray(ed, np.float64)
        edd = np.ascontiguousarray(edd, np.float64)
        self.L.tgnh_oracle_set_chain_state(self.h, _dp(eta), _dp(ed), _dp(edd))

    def thermostat_params(self):
        dof = np.zeros(self.T); nkbt = np.zeros(self.T); q = np.zeros((self.T, self.M))
        self.L.tgnh_oracle_get_thermostat_params(self.h, _dp(dof), _dp(nkbt), _dp(q))
        return dof, nkbt, q


def harmonic_forces(system, pos, ext_force=None):
    """Drude springs -k (x_d - x_p) + optional external force; numpy fp64 (same formula as the C oracle)."""
    f = np.zeros_like(pos) if ext_force is None else ext_force.copy()
    d = pos[system.pair_drude] - pos[system.pair_parent]
    fd = -system.k_spring[:, None] * d
    np.add.at(f, system.pair_drude, fd)
    np.add.at(f, system.pair_parent, -fd)
    return f
