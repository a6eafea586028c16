"""TEST INFRASTRUCTURE — ctypes wrapper over oracle/_refcuda/librefcuda.so and libb200cuda.so (oracle/cuda_driver.cpp).

    flavour "reference": the reference's OWN CUDA platform (openmmapi + platforms/cuda sources and kernel strings, unmodified
                         from /root/reference) on the GPU behind the CUDA-platform stand-in of shim/cuda
    flavour "b200":      the same unmodified reference DrudeTGNHIntegrator and stand-in platform, "IntegrateDrudeTGNHStep"
                         served by this repo's plugin (the -DTGNH_WITH_OPENMM, CudaContext-facing build) over libtgnh.so

Both exist only where /root/reference was mounted at build time (this container); the libraries then travel with the repo.
Needs a GPU (libcuda) to load.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATHS = {"reference": os.path.join(_HERE, "_refcuda", "librefcuda.so"), "b200": os.path.join(_HERE, "_refcuda", "libb200cuda.so")}
_libs = {}
PRECISION = {"single": 0, "mixed": 1, "double": 2}


def available(flavour="reference"):
    return os.path.exists(PATHS[flavour])


def lib(flavour):
    if flavour not in _libs:
        L = C.CDLL(PATHS[flavour])
        dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
        L.cudadrv_last_error.restype = C.c_char_p
        L.cudadrv_flavour.restype = C.c_char_p
        L.cudadrv_create.restype = vp
        L.cudadrv_create.argtypes = [C.c_int, dp, C.c_int, ip, ip, ip, ip, C.c_int] + [C.c_double] * 5 + [C.c_int] * 4 + [C.c_double] + [C.c_int] * 3 + [dp, C.c_int, C.c_int, ip, ip]
        L.cudadrv_destroy.argtypes = [vp]
        L.cudadrv_num_residues.argtypes = [vp]
        L.cudadrv_set_state.argtypes = [vp, dp, dp, dp]
        L.cudadrv_set_velocities.argtypes = [vp, dp]
        L.cudadrv_step.argtypes = [vp, C.c_int]
        L.cudadrv_time_steps.argtypes = [vp, C.c_int]
        L.cudadrv_time_steps.restype = C.c_double
        L.cudadrv_get_state.argtypes = [vp, dp, dp, dp, dp]
        L.cudadrv_get_thermostat.argtypes = [vp, dp, dp, dp, dp]
        L.cudadrv_get_thermostat_params.argtypes = [vp, dp, dp]
        L.cudadrv_counters.argtypes = [vp, C.POINTER(C.c_longlong)]
        L.cudadrv_last_kernel_source.argtypes = [vp, C.c_char_p, C.c_int]
        L.cudadrv_kesum.argtypes = [vp]
        L.cudadrv_kesum.restype = C.c_double
        L.cudadrv_kernel_generation.argtypes = [vp]
        _libs[flavour] = L
    return _libs[flavour]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


class CudaDriverError(RuntimeError):
    pass


class CudaSim:
    """DrudeTGNHIntegrator (the reference's) + a CUDA-platform kernel on a synth.DrudeSystem, driven like a user script."""

    COUNTERS = ("launches", "initialize_contexts", "force_evaluations", "reorders", "apply_constraints", "apply_velocity_constraints",
                "compute_virtual_sites", "step_count")

    def __init__(self, system, flavour="reference", precision="double", force_model=0, has_cm_motion_remover=False, reorder_interval=0,
                 assign_groups=True, with_constraints=True):
        s = system
        self.L = L = lib(flavour)
        self.flavour, self.n = flavour, s.num_particles
        self.T, self.M = s.num_temp_groups + 2, s.num_nh_chains
        self._keep = [np.ascontiguousarray(s.masses, np.float64), np.ascontiguousarray(s.pair_drude, np.int32),
                      np.ascontiguousarray(s.pair_parent, np.int32), np.ascontiguousarray(s.res_id, np.int32),
                      np.ascontiguousarray(s.temp_group, np.int32), np.ascontiguousarray(s.k_spring, np.float64)]
        m, pd, pp, res, tg, k = self._keep
        cons = np.ascontiguousarray(s.constraints, np.int32).reshape(-1, 2) if with_constraints else np.zeros((0, 2), np.int32)
        ca, cb = np.ascontiguousarray(cons[:, 0]), np.ascontiguousarray(cons[:, 1])
        self._keep += [ca, cb]
        self.h = L.cudadrv_create(s.num_particles, _dp(m), len(pd), _ip(pd), _ip(pp), _ip(res), _ip(tg) if assign_groups else None,
                                  s.num_temp_groups, s.temperature, s.coupling_time, s.drude_temperature, s.drude_coupling_time, s.step_size,
                                  s.drude_steps, s.num_nh_chains, int(s.use_drude_nh_chains), int(s.use_com_temp_group), s.max_drude_distance,
                                  int(has_cm_motion_remover), PRECISION[precision], force_model, _dp(k), reorder_interval,
                                  len(cons), _ip(ca) if len(cons) else None, _ip(cb) if len(cons) else None)
        if not self.h:
            raise CudaDriverError(L.cudadrv_last_error().decode())

    def _chk(self, rc):
        if rc:
            raise CudaDriverError(self.L.cudadrv_last_error().decode())

    def set_state(self, pos, vel, force):
        self._chk(self.L.cudadrv_set_state(self.h, _dp(np.ascontiguousarray(pos)), _dp(np.ascontiguousarray(vel)), _dp(np.ascontiguousarray(force))))

    def set_velocities(self, vel):
        self._chk(self.L.cudadrv_set_velocities(self.h, _dp(np.ascontiguousarray(vel))))

    def step(self, nsteps=1):
        self._chk(self.L.cudadrv_step(self.h, nsteps))

    def time_steps(self, nsteps):
        ms = self.L.cudadrv_time_steps(self.h, nsteps)
        if ms < 0:
            raise CudaDriverError(self.L.cudadrv_last_error().decode())
        return ms

    def get_state(self, energy=False):
        pos = np.zeros((self.n, 3)); vel = np.zeros((self.n, 3)); force = np.zeros((self.n, 3))
        ke = C.c_double()
        self._chk(self.L.cudadrv_get_state(self.h, _dp(pos), _dp(vel), _dp(force), C.byref(ke) if energy else None))
        return (pos, vel, force, ke.value) if energy else (pos, vel, force)

    def thermostat(self):
        eta = np.zeros((self.T, self.M)); ed = np.zeros((self.T, self.M + 1)); edd = np.zeros((self.T, self.M)); vs = np.zeros(self.T)
        self._chk(self.L.cudadrv_get_thermostat(self.h, _dp(eta), _dp(ed), _dp(edd), _dp(vs)))
        return eta, ed, edd, vs

    def thermostat_params(self):
        nkbt = np.zeros(self.T); q = np.zeros((self.T, self.M))
        self._chk(self.L.cudadrv_get_thermostat_params(self.h, _dp(nkbt), _dp(q)))
        return nkbt, q

    @property
    def ke_sum(self):
        return self.L.cudadrv_kesum(self.h)

    @property
    def kernel_generation(self):
        return self.L.cudadrv_kernel_generation(self.h)

    def counters(self):
        out = (C.c_longlong * 8)()
        self._chk(self.L.cudadrv_counters(self.h, out))
        return dict(zip(self.COUNTERS, [int(x) for x in out]))

    def kernel_source(self):
        n = self.L.cudadrv_last_kernel_source(self.h, None, 0)
        buf = C.create_string_buffer(n + 1)
        self.L.cudadrv_last_kernel_source(self.h, buf, n + 1)
        return buf.value.decode()

    def close(self):
        if getattr(self, "h", None):
            self.L.cudadrv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
