"""ctypes binding of libtgnh.so (include/tgnh.h) — the only way Python reaches the CUDA path.

There is no fallback: if the library is missing or no B200 is visible, calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TGNH_LIB") or os.path.join(_HERE, "libtgnh.so")      # TGNH_LIB: experimental builds of the same library

OK, ERR_INVALID_ARGUMENT, ERR_TEMP_GROUP, ERR_UNSUPPORTED, ERR_CUDA, ERR_NCCL, ERR_NO_DEVICE = range(7)
FORCE_F32_SOA, FORCE_I64_SOA = 0, 1
PRECISION_SINGLE, PRECISION_MIXED, PRECISION_DOUBLE = 0, 1, 2
HALF2_DEFAULT, HALF2_DEFER_SCALE, HALF2_KICK_ONLY = 0, 1, 2
HOST_FORCES_UNCHANGED = 1
UNIQUE_ID_BYTES = 128
BOLTZ = 1.380649e-23 * 6.02214076e23 / 1000.0

# every symbol include/tgnh.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "tgnh_create", "tgnh_destroy", "tgnh_last_error", "tgnh_build_info", "tgnh_half1", "tgnh_half1_kick", "tgnh_half1_drift",
    "tgnh_thermostat", "tgnh_half2", "tgnh_flush",
    "tgnh_step", "tgnh_step_host", "tgnh_step_host2", "tgnh_set_posq_correction", "tgnh_invalidate", "tgnh_num_thermostats", "tgnh_num_nh_chains", "tgnh_get_kinetic_energies",
    "tgnh_kinetic_energy", "tgnh_compute_kinetic_energies", "tgnh_get_chain_state", "tgnh_set_chain_state",
    "tgnh_get_vscale", "tgnh_get_thermostat_params", "tgnh_launch_count", "tgnh_exchange_kind", "tgnh_get_exchange_timing", "tgnh_plan_tiles", "tgnh_plan_descriptors", "tgnh_plan_chunks", "tgnh_kernel_generation", "tgnh_residue_per_lane", "tgnh_chunks_per_tile", "tgnh_lazy_second_kick", "tgnh_set_profiling", "tgnh_get_profile", "tgnh_comm_get_unique_id",
    "tgnh_comm_create", "tgnh_comm_destroy",
]


class TgnhError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libtgnh error {code}: {message}")
        self.code = code


class Params(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "num_particles", "padded_num_particles", "num_pairs", "num_residues", "num_temp_groups", "num_constraints",
        "num_nh_chains", "drude_steps_per_real_step", "use_drude_nh_chains", "use_com_temp_group",
        "has_cm_motion_remover", "force_format", "device", "precision")] + [(n, C.c_double) for n in (
        "temperature", "coupling_time", "drude_temperature", "drude_coupling_time", "step_size",
        "max_drude_distance")] + [
        ("masses", C.POINTER(C.c_double)), ("pair_drude", C.POINTER(C.c_int32)), ("pair_parent", C.POINTER(C.c_int32)),
        ("particle_temp_group", C.POINTER(C.c_int32)), ("particle_res_id", C.POINTER(C.c_int32)),
        ("constraint_p", C.POINTER(C.c_int32)), ("constraint_p1", C.POINTER(C.c_int32)), ("comm", C.c_void_p)]


_lib = None


def lib():
    """Load libtgnh.so; raises if it has not been built (no CPU path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        dp, vp = C.POINTER(C.c_double), C.c_void_p
        L.tgnh_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
        L.tgnh_destroy.argtypes = [vp]
        L.tgnh_destroy.restype = None
        L.tgnh_last_error.restype = C.c_char_p
        L.tgnh_build_info.restype = C.c_char_p
        L.tgnh_half1.argtypes = [vp, vp, vp, vp, vp]
        L.tgnh_half2.argtypes = [vp, vp, vp, vp, C.c_int]
        L.tgnh_half1_kick.argtypes = [vp, vp, vp, vp, vp]
        L.tgnh_half1_drift.argtypes = [vp, vp, vp, vp, vp]
        L.tgnh_thermostat.argtypes = [vp, vp, vp, C.c_int]
        L.tgnh_flush.argtypes = [vp, vp, vp]
        L.tgnh_step.argtypes = [vp, vp, vp, vp, vp, C.c_int]
        L.tgnh_step_host.argtypes = [vp, vp, vp, vp, C.c_int, dp]
        L.tgnh_step_host2.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, dp]
        L.tgnh_set_posq_correction.argtypes = [vp, vp]
        L.tgnh_invalidate.argtypes = [vp]
        L.tgnh_num_thermostats.argtypes = [vp]
        L.tgnh_num_nh_chains.argtypes = [vp]
        L.tgnh_get_kinetic_energies.argtypes = [vp, vp, dp]
        L.tgnh_kinetic_energy.argtypes = [vp, vp, dp]
        L.tgnh_compute_kinetic_energies.argtypes = [vp, vp, vp, dp]
        L.tgnh_get_chain_state.argtypes = [vp, vp, dp, dp, dp]
        L.tgnh_set_chain_state.argtypes = [vp, vp, dp, dp, dp]
        L.tgnh_get_vscale.argtypes = [vp, vp, dp]
        L.tgnh_get_thermostat_params.argtypes = [vp, dp, dp, dp]
        L.tgnh_launch_count.argtypes = [vp]
        L.tgnh_launch_count.restype = C.c_int64
        L.tgnh_exchange_kind.argtypes = [vp]
        L.tgnh_get_exchange_timing.argtypes = [vp, vp, dp]
        ip32 = C.POINTER(C.c_int32)
        L.tgnh_plan_tiles.argtypes = [C.POINTER(Params), ip32, C.c_int32, ip32, ip32, ip32]
        L.tgnh_plan_descriptors.argtypes = [C.POINTER(Params), C.POINTER(C.c_uint32)]
        L.tgnh_plan_chunks.argtypes = [C.POINTER(Params), ip32, C.c_int32, ip32, C.POINTER(C.c_uint8), C.POINTER(C.c_float), ip32, ip32]
        L.tgnh_kernel_generation.argtypes = [vp]
        L.tgnh_lazy_second_kick.argtypes = [vp]
        L.tgnh_residue_per_lane.argtypes = [vp]
        L.tgnh_set_profiling.argtypes = [vp, C.c_int]
        L.tgnh_get_profile.argtypes = [vp, dp, C.POINTER(C.c_int64)]
        L.tgnh_comm_get_unique_id.argtypes = [vp]
        L.tgnh_comm_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
        L.tgnh_comm_destroy.argtypes = [vp]
        L.tgnh_comm_destroy.restype = None
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        raise TgnhError(rc, lib().tgnh_last_error().decode())


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None and len(a) else None


class Comm:
    """NCCL communicator for particle-range sharding (one process per GPU)."""

    def __init__(self, unique_id: bytes, world_size: int, rank: int, device: int):
        h = C.c_void_p()
        buf = C.create_string_buffer(unique_id, UNIQUE_ID_BYTES)
        check(lib().tgnh_comm_create(buf, world_size, rank, device, C.byref(h)))
        self.h, self.world_size, self.rank = h, world_size, rank

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(UNIQUE_ID_BYTES)
        check(lib().tgnh_comm_get_unique_id(buf))
        return buf.raw

    def close(self):
        if self.h:
            lib().tgnh_comm_destroy(self.h)
            self.h = None


def make_params(system, *, force_format=FORCE_F32_SOA, precision=PRECISION_SINGLE, padded=None, device=-1, comm=None,
                has_cm_motion_remover=False, constraints=None, **overrides):
    """tgnh_params for a synth.DrudeSystem.  Returns (params, arrays that must outlive it, padded particle count)."""
    s = system
    n = s.num_particles
    padded = padded or ((n + 31) // 32) * 32
    cons = s.constraints if constraints is None else constraints
    cons = np.ascontiguousarray(cons, np.int32).reshape(-1, 2)
    k = dict(
        masses=np.ascontiguousarray(s.masses, np.float64), pd=np.ascontiguousarray(s.pair_drude, np.int32),
        pp=np.ascontiguousarray(s.pair_parent, np.int32), tg=np.ascontiguousarray(s.temp_group, np.int32),
        res=np.ascontiguousarray(s.res_id, np.int32), c0=np.ascontiguousarray(cons[:, 0]),
        c1=np.ascontiguousarray(cons[:, 1]))
    p = Params(
        num_particles=n, padded_num_particles=padded, num_pairs=len(k["pd"]), num_residues=s.num_residues,
        num_temp_groups=s.num_temp_groups, num_constraints=len(cons), num_nh_chains=s.num_nh_chains,
        drude_steps_per_real_step=s.drude_steps, use_drude_nh_chains=int(s.use_drude_nh_chains),
        use_com_temp_group=int(s.use_com_temp_group), has_cm_motion_remover=int(has_cm_motion_remover),
        force_format=force_format, device=device, precision=precision, temperature=s.temperature, coupling_time=s.coupling_time,
        drude_temperature=s.drude_temperature, drude_coupling_time=s.drude_coupling_time, step_size=s.step_size,
        max_drude_distance=s.max_drude_distance, masses=_dp(k["masses"]), pair_drude=_ip(k["pd"]),
        pair_parent=_ip(k["pp"]), particle_temp_group=_ip(k["tg"]), particle_res_id=_ip(k["res"]),
        constraint_p=_ip(k["c0"]), constraint_p1=_ip(k["c1"]), comm=comm.h if comm is not None else None)
    for key, val in overrides.items():
        setattr(p, key, val)
    return p, k, padded


def plan_tiles(system, **kw):
    """tgnh_plan_tiles: the host-side plan of tgnh_create without a device.  Returns (tile_start[num_tiles + 1],
    number of big residues, residue_uniform); raises TgnhError for every table error tgnh_create would report."""
    p, keep, _ = make_params(system, **kw)
    nt, nb, uni = C.c_int32(), C.c_int32(), C.c_int32()
    check(lib().tgnh_plan_tiles(C.byref(p), None, 0, C.byref(nt), C.byref(nb), C.byref(uni)))
    ts = np.zeros(nt.value + 1, np.int32)
    check(lib().tgnh_plan_tiles(C.byref(p), ts.ctypes.data_as(C.POINTER(C.c_int32)), len(ts), C.byref(nt), C.byref(nb), C.byref(uni)))
    return ts, nb.value, bool(uni.value)


def plan_chunks(system, **kw):
    """tgnh_plan_chunks: the warp-chunk plan (csrc/tgnh_v2.cuh), host-only.  Returns (chunk_start[num_chunks + 1],
    species[N] uint8, table[256, 8] float32, number of species, longest residue); raises TgnhError(ERR_UNSUPPORTED) for
    systems that run through the first-generation kernels."""
    p, keep, _ = make_params(system, **kw)
    nc, ns, mr = C.c_int32(), C.c_int32(), C.c_int32()
    ip32 = C.POINTER(C.c_int32)
    check(lib().tgnh_plan_chunks(C.byref(p), None, 0, C.byref(nc), None, None, C.byref(ns), C.byref(mr)))
    cs = np.zeros(nc.value + 1, np.int32)
    spec = np.zeros(system.num_particles, np.uint8)
    table = np.zeros((256, 8), np.float32)
    check(lib().tgnh_plan_chunks(C.byref(p), cs.ctypes.data_as(ip32), len(cs), C.byref(nc), spec.ctypes.data_as(C.POINTER(C.c_uint8)),
                                 table.ctypes.data_as(C.POINTER(C.c_float)), C.byref(ns), C.byref(mr)))
    return cs, spec, table, ns.value, mr.value


def plan_descriptors(system, **kw):
    """tgnh_plan_descriptors: the per-particle descriptor words (uint32[N]), host-only."""
    p, keep, _ = make_params(system, **kw)
    out = np.zeros(system.num_particles, np.uint32)
    check(lib().tgnh_plan_descriptors(C.byref(p), out.ctypes.data_as(C.POINTER(C.c_uint32))))
    return out


class Handle:
    """Owns one tgnh_handle.  Buffers are passed as raw device pointers (ints), e.g. tensor.data_ptr()."""

    def __init__(self, system, *, force_format=FORCE_F32_SOA, precision=PRECISION_SINGLE, padded=None, device=-1, comm=None,
                 has_cm_motion_remover=False, constraints=None, **overrides):
        p, self._keep, padded = make_params(system, force_format=force_format, precision=precision, padded=padded, device=device, comm=comm,
                                            has_cm_motion_remover=has_cm_motion_remover, constraints=constraints, **overrides)
        self.padded = padded
        self.num_particles = system.num_particles
        self.M = p.num_nh_chains
        h = C.c_void_p()
        check(lib().tgnh_create(C.byref(p), C.byref(h)))
        self.h = h
        self.T = lib().tgnh_num_thermostats(h)

    def close(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.tgnh_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:       # interpreter shutdown
            pass

    # ---- the step ----
    def half1(self, velm, posq, force, stream=0):
        check(lib().tgnh_half1(self.h, stream, velm, posq, force))

    def half1_kick(self, velm, force, pos_delta, stream=0):
        check(lib().tgnh_half1_kick(self.h, stream, velm, force, pos_delta))

    def half1_drift(self, velm, posq, pos_delta, stream=0):
        check(lib().tgnh_half1_drift(self.h, stream, velm, posq, pos_delta))

    def thermostat(self, velm, flags=HALF2_DEFAULT, stream=0):
        check(lib().tgnh_thermostat(self.h, stream, velm, flags))

    def half2(self, velm, force, flags=HALF2_DEFAULT, stream=0):
        check(lib().tgnh_half2(self.h, stream, velm, force, flags))

    def flush(self, velm, stream=0):
        check(lib().tgnh_flush(self.h, stream, velm))

    def step(self, velm, posq, force, nsteps=1, stream=0):
        check(lib().tgnh_step(self.h, stream, velm, posq, force, nsteps))

    def step_host(self, velm_host, posq_host, force_host, nsteps=1):
        """host numpy arrays (or pinned pointers as ints); returns the 2*KE vector consumed by the last chain update"""
        ke2 = np.zeros(self.T)
        as_ptr = lambda a: a if isinstance(a, int) else a.ctypes.data
        check(lib().tgnh_step_host(self.h, as_ptr(velm_host), as_ptr(posq_host), as_ptr(force_host), nsteps, _dp(ke2)))
        return ke2

    def step_host2(self, velm_host, posq_host, force_host, nsteps=1, posq_correction_host=None, forces_unchanged=False):
        """tgnh_step_host2: as step_host, plus the mixed layout's posqCorrection array and the forces-unchanged promise"""
        ke2 = np.zeros(self.T)
        as_ptr = lambda a: None if a is None else (a if isinstance(a, int) else a.ctypes.data)
        check(lib().tgnh_step_host2(self.h, as_ptr(velm_host), as_ptr(posq_host), as_ptr(posq_correction_host), as_ptr(force_host), nsteps,
                                    HOST_FORCES_UNCHANGED if forces_unchanged else 0, _dp(ke2)))
        return ke2

    def set_posq_correction(self, ptr):
        check(lib().tgnh_set_posq_correction(self.h, ptr))

    def invalidate(self):
        check(lib().tgnh_invalidate(self.h))

    # ---- thermostat state ----
    def kinetic_energies(self, stream=0):
        out = np.zeros(self.T); check(lib().tgnh_get_kinetic_energies(self.h, stream, _dp(out))); return out

    def kinetic_energy(self, stream=0):
        out = C.c_double(); check(lib().tgnh_kinetic_energy(self.h, stream, C.byref(out))); return out.value

    def compute_kinetic_energies(self, velm, stream=0):
        out = np.zeros(self.T); check(lib().tgnh_compute_kinetic_energies(self.h, stream, velm, _dp(out))); return out

    def vscale(self, stream=0):
        out = np.zeros(self.T); check(lib().tgnh_get_vscale(self.h, stream, _dp(out))); return out

    def chain_state(self, stream=0):
        eta = np.zeros((self.T, self.M)); ed = np.zeros((self.T, self.M + 1)); edd = np.zeros((self.T, self.M))
        check(lib().tgnh_get_chain_state(self.h, stream, _dp(eta), _dp(ed), _dp(edd)))
        return eta, ed, edd

    def set_chain_state(self, eta, ed, edd, stream=0):
        eta = np.ascontiguousarray(eta, np.float64); ed = np.ascontiguousarray(ed, np.float64)
        edd = np.ascontiguousarray(edd, np.float64)
        check(lib().tgnh_set_chain_state(self.h, stream, _dp(eta), _dp(ed), _dp(edd)))

    def thermostat_params(self):
        dof = np.zeros(self.T); nkbt = np.zeros(self.T); q = np.zeros((self.T, self.M))
        check(lib().tgnh_get_thermostat_params(self.h, _dp(dof), _dp(nkbt), _dp(q)))
        return dof, nkbt, q

    def set_profiling(self, enabled=True):
        check(lib().tgnh_set_profiling(self.h, int(enabled)))

    def profile(self):
        """{'half1': (ms, launches), 'half2': ..., 'reduce': ...} accumulated since the last call"""
        ms = np.zeros(3); cnt = np.zeros(3, np.int64)
        check(lib().tgnh_get_profile(self.h, _dp(ms), cnt.ctypes.data_as(C.POINTER(C.c_int64))))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(("half1", "half2", "reduce"))}

    @property
    def launch_count(self):
        return lib().tgnh_launch_count(self.h)

    def exchange_timing(self, stream=0):
        """(wait for all ranks' sums, publish -> wait start) of the most recent reduction, microseconds"""
        out = np.zeros(2); check(lib().tgnh_get_exchange_timing(self.h, stream, _dp(out))); return float(out[0]), float(out[1])

    @property
    def kernel_generation(self):
        """2 = the two halves run through the warp-chunk kernels (tgnh_v2.cuh), 1 = first-generation kernels"""
        return lib().tgnh_kernel_generation(self.h)

    @property
    def residue_per_lane(self):
        """k > 0: the reducing launches take a whole k-particle residue per lane (tgnh.h: tgnh_residue_per_lane)"""
        return lib().tgnh_residue_per_lane(self.h)

    @property
    def lazy_second_kick(self):
        """True when step(n) leaves the second half kick of all steps but the last to the next first half (tgnh.h)"""
        return bool(lib().tgnh_lazy_second_kick(self.h))

    @property
    def exchange_kind(self):
        """0 = not sharded, 1 = NCCL all-reduce, 2 = peer-mapped inboxes over NVLink"""
        return lib().tgnh_exchange_kind(self.h)
