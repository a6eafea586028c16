"""The C++ host side (plugin/): CPU unit tests through its own test executable, and — on a GPU — the whole stack
DrudeTGNHIntegrator -> KernelImpl -> C-ABI -> kernels against the oracle, driven like a user script
(Context::setPositions / setVelocities, integrator.step, Context::getState)."""
import os
import subprocess

import numpy as np
import pytest

from openmm_drudenose_b200 import synth
from oracle import oracle as O
from util import rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plugin_cpp_unit_tests():
    """plugin/tests/test_plugin.cpp: constructor / setters, temperature-group validation, XML (v1 byte-compatible with the
    reference's proxy, v2 complete), plugin registration, and the loud failure without a B200."""
    exe = os.path.join(ROOT, "plugin", "tests", "test_plugin")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip().endswith("Done"), out.stdout + out.stderr


def test_plugin_libraries_export_the_openmm_entry_points():
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "plugin", "libDrudeTGNHPluginB200.so"))
    for sym in ("registerPlatforms", "registerKernelFactories", "registerDrudeTGNHCudaKernelFactories"):
        assert hasattr(lib, sym)
    api = ctypes.CDLL(os.path.join(ROOT, "plugin", "libOpenMMDrudeTGNH.so"))
    assert hasattr(api, "registerDrudeTGNHSerializationProxies")


@pytest.mark.gpu
@pytest.mark.parametrize("model", [0, 1])
def test_plugin_stack_against_oracle(cuda, model):
    """20 steps through the integrator class with OpenMM-format int64 forces; forces re-evaluated by the Context each step
    (model 1: Drude springs from the current positions)."""
    from plugin_driver import PluginSim
    kw = dict(quantize_masses=True)
    if model:
        kw.update(pair_force="none", cold_drudes=True, drude_sigma=1.4e-4, force_sigma=2.0)
    s = synth.water_box(1500, 3, **kw)
    s.positions = s.positions - s.positions.mean(0)          # small coordinates: fp32 resolves the Drude displacement
    s.positions = s.positions.astype(np.float32).astype(np.float64)
    sim = PluginSim(s, force_model=model)
    o = O.Oracle(s, O.TG)
    pa, va = s.positions.copy(), s.velocities.copy()
    pb, vb = pa.copy(), va.copy()
    ext = np.rint(s.forces * 4294967296.0) / 4294967296.0
    fa = O.harmonic_forces(s, pa, ext) if model else ext.copy()
    fb = fa.copy()
    steps = 20
    sim.step(pa, va, fa, steps, ext if model else None)
    o.step(pb, vb, fb, steps, model, ext if model else None, s.k_spring if model else None)
    # 20 free-running fp32 steps; model 1 feeds fp32 positions back into the spring forces: the 1.4e-4 nm Drude displacement
    # sits on coordinates of a few nm (ulp 2.4e-7), so every force carries ~0.2 % noise (the fp32 layout, not the kernels)
    assert rel_err(va, vb) < (2e-3 if model else 2e-5)
    assert rel_err(pa, pb) < 1e-5
    ke = sim.kinetic_energy()
    assert abs(ke - o.ke_sum) / o.ke_sum < (1e-4 if model else 1e-6)
    sim.close()


@pytest.mark.gpu
def test_plugin_constrained_system_takes_the_split_sequence(cuda):
    """A System with constraints (C2 shape: SWM4 waters, 3 constraints each, massless M site, CMMotionRemover): the KernelImpl
    calls kick -> applyConstraints -> drift -> forces -> kick -> applyVelocityConstraints -> thermostat
    (CudaDrudeTGNHKernels.cpp:356-402).  The shim's constraint calls are identities, so the result equals the oracle with
    the constraints entering only the DOF bookkeeping."""
    from plugin_driver import PluginSim
    s = synth.swm4_box(800, quantize_masses=True)
    s.positions = (s.positions - s.positions.mean(0)).astype(np.float32).astype(np.float64)
    sim = PluginSim(s, has_cm_motion_remover=True, with_constraints=True)
    o = O.Oracle(s, O.TG, constraints=s.constraints, has_cm_motion_remover=True)
    ext = np.rint(s.forces * 4294967296.0) / 4294967296.0
    pa, va, fa = s.positions.copy(), s.velocities.copy(), ext.copy()
    pb, vb, fb = pa.copy(), va.copy(), ext.copy()
    sim.step(pa, va, fa, 10)
    o.step(pb, vb, fb, 10)
    assert sim.constraint_calls() == 20
    assert rel_err(va, vb) < 2e-5 and rel_err(pa, pb) < 1e-5
    assert abs(sim.kinetic_energy() - o.ke_sum) / o.ke_sum < 1e-6
    sim.close()


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["mixed", "double"])
@pytest.mark.parametrize("constrained", [False, True])
def test_plugin_stack_mixed_precision(cuda, constrained, precision):
    """Context created with Precision=mixed (what example/nacl_tg.py:60 asks for): double4 velm, posqCorrection, int64
    forces recomputed from the positions each step (Drude springs).  100 steps agree with the oracle to 1e-9 — the
    spring forces see positions through posq + posqCorrection (~48 bits)."""
    from plugin_driver import PluginSim
    kw = dict(quantize_masses=True, pair_force="none", cold_drudes=True, drude_sigma=1.4e-4, force_sigma=2.0)
    s = synth.swm4_box(600, **kw) if constrained else synth.water_box(1500, 3, **kw)
    sim = PluginSim(s, force_model=1, precision=precision, with_constraints=constrained, has_cm_motion_remover=constrained)
    o = O.Oracle(s, O.TG, constraints=s.constraints if constrained else None, has_cm_motion_remover=constrained)
    pa, va = s.positions.copy(), s.velocities.copy()
    pb, vb = pa.copy(), va.copy()
    ext = np.rint(s.forces * 4294967296.0) / 4294967296.0
    fa = O.harmonic_forces(s, pa, ext)
    fb = fa.copy()
    sim.step(pa, va, fa, 100, ext)
    o.step(pb, vb, fb, 100, 1, ext, s.k_spring)
    assert rel_err(va, vb) < 1e-6          # int64 force quantisation (2^-32) + float-float positions feeding back for 100 steps
    assert rel_err(pa, pb) < 1e-7
    assert abs(sim.kinetic_energy() - o.ke_sum) / o.ke_sum < 1e-8
    if constrained:
        assert sim.constraint_calls() == 200
    sim.close()


@pytest.mark.gpu
def test_plugin_single_pair_hard_wall(cuda):
    """The reference tests' 2-particle system (testSinglePair): the hard wall bound holds at every sample."""
    from plugin_driver import PluginSim
    sp = synth.single_pair()
    sim = PluginSim(sp, force_model=1)
    p, v = sp.positions.copy(), sp.velocities.copy()
    f = O.harmonic_forces(sp, p)
    for _ in range(200):
        sim.step(p, v, f, 10)
        assert np.linalg.norm(p[0] - p[1]) <= sp.max_drude_distance * (1 + 1e-4)
    sim.close()


def test_python_module_surface():
    """`drudetgnhplugin` (pybind11 stand-in for the SWIG module): class, method list, Python-side defaults of
    python/drudetgnhplugin.i:60-92 — note useDrudeNHChains defaults to True in Python, False in C++ (SURVEY.md finding 3)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "plugin", "python"))
    import drudetgnhplugin as dp
    integ = dp.DrudeTGNHIntegrator(300.0, 0.1, 1.0, 0.005, 0.001, 20)          # the README's call: 6th argument = drudeStepsPerRealStep
    assert float(integ.getTemperature()) == 300.0 and float(integ.getDrudeCouplingTime()) == 0.005
    assert integ.getDrudeStepsPerRealStep() == 20 and integ.getNumNHChains() == 1
    assert integ.getUseDrudeNHChains() == 1 and integ.getUseCOMTempGroup() == 1
    integ.setMaxDrudeDistance(0.02)
    assert float(integ.getMaxDrudeDistance()) == 0.02
    with pytest.raises(dp.OpenMMException, match="cannot be negative"):
        integ.setMaxDrudeDistance(-0.1)
    methods = ["getTemperature", "setTemperature", "getCouplingTime", "setCouplingTime", "getDrudeTemperature", "setDrudeTemperature",
               "getDrudeCouplingTime", "setDrudeCouplingTime", "getMaxDrudeDistance", "setMaxDrudeDistance", "step",
               "getDrudeStepsPerRealStep", "setDrudeStepsPerRealStep", "getNumNHChains", "setNumNHChains", "getUseDrudeNHChains",
               "setUseDrudeNHChains", "getUseCOMTempGroup", "setUseCOMTempGroup", "getNumTempGroups", "addTempGroup",
               "addParticleTempGroup", "setParticleTempGroup", "getParticleTempGroup"]
    assert all(hasattr(integ, m) for m in methods)
    # temperature groups: the first addTempGroup() returns 0; the README's "default group 0, added group 1" recipe throws (D11)
    assert integ.addTempGroup() == 0
    with pytest.raises(dp.OpenMMException, match="Index out of range"):
        integ.addParticleTempGroup(1)
    assert integ.addTempGroup() == 1
    assert integ.addParticleTempGroup(1) == 0 and integ.getParticleTempGroup(0) == 1
    with pytest.raises(dp.OpenMMException, match="not bound to a context"):
        integ.step(1)
    assert 'version="1"' in dp.serialize(integ)              # default: the upstream plugin's format
    xml = dp.serialize(integ, version=2)
    copy = dp.deserialize(xml)
    assert copy.getNumTempGroups() == 2 and copy.getParticleTempGroup(0) == 1 and float(copy.getMaxDrudeDistance()) == 0.02


def _script_style_run(dp, s, precision, steps, springs):
    """What a user script does (example/nacl_tg.py:30-75), written against drudetgnhplugin + drudetgnhplugin.shim."""
    mm = dp.shim
    system = mm.System()
    for m in s.masses:
        system.addParticle(float(m))
    drude = mm.DrudeForce()
    for d, p in zip(s.pair_drude, s.pair_parent):
        drude.addParticle(int(d), int(p), -1, -1, -1, -1.0, 1.0, 1.0, 1.0)
    system.addForce(drude)
    bonds = mm.BondForce()                       # molecules = the integrator's residues
    for i in range(1, s.num_particles):
        if s.res_id[i] == s.res_id[i - 1]:
            bonds.addBond(i - 1, i)
    system.addForce(bonds)
    integ = dp.DrudeTGNHIntegrator(s.temperature, s.coupling_time, s.drude_temperature, s.drude_coupling_time, s.step_size, s.drude_steps,
                                   s.num_nh_chains, int(s.use_drude_nh_chains), int(s.use_com_temp_group))
    integ.setMaxDrudeDistance(s.max_drude_distance)
    for _ in range(s.num_temp_groups):
        integ.addTempGroup()
    for g in s.temp_group:
        integ.addParticleTempGroup(int(g))
    platform = mm.Platform.getPlatformByName("CUDA")
    context = mm.Context(system, integ, platform, {"Precision": precision})
    context.setPositions(s.positions)
    context.setVelocities(s.velocities)
    ext = np.rint(s.forces * 4294967296.0) / 4294967296.0
    if springs:
        context.setForceModel(ext, [int(x) for x in s.pair_drude], [int(x) for x in s.pair_parent], [float(k) for k in s.k_spring])
    else:
        context.setForceModel(ext)
    integ.step(steps)
    st = context.getState(getPositions=True, getVelocities=True, getEnergy=True)
    return st.getPositions(), st.getVelocities(), st.getKineticEnergy(), ext


def test_python_shim_surface_and_loud_failure_without_a_device():
    """The stand-in classes a script needs exist under OpenMM's names; without a CUDA device, creating the Context fails with an
    OpenMMException (no silent CPU path)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "plugin", "python"))
    import drudetgnhplugin as dp
    for name in ("System", "DrudeForce", "BondForce", "CMMotionRemover", "Platform", "Context", "State"):
        assert hasattr(dp.shim, name)
    assert dp.shim.Platform.getPlatformByName("CUDA").getName() == "CUDA"
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present: covered by test_python_script_runs_the_integrator_on_the_gpu")
    s = synth.water_box(8, 2)
    with pytest.raises(dp.OpenMMException, match="no CUDA device"):
        _script_style_run(dp, s, "single", 1, False)


@pytest.mark.gpu
@pytest.mark.parametrize("precision,springs", [("single", False), ("mixed", True)])
def test_python_script_runs_the_integrator_on_the_gpu(cuda, precision, springs):
    """integrator.step(n) from Python: System / DrudeForce / Context(platform "CUDA", Precision) / getState, the way
    example/nacl_tg.py drives the reference, through the pybind11 module; 20 steps against the oracle."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "plugin", "python"))
    import drudetgnhplugin as dp
    kw = dict(quantize_masses=True)
    if springs:
        kw.update(pair_force="none", cold_drudes=True, drude_sigma=1.4e-4, force_sigma=2.0)
    s = synth.water_box(1500, 3, **kw)
    s.positions = (s.positions - s.positions.mean(0)).astype(np.float32).astype(np.float64)
    if precision == "single":
        s.velocities = s.velocities.astype(np.float32).astype(np.float64)
    steps = 20
    p, v, ke, ext = _script_style_run(dp, s, precision, steps, springs)
    o = O.Oracle(s, O.TG)
    pb, vb = s.positions.copy(), s.velocities.copy()
    fb = O.harmonic_forces(s, pb, ext) if springs else ext.copy()
    o.step(pb, vb, fb, steps, 1 if springs else 0, ext if springs else None, s.k_spring if springs else None)
    tol_v, tol_x = (2e-5, 1e-5) if precision == "single" else (1e-6, 1e-7)
    assert rel_err(v, vb) < tol_v and rel_err(p, pb) < tol_x
    assert abs(ke - o.ke_sum) / o.ke_sum < (1e-6 if precision == "single" else 1e-7)
