"""The reference's OWN CUDA platform, executed (oracle/_refcuda, oracle/cuda_driver.cpp).

`platforms/cuda` is the only place where the reference implements temperature groups and the molecular-COM thermostat
(SURVEY.md finding 1).  Its sources — CudaDrudeTGNHKernels.cpp (host step driver, host fp64 Nose-Hoover chain), the kernel
strings vectorOps.cu + drudeTGNH.cu, CudaDrudeTGNHKernelFactory.cpp and openmmapi's DrudeTGNHIntegrator.cpp — are compiled
unmodified from /root/reference against shim/cuda (a stand-in for OpenMM's CudaContext / CudaArray / CudaIntegrationUtilities /
CudaPlatform with real device arrays, NVRTC behind OpenMM's kernel prelude and OpenMM's launch rule) and run on the B200.

  * the oracle's TG layer (what every other -m gpu test compares against) is pinned against that execution, per step;
  * this repo's plugin is driven by the reference's own, unmodified DrudeTGNHIntegrator on the same stand-in platform (the
    -DTGNH_WITH_OPENMM build of plugin/src/B200DrudeTGNHKernelFactory.cpp) and compared with the reference's CUDA platform directly,
    including Context::setVelocities between steps and atom reordering.
"""
import os

import numpy as np
import pytest

from openmm_drudenose_b200 import synth
from oracle import oracle as O
from util import group_temperatures, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_cuda_kernels_jit_for_sm_100a():
    """CPU: the reference's kernel strings (vectorOps + drudeTGNH, CudaDrudeTGNHKernels.cpp:269) compile with NVRTC for sm_100a behind
    OpenMM's prelude in single, mixed and double precision."""
    import subprocess
    exe = os.path.join(ROOT, "oracle", "_refcuda", "nvrtc_check")
    if not os.path.exists(exe):
        pytest.skip("oracle/_refcuda was not built (no /root/reference at build time)")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("Done"), out.stdout + out.stderr
    for mode in ("single", "mixed", "double"):
        assert f"{mode}: ok" in out.stdout


def test_reference_cuda_libraries_are_built():
    """CPU: both driver builds exist wherever the reference tree was mounted at build time (they need libcuda to load)."""
    if not os.path.exists("/root/reference") and not os.path.exists(os.path.join(ROOT, "oracle", "_refcuda", "librefcuda.so")):
        pytest.skip("no reference tree and no prebuilt oracle/_refcuda")
    for name in ("librefcuda.so", "libb200cuda.so"):
        assert os.path.exists(os.path.join(ROOT, "oracle", "_refcuda", name)), name


# ---------------------------------------------------------------------------------------------------------------------
gpu = pytest.mark.gpu
KW = dict(quantize_masses=True)
TOL_PIN = 1e-12       # oracle-TG against the reference's CUDA platform in double precision, per step


def _quantize_forces(s):
    s.forces = np.rint(s.forces * 4294967296.0) / 4294967296.0      # the platform's force buffer is fixed point, 2^-32
    return s


def _systems():
    return {
        "water_G4_com": lambda: synth.water_box(1500, 4, **KW),
        "water_G1_nocom": lambda: synth.water_box(1000, 1, use_com_temp_group=False, **KW),
        "water_M1_nodrudechain": lambda: synth.water_box(800, 2, num_nh_chains=1, use_drude_nh_chains=False, **KW),
        "water_M6": lambda: synth.water_box(500, 3, num_nh_chains=6, **KW),
        "nacl_C1": lambda: synth.nacl_box(**KW),
        "swm4_C2": lambda: synth.swm4_box(1500, **KW),
        "ionic_C3": lambda: synth.ionic_liquid(100, **KW),
        "ragged": lambda: synth.build([synth.WATER4, synth.SOD, synth.SWM4], np.arange(901) % 3, np.arange(901) % 2, 2, **KW),
        "split_groups": lambda: _split(synth.water_box(900, 2, **KW)),
        "tiny": lambda: synth.water_box(3, 1, **KW),
    }


def _split(s):
    tg = s.temp_group.copy()
    tg[2::4] = 1 - tg[2::4]                                # residues that span two temperature groups
    s.temp_group = tg
    return s


SYSTEMS = _systems()


def _refcuda():
    from oracle import refcuda
    if not refcuda.available("reference"):
        pytest.skip("oracle/_refcuda was not built (no /root/reference at build time)")
    return refcuda


@gpu
@pytest.mark.parametrize("name", sorted(SYSTEMS))
def test_oracle_tg_pinned_by_the_reference_cuda_platform(cuda, name):
    """Double precision, 4 consecutive free-running steps: positions, velocities, scale factors and every chain variable of the oracle's TG
    layer agree with the reference's own CUDA platform to 1e-12; the thermostat tables (N kT, Q) to 1e-14."""
    R = _refcuda()
    s = _quantize_forces(SYSTEMS[name]())
    sim = R.CudaSim(s, "reference", "double")
    o = O.Oracle(s, O.TG, constraints=s.constraints)
    nkbt, q = sim.thermostat_params()
    _, nkbt_o, q_o = o.thermostat_params()
    np.testing.assert_allclose(nkbt, nkbt_o, rtol=1e-14, atol=0)
    np.testing.assert_allclose(q, q_o, rtol=1e-14, atol=0)
    p, v, f = s.positions.copy(), s.velocities.copy(), s.forces.copy()
    sim.set_state(p, v, f)
    for step in range(4):
        sim.step(1)
        o.step(p, v, f, 1)
        gp, gv, _ = sim.get_state()
        assert rel_err(gv, v) < TOL_PIN, (step, rel_err(gv, v))
        assert rel_err(gp, p) < TOL_PIN, (step, rel_err(gp, p))
        eta, ed, edd, vs = sim.thermostat()
        eta_o, ed_o, edd_o = o.chain_state()
        np.testing.assert_allclose(vs, o.vscale, rtol=TOL_PIN)
        for got, ref in ((eta, eta_o), (ed, ed_o), (edd, edd_o)):
            np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-12 * max(np.abs(ref).max(), 1e-300))
        assert abs(sim.ke_sum - o.ke_sum) <= 1e-12 * abs(o.ke_sum)
    sim.close()


@gpu
def test_oracle_tg_pinned_hard_wall(cuda):
    """Pairs placed robustly inside / outside the wall, double precision: the reflection formulas (drudeTGNH.cu:487-572) of the oracle
    against the reference's kernel, 3 steps."""
    R = _refcuda()
    s = synth.water_box(2048, 2, quantize_masses=True, drude_sigma=0.0, pair_force="none", cold_drudes=True, force_sigma=5.0)
    rng = np.random.default_rng(7)
    npair = s.num_pairs
    direction = rng.standard_normal((npair, 3)); direction /= np.linalg.norm(direction, axis=1)[:, None]
    dist = np.where(np.arange(npair) % 2 == 0, rng.uniform(0.022, 0.035, npair), rng.uniform(0.001, 0.017, npair))
    s.positions[s.pair_drude] = s.positions[s.pair_parent] + direction * dist[:, None]
    _quantize_forces(s)
    sim = R.CudaSim(s, "reference", "double")
    o = O.Oracle(s, O.TG)
    p, v, f = s.positions.copy(), s.velocities.copy(), s.forces.copy()
    sim.set_state(p, v, f)
    for step in range(3):
        sim.step(1)
        o.step(p, v, f, 1)
        gp, gv, _ = sim.get_state()
        if step == 0:
            r = np.linalg.norm(gp[s.pair_drude] - gp[s.pair_parent], axis=1)
            assert (np.abs(r - dist) > 1e-3).sum() > npair // 3          # the wall really acted
        assert rel_err(gv, v) < 1e-11 and rel_err(gp, p) < 1e-11
    sim.close()


@gpu
def test_mixed_precision_reference_matches_oracle(cuda):
    """OpenMM's mixed mode (double velm, float posq + posqCorrection): the same agreement at the resolution of the float+float positions.
    (Pairs at the wall are excluded here: in mixed mode OpenMM's SQRT is sqrtf, so the reference's bond length has float precision.)"""
    R = _refcuda()
    s = _quantize_forces(synth.water_box(1500, 4, max_drude_distance=2.0, **KW))
    sim = R.CudaSim(s, "reference", "mixed")
    assert "#define USE_MIXED_PRECISION 1" in sim.kernel_source() and "#define SQRT sqrtf" in sim.kernel_source()
    o = O.Oracle(s, O.TG)
    p, v, f = s.positions.copy(), s.velocities.copy(), s.forces.copy()
    sim.set_state(p, v, f)
    sim.step(5)
    o.step(p, v, f, 5)
    gp, gv, _ = sim.get_state()
    assert rel_err(gv, v) < 1e-12 and rel_err(gp, p) < 1e-12
    sim.close()


@gpu
def test_two_fp64_implementations_over_1000_steps(cuda):
    """How far do two fp64 implementations of the same algorithm drift apart in 1000 steps?  The reference's CUDA platform (double
    precision, FMA contraction, its own summation order) against the oracle's TG layer (no contraction, particle order), G = 4, COM
    thermostat, M = 3, Drude chains ON.  Every thermostat except the Drude one agrees to 1e-9; the Drude thermostat's chain
    (tau = 5 fs, chaotic, tests/test_oracle.py::test_fp32_state_sensitivity) is the variable that no two implementations keep to 1e-6 —
    the measured divergence is recorded in DESIGN.md and bounds what the fp32-state tests can be asked to hold."""
    R = _refcuda()
    s = _quantize_forces(synth.water_box(5000, 4, quantize_masses=True, cold_drudes=True, drude_sigma=1.4e-4, force_sigma=2.0,
                                         max_drude_distance=2.0, use_drude_nh_chains=True))
    sim = R.CudaSim(s, "reference", "double")
    o = O.Oracle(s, O.TG)
    p, v, f = s.positions.copy(), s.velocities.copy(), s.forces.copy()
    sim.set_state(p, v, f)
    sim.step(1000)
    o.step(p, v, f, 1000)
    eta, ed, edd, vs = sim.thermostat()
    eta_o, ed_o, _ = o.chain_state()
    np.testing.assert_allclose(vs[:-1], o.vscale[:-1], rtol=1e-9)
    np.testing.assert_allclose(ed[:-1], ed_o[:-1], rtol=1e-6, atol=1e-9 * np.abs(ed_o[:-1]).max())
    np.testing.assert_allclose(eta[:-1], eta_o[:-1], rtol=1e-6, atol=1e-9 * np.abs(eta_o[:-1]).max())
    drude = float(np.max(np.abs(ed[-1] - ed_o[-1])) / np.abs(ed_o[-1]).max())
    print(f"\nDrude-thermostat chain after 1000 steps, reference CUDA (fp64) vs oracle-TG (fp64): max |d eta_dot| / max |eta_dot| = {drude:.3e}; "
          f"scale factor difference {abs(vs[-1] - o.vscale[-1]):.3e}")
    assert drude < 1e-2
    sim.close()


# ---- this repo's plugin under the reference's own integrator, against the reference's own CUDA platform ---------------------------------------
def _pair(R, s, precision_b200, precision_ref="double", **kw):
    """(reference CUDA platform, this repo's plugin) on the same system.  The reference runs in double precision unless asked otherwise:
    in its mixed mode OpenMM's SQRT macro is sqrtf, so the hard-wall kernel's bond length and direction (drudeTGNH.cu:488-494) carry
    float rounding (6e-8) into the velocities of every pair that meets the wall — a property of that build mode, not a target."""
    if not R.available("b200"):
        pytest.skip("oracle/_refcuda/libb200cuda.so was not built")
    return R.CudaSim(s, "reference", precision_ref, **kw), R.CudaSim(s, "b200", precision_b200, **kw)


@gpu
@pytest.mark.parametrize("name", ["water_G4_com", "nacl_C1", "ionic_C3", "split_groups", "water_G1_nocom"])
@pytest.mark.parametrize("precision", ["single", "mixed", "double"])
def test_plugin_under_reference_integrator_single_steps(cuda, name, precision):
    """libtgnh behind the reference's unmodified DrudeTGNHIntegrator and kernel interface (3 virtuals, no extension) against the
    reference's CUDA platform: 5 steps, per-step tolerance 1e-5 (single) / 1e-10 (mixed, double)."""
    R = _refcuda()
    s = _quantize_forces(SYSTEMS[name]())
    ref, mine = _pair(R, s, precision)
    p, v, f = s.positions.copy(), s.velocities.copy(), s.forces.copy()
    if precision == "single":
        p = p.astype(np.float32).astype(np.float64); v = v.astype(np.float32).astype(np.float64)
    ref.set_state(p, v, f); mine.set_state(p, v, f)
    ref.step(5); mine.step(5)
    rp, rv, _ = ref.get_state()
    mp, mv, _ = mine.get_state()
    tol_v, tol_x = (5e-5, 1e-5) if precision == "single" else (1e-10, 1e-10)
    assert rel_err(mv, rv) < tol_v and rel_err(mp, rp) < tol_x
    np.testing.assert_allclose(mine.thermostat()[3], ref.thermostat()[3], rtol=1e-6 if precision == "single" else 1e-10)
    c = mine.counters()
    assert c["initialize_contexts"] == 1 and c["force_evaluations"] >= 5
    if precision == "single" and name != "ionic_C3":
        assert mine.kernel_generation == 2
    ref.close(); mine.close()


@gpu
@pytest.mark.parametrize("drude_chain", [False, True])
def test_plugin_against_reference_cuda_1000_steps(cuda, drude_chain):
    """BASELINE.json's bar against the reference's own code, 1000 steps from identical state; the plugin in OpenMM's single-precision
    layout, the reference in its mixed mode (its single mode reads the double scale factors as floats).

    Group temperatures and thermostat scale factors: within 1e-6 (measured 7e-10 for the factors).  Chain variables: within 5e-6 of
    the chain's largest variable and 2e-5 per variable (measured 6e-8 ... 6.5e-6).  Why not 1e-6 per variable in THIS layout: with
    FIXED forces, fl32(v + dv) with v on the float grid and dv the same every step has the same rounding error every step (it
    depends only on dv mod ulp(v)), so per-particle errors persist instead of averaging out over time; summed over a group of 12 500
    molecules that is a scale error of ~1e-9 per step with one sign (reproduced in numpy without any kernel, DESIGN.md "Parity"), and
    a scale error of 1e-9 per step moves a thermostat's chain velocities by 1e-6 relative (COM thermostat 5e-6) — shown by
    perturbing the fp64 oracle.  The mixed layout (double velocities) holds 3e-13 on the same run (next test)."""
    R = _refcuda()
    s = _quantize_forces(synth.water_box(12500, 4, quantize_masses=True, cold_drudes=True, drude_sigma=1.4e-4, force_sigma=2.0,
                                         max_drude_distance=2.0, use_drude_nh_chains=drude_chain))
    ref, mine = _pair(R, s, "single", "mixed")
    p = s.positions.astype(np.float32).astype(np.float64); v = s.velocities.astype(np.float32).astype(np.float64)
    ref.set_state(p, v, s.forces); mine.set_state(p, v, s.forces)
    ref.step(1000); mine.step(1000)
    eta_r, ed_r, _, vs_r = ref.thermostat()
    eta_m, ed_m, _, vs_m = mine.thermostat()
    live = slice(0, -1)          # the Drude thermostat: stiff (tau = 5 fs); with a chain it is chaotic even between two fp64 codes (see above)
    np.testing.assert_allclose(vs_m[live], vs_r[live], rtol=1e-6)
    assert np.max(np.abs(ed_m[live] - ed_r[live])) < 5e-6 * np.abs(ed_r[live]).max()
    assert np.max(np.abs(eta_m[live] - eta_r[live])) < 5e-6 * np.abs(eta_r[live]).max()
    nz = np.abs(ed_r[live]) > 0
    assert np.max(np.abs(ed_m[live] - ed_r[live])[nz] / np.abs(ed_r[live])[nz]) < 2e-5
    if not drude_chain:
        assert abs(vs_m[-1] - vs_r[-1]) < 1e-6
    # temperatures: from the velocities both platforms hold at the end (the reference keeps no per-group energies)
    o = O.Oracle(s, O.TG)
    dof = o.thermostat_params()[0]
    t_r = group_temperatures(o.compute_ke2(np.ascontiguousarray(ref.get_state()[1])), dof)
    t_m = group_temperatures(o.compute_ke2(np.ascontiguousarray(mine.get_state()[1])), dof)
    np.testing.assert_allclose(t_m[live], t_r[live], rtol=1e-6)
    ref.close(); mine.close()


@gpu
def test_plugin_mixed_against_reference_cuda_1000_steps(cuda):
    """The same 1000 steps with the plugin in OpenMM's mixed layout against the reference in double precision: every chain variable of the
    particle thermostats within 1e-11 (measured 3e-13), the Drude thermostat's (no chain) within 1e-7."""
    R = _refcuda()
    s = _quantize_forces(synth.water_box(12500, 4, quantize_masses=True, cold_drudes=True, drude_sigma=1.4e-4, force_sigma=2.0,
                                         max_drude_distance=2.0, use_drude_nh_chains=False))
    ref, mine = _pair(R, s, "mixed", "double")
    p = s.positions.astype(np.float32).astype(np.float64); v = s.velocities.astype(np.float32).astype(np.float64)
    ref.set_state(p, v, s.forces); mine.set_state(p, v, s.forces)
    ref.step(1000); mine.step(1000)
    eta_r, ed_r, _, vs_r = ref.thermostat()
    eta_m, ed_m, _, vs_m = mine.thermostat()
    np.testing.assert_allclose(vs_m, vs_r, rtol=1e-12)
    np.testing.assert_allclose(ed_m[:-1, :-1], ed_r[:-1, :-1], rtol=1e-11)
    np.testing.assert_allclose(eta_m[:-1], eta_r[:-1], rtol=1e-11)
    np.testing.assert_allclose(ed_m[-1, 0], ed_r[-1, 0], rtol=1e-7)
    assert rel_err(mine.get_state()[1], ref.get_state()[1]) < 1e-9
    ref.close(); mine.close()


@gpu
@pytest.mark.parametrize("precision", ["single", "mixed"])
def test_set_velocities_between_steps(cuda, precision):
    """Context::setVelocities between steps, through the reference's own DrudeTGNHIntegrator::stateChanged (which knows nothing of this
    repo's kernel): the next thermostat half-step must see the new velocities."""
    R = _refcuda()
    s = _quantize_forces(synth.water_box(1500, 4, **KW))
    ref, mine = _pair(R, s, precision)
    p = s.positions.astype(np.float32).astype(np.float64); v = s.velocities.astype(np.float32).astype(np.float64)
    ref.set_state(p, v, s.forces); mine.set_state(p, v, s.forces)
    ref.step(3); mine.step(3)
    hot = (1.5 * ref.get_state()[1]).astype(np.float32).astype(np.float64)       # 2.25 x the kinetic energy
    ref.set_velocities(hot); mine.set_velocities(hot)
    ref.step(1); mine.step(1)
    np.testing.assert_allclose(mine.thermostat()[3], ref.thermostat()[3], rtol=1e-6)
    assert abs(mine.ke_sum / ref.ke_sum - 1) < 1e-6
    ref.step(2); mine.step(2)
    assert rel_err(mine.get_state()[1], ref.get_state()[1]) < (1e-4 if precision == "single" else 1e-10)
    ref.close(); mine.close()


@gpu
@pytest.mark.parametrize("precision", ["single", "mixed"])
def test_atom_reordering(cuda, precision):
    """cu.reorderAtoms() at the end of a step moves molecules to other slots while the force buffer keeps the old order; the next execute
    must refresh the forces (CudaDrudeTGNHKernels.cpp:344-347).  Per-particle fixed forces that differ between the molecules that trade
    places: a stale buffer would kick every swapped molecule with its neighbour's forces."""
    R = _refcuda()
    # two temperature groups in blocks, of different molecule types: neighbours inside a block are interchangeable, the two molecules
    # at the block boundary are not (the reference registers no CudaForceInfo: OpenMM would happily swap equal molecules of different
    # temperature groups under it)
    half = np.arange(1000) >= 500
    s = _quantize_forces(synth.build([synth.WATER4, synth.SWM4], half.astype(int), half.astype(int), 2, **KW))
    p = s.positions.astype(np.float32).astype(np.float64); v = s.velocities.astype(np.float32).astype(np.float64)
    plain = R.CudaSim(s, "reference", "double")
    ref, mine = _pair(R, s, precision, reorder_interval=2)
    for sim in (plain, ref, mine):
        sim.set_state(p, v, s.forces)
        sim.step(7)
    assert ref.counters()["reorders"] >= 3 and mine.counters()["reorders"] >= 3
    assert mine.counters()["force_evaluations"] > plain.counters()["force_evaluations"]
    pp, pv, _ = plain.get_state()
    rp, rv, _ = ref.get_state()
    mp, mv, _ = mine.get_state()
    assert rel_err(rv, pv) < 1e-12 and rel_err(rp, pp) < 1e-12              # reordering is a relabelling for the reference
    tol = 1e-4 if precision == "single" else 1e-10
    assert rel_err(mv, pv) < tol and rel_err(mp, pp) < (1e-5 if precision == "single" else 1e-10)
    for sim in (plain, ref, mine):
        sim.close()


@gpu
def test_force_info_keeps_temperature_groups_apart(cuda):
    """The CudaForceInfo this repo's kernel registers (tgnh_plan_descriptors) calls molecules of different temperature groups different:
    with groups alternating molecule by molecule no neighbouring pair may trade places."""
    R = _refcuda()
    if not R.available("b200"):
        pytest.skip("oracle/_refcuda/libb200cuda.so was not built")
    s = _quantize_forces(synth.water_box(200, 2, **KW))           # molecule k -> group k mod 2
    mine = R.CudaSim(s, "b200", "single", reorder_interval=1)
    mine.set_state(s.positions, s.velocities, s.forces)
    mine.step(4)
    assert mine.counters()["reorders"] == 0
    mine.close()
