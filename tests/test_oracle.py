"""CPU tests of the oracle (oracle/): golden vectors, the two layers on their overlap domain, the reference
tests' statistical invariants, DOF bookkeeping and error behaviour.  No GPU."""
import os

import numpy as np
import pytest

from openmm_drudenose_b200 import synth
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_boltz_constant_is_shared():
    from openmm_drudenose_b200 import capi
    assert O.BOLTZ == capi.BOLTZ == synth.BOLTZ == 1.380649e-23 * 6.02214076e23 / 1000.0


def test_golden_full_step():
    g = np.load(os.path.join(GOLD, "tgnh_golden.npz"))
    s = synth.water_box(int(g["molecules"]), int(g["groups"]), quantize_masses=True)
    np.testing.assert_array_equal(s.positions, g["positions0"])
    np.testing.assert_array_equal(s.velocities, g["velocities0"])
    o = O.Oracle(s, O.TG)
    p, v, f = s.positions.copy(), s.velocities.copy(), s.forces.copy()
    np.testing.assert_allclose(o.compute_ke2(v.copy()), g["ke2_initial"], rtol=1e-13)
    o.step(p, v, f, int(g["steps"]))
    np.testing.assert_allclose(p, g["positions"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(v, g["velocities"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(o.ke2, g["ke2"], rtol=1e-12)
    np.testing.assert_allclose(o.vscale, g["vscale"], rtol=1e-13)
    np.testing.assert_allclose(o.chain_state()[1], g["eta_dot"], rtol=1e-10, atol=1e-14)
    dof, nkbt, q = o.thermostat_params()
    np.testing.assert_array_equal(dof, g["dof"]); np.testing.assert_array_equal(q, g["eta_mass"])


def test_golden_chain_known_answers():
    g = np.load(os.path.join(GOLD, "chain_golden.npz"))
    s = synth.water_box(int(g["molecules"]), int(g["groups"]), quantize_masses=True, num_nh_chains=int(g["chains"]))
    o = O.Oracle(s, O.TG)
    v = s.velocities.copy()
    for i, rec in enumerate(g["records"]):
        v *= (1.0 + 0.25 * i)
        o.propagate_nh_chain(v)
        got = np.concatenate([o.ke2, o.vscale, o.chain_state()[1].ravel()])
        np.testing.assert_allclose(got, rec, rtol=1e-11, atol=1e-13)


def test_layers_agree_on_overlap_domain():
    """G = 1, no COM group, Drude chains on, no constraints: platforms/reference and platforms/cuda coincide
    (SURVEY.md finding 2).  The two restatements keep each platform's own operation order, so they agree to
    rounding, not bit for bit."""
    g = np.load(os.path.join(GOLD, "overlap_golden.npz"))
    np.testing.assert_allclose(g["velocities_tg"], g["velocities_ref"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(g["positions_tg"], g["positions_ref"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(g["vscale_tg"][[0, 2]], g["vscale_ref"][[0, 2]], rtol=1e-12)
    np.testing.assert_allclose(g["eta_dot_tg"][[0, 2]], g["eta_dot_ref"][[0, 2]], rtol=1e-9, atol=1e-12)
    # and the live oracle reproduces the stored vectors
    s = synth.water_box(int(g["molecules"]), 1, use_com_temp_group=False, quantize_masses=True, pair_force="none",
                        cold_drudes=True, drude_sigma=1.4e-4, force_sigma=0.0)
    for name, which in (("tg", O.TG), ("ref", O.REF)):
        o = O.Oracle(s, which)
        p, v = s.positions.copy(), s.velocities.copy()
        f = O.harmonic_forces(s, p)
        o.step(p, v, f, int(g["steps"]), O.FORCE_HARMONIC, None, s.k_spring)
        np.testing.assert_allclose(v, g[f"velocities_{name}"], rtol=0, atol=1e-12)


def test_single_pair_statistics():
    """testSinglePair (platforms/cuda/tests/TestCudaDrudeTGNHIntegrator.cpp:54-109), shortened: mean internal
    kinetic energy = 1.5 kB T_drude within a few %, mean COM kinetic energy = 1.5 kB T within 15 %, and the
    hard wall bound distance <= r_max (1 + 1e-6) at every sample."""
    sp = synth.single_pair()
    o = O.Oracle(sp, O.TG)
    p, v = sp.positions.copy(), sp.velocities.copy()
    f = O.harmonic_forces(sp, p)
    o.step(p, v, f, 1000, O.FORCE_HARMONIC, None, sp.k_spring)
    m1, m2 = 1.0, 0.1
    ke_cm = ke_int = 0.0
    n = 4000
    for _ in range(n):
        o.step(p, v, f, 10, O.FORCE_HARMONIC, None, sp.k_spring)
        vcm = (v[0] * m1 + v[1] * m2) / (m1 + m2)
        ke_cm += 0.5 * (m1 + m2) * vcm @ vcm
        vi = v[0] - v[1]
        ke_int += 0.5 * (m1 * m2 / (m1 + m2)) * vi @ vi
        assert np.linalg.norm(p[0] - p[1]) <= sp.max_drude_distance * (1 + 1e-6)
    assert abs(ke_cm / n / (1.5 * O.BOLTZ * 300.0) - 1) < 0.15
    assert abs(ke_int / n / (1.5 * O.BOLTZ * 10.0) - 1) < 0.03


def test_kinetic_energy_scaling_identity():
    """KE'_g = s_g^2 KE_g for residue-uniform temperature groups: what the device path's deferred scaling relies on."""
    s = synth.water_box(500, 3, quantize_masses=True)
    o = O.Oracle(s, O.TG)
    v = s.velocities.copy()
    o.propagate_nh_chain(v)
    before, sc = o.ke2, o.vscale
    after = o.compute_ke2(v)
    np.testing.assert_allclose(after, before * sc ** 2, rtol=1e-12)


def test_dof_bookkeeping():
    """CudaDrudeTGNHKernels.cpp:114-212: 3 per massive particle, -3 per pair, -1 per constraint, COM group 3R (-3
    with a CMMotionRemover), COM share 3 m / M_res removed from the relative groups."""
    s = synth.swm4_box(50)                                  # 5 sites, one massless, 3 constraints per molecule
    o = O.Oracle(s, O.TG, constraints=s.constraints, has_cm_motion_remover=True)
    dof, nkbt, q = o.thermostat_params()
    assert dof[1] == 3 * 50 - 3 and dof[2] == 3 * 50
    assert abs(dof[0] - (50 * (4 * 3 - 3 - 3) - 3 * 50)) < 1e-9       # 12 particle dof - 3 (pair) - 3 (constraints) - 3 (COM share)
    np.testing.assert_allclose(nkbt[:2], dof[:2] * O.BOLTZ * 300.0)
    o2 = O.Oracle(s, O.REF, constraints=s.constraints, has_cm_motion_remover=True)
    dof2 = o2.thermostat_params()[0]
    assert dof2[0] == 50 * (12 - 3) - 150 - 3                           # D3: everything lands on the one real thermostat


def test_error_behaviour():
    s = synth.water_box(10, 2)
    s.temp_group = s.temp_group.copy()
    s.temp_group[1] = 1 - s.temp_group[1]                   # Drude in another group than its parent
    with pytest.raises(O.OracleError, match="Temperature group for drude particle"):
        O.Oracle(s, O.TG)
    s = synth.water_box(10, 1)
    with pytest.raises(O.OracleError, match="Temperature group of constrained"):
        s2 = synth.water_box(10, 2); O.Oracle(s2, O.TG, constraints=np.array([[0, 4]], np.int32))
    # reference platform: more than 2 r_max beyond the wall throws (ReferenceDrudeTGNHKernels.cpp:311-312); CUDA never does (D5)
    s = synth.water_box(4, 1, use_com_temp_group=False)
    s.positions[s.pair_drude[0]] = s.positions[s.pair_parent[0]] + np.array([0.0, 0.0, 0.2])
    p, v, f = s.positions.copy(), s.velocities.copy(), np.zeros_like(s.forces)
    with pytest.raises(O.OracleError, match="too far beyond hard wall"):
        O.Oracle(s, O.REF).step(p, v, f, 1)
    O.Oracle(s, O.TG).step(s.positions.copy(), s.velocities.copy(), f, 1)


def test_threads_do_not_change_results_beyond_rounding():
    s = synth.water_box(3000, 4, quantize_masses=True)
    res = []
    for nt in (1, 4):
        O.lib().tgnh_oracle_set_threads(nt)
        o = O.Oracle(s, O.TG)
        p, v, f = s.positions.copy(), s.velocities.copy(), s.forces.copy()
        o.step(p, v, f, 3)
        res.append((p, v, o.ke2))
    O.lib().tgnh_oracle_set_threads(1)
    np.testing.assert_allclose(res[0][1], res[1][1], rtol=0, atol=1e-10)
    np.testing.assert_allclose(res[0][2], res[1][2], rtol=1e-12)


def test_fp32_state_sensitivity():
    """What an fp32 velocity / position layout can reproduce over 1000 steps, measured inside the fp64 oracle by
    rounding its own state to fp32 after every step (no hard-wall events: the wall sits at 2 nm).
    Without Drude chains every thermostat stays within 1e-7.  With a Drude Nose-Hoover CHAIN at tau = 5 fs that one
    thermostat is chaotic: its temperature moves by > 1e-4 while the others stay within 1e-7.  These are the
    bounds tests/test_gpu_parity.py::test_thousand_steps_thermostat_parity asserts."""
    out = {}
    for chain in (False, True):
        s = synth.water_box(6000, 4, quantize_masses=True, cold_drudes=True, drude_sigma=1.4e-4, force_sigma=2.0,
                            max_drude_distance=2.0, use_drude_nh_chains=chain)
        temps = []
        for rounded in (False, True):
            o = O.Oracle(s, O.TG)
            p, v, f = s.positions.copy(), s.velocities.copy(), s.forces.copy()
            for _ in range(1000):
                o.step(p, v, f, 1)
                if rounded:
                    v = v.astype(np.float32).astype(np.float64); p = p.astype(np.float32).astype(np.float64)
            temps.append(o.ke2 / (o.thermostat_params()[0] * O.BOLTZ))
        out[chain] = np.abs(temps[1] / temps[0] - 1)
    assert out[False].max() < 1e-6
    assert out[True][:-1].max() < 1e-6 and out[True][-1] > 1e-4


def test_fixed_forces_make_fp32_rounding_errors_persistent():
    """Why chain variables of a single-precision run cannot be held to 1e-6 per variable in the integrator-only tests (DESIGN.md "Parity").
    No kernel involved: numpy emulation of the half kicks v <- fl32(v + dv).  With v on the float grid and dv the SAME every step (fixed
    synthetic forces) the rounding error of a component is the same every step, so the mass-weighted radial error of a whole temperature
    group, sum m (v32 - v64).v64 / sum m |v64|^2, keeps its value step after step (~1e-9 for 12 500 molecules) instead of averaging out;
    with forces that change every step (real MD) it scatters around zero.  A scale error of 1e-9 per step moves a thermostat's chain
    velocities by 1e-6 relative (second half of the test, on the fp64 oracle)."""
    s = synth.water_box(12500, 4, quantize_masses=True, cold_drudes=True, drude_sigma=1.4e-4, force_sigma=2.0, max_drude_distance=2.0,
                        use_drude_nh_chains=False)
    f32 = lambda a: np.asarray(a, np.float64).astype(np.float32).astype(np.float64)
    m, w, dt = s.masses[:, None], 1.0 / s.masses[:, None], s.step_size
    sel = s.temp_group == 2

    def bias(a, r):
        return float(np.sum(m[sel] * (a[sel] - r[sel]) * r[sel]) / np.sum(m[sel] * r[sel] ** 2))

    def run(forces_of_step):
        v = f32(s.velocities)
        out = []
        for t in range(60):
            dv = 0.5 * dt * w * forces_of_step(t)
            exact = v + dv + dv
            v = f32(f32(v + dv) + dv)                        # two half kicks, each stored in single precision
            if t >= 50:
                out.append(bias(v, exact))
        return np.array(out)

    fixed = run(lambda t: s.forces)
    rng = np.random.default_rng(5)
    changing = run(lambda t: f32(2.0 * rng.standard_normal(s.forces.shape)))
    assert abs(fixed.mean()) > 3e-10 and fixed.std() < 0.1 * abs(fixed.mean())          # one value, step after step
    assert abs(changing.mean()) < 3 * changing.std() / np.sqrt(len(changing)) + 1e-10      # scatters around zero
    # response of the chain to such a scale error (fp64 oracle, 300 steps): d(eta_dot)/eta_dot ~ 1e3 x the error per step
    O.lib().tgnh_oracle_set_threads(os.cpu_count() or 1)
    ref = O.Oracle(s, O.TG); p, v, f = s.positions.copy(), s.velocities.copy(), s.forces.copy(); ref.step(p, v, f, 300)
    o = O.Oracle(s, O.TG); p, v, f = s.positions.copy(), s.velocities.copy(), s.forces.copy()
    for _ in range(300):
        o.step(p, v, f, 1)
        v *= 1.0 + 1e-9
    O.lib().tgnh_oracle_set_threads(1)
    rel = np.abs(o.chain_state()[1][:4, 0] / ref.chain_state()[1][:4, 0] - 1)
    assert np.all(rel > 1e-7) and np.all(rel < 1e-5), rel
