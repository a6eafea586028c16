"""The oracle against the reference's OWN code: /root/reference's openmmapi + reference-platform + serialization sources,
compiled unmodified against the OpenMM API shim into oracle/_ref/libref.so (oracle/Makefile).  No GPU.
Skipped only where that library could not be built (no /root/reference at build time and no shipped .so)."""
import numpy as np
import pytest

from openmm_drudenose_b200 import synth
from oracle import oracle as O
from oracle import ref as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libref.so not built (reference tree absent)")

CASES = {
    "overlap_M3_drude_chain": (dict(use_com_temp_group=False), 0),
    "harmonic_M2": (dict(use_com_temp_group=False, num_nh_chains=2, pair_force="none", cold_drudes=True, drude_sigma=1.4e-4, force_sigma=0.0), 1),
    "default_no_drude_chain_D1_layout": (dict(use_com_temp_group=False, use_drude_nh_chains=False), 0),
    "M1": (dict(use_com_temp_group=False, num_nh_chains=1, use_drude_nh_chains=False), 0),
    "hard_wall_hits": (dict(use_com_temp_group=False, pair_force="frozen_spring", max_drude_distance=0.02), 0),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_ref_layer_is_bit_identical_to_the_reference_platform(name):
    """ReferenceIntegrateDrudeTGNHStepKernel::execute (platforms/reference/src/ReferenceDrudeTGNHKernels.cpp:221-415) run
    through the reference's own DrudeTGNHIntegrator::step vs tgnh_oracle.c's TGNH_ORACLE_REF layer: identical bits."""
    kw, model = CASES[name]
    s = synth.water_box(300, 1, **kw)
    sim = R.ReferenceSim(s, force_model=model)
    o = O.Oracle(s, O.REF)
    pa, va = s.positions.copy(), s.velocities.copy()
    pb, vb = pa.copy(), va.copy()
    fa = O.harmonic_forces(s, pa) if model else s.forces.copy()
    fb = fa.copy()
    sim.step(pa, va, fa, 40)
    o.step(pb, vb, fb, 40, model, None, s.k_spring if model else None)
    assert np.array_equal(va, vb) and np.array_equal(pa, pb)
    assert sim.num_residues == s.num_residues          # DrudeTGNHIntegrator::initialize residue tables (:136-153)


def test_tg_layer_matches_the_reference_platform_on_the_overlap_domain():
    """G = 1, no COM group, Drude chains: the CUDA-platform restatement (TG layer) reproduces the REAL reference platform to
    rounding (the two platforms order a few operations differently)."""
    s = synth.water_box(400, 1, use_com_temp_group=False, quantize_masses=True)
    sim = R.ReferenceSim(s)
    o = O.Oracle(s, O.TG)
    pa, va, fa = s.positions.copy(), s.velocities.copy(), s.forces.copy()
    pb, vb, fb = pa.copy(), va.copy(), fa.copy()
    sim.step(pa, va, fa, 100)
    o.step(pb, vb, fb, 100)
    np.testing.assert_allclose(vb, va, rtol=0, atol=1e-11)
    np.testing.assert_allclose(pb, pa, rtol=0, atol=1e-12)


def test_reference_platform_ignores_temperature_groups():
    """SURVEY.md finding 1: platforms/reference never reads the temperature-group tables."""
    s = synth.water_box(200, 3, use_com_temp_group=False)
    a = R.ReferenceSim(s)
    one = synth.water_box(200, 1, use_com_temp_group=False)
    b = R.ReferenceSim(one)
    pa, va, fa = s.positions.copy(), s.velocities.copy(), s.forces.copy()
    pb, vb, fb = pa.copy(), va.copy(), fa.copy()
    a.step(pa, va, fa, 10); b.step(pb, vb, fb, 10)
    assert np.array_equal(va, vb)


def test_reference_integrator_validation():
    """DrudeTGNHIntegrator's own checks (openmmapi/src/DrudeTGNHIntegrator.cpp:66-134)."""
    s = synth.water_box(10, 2)
    s.num_temp_groups = 1                                 # particle groups reference group 1, only one exists
    with pytest.raises(R.RefError, match="Index out of range"):
        R.ReferenceSim(s)
    s = synth.water_box(10, 1)
    s.max_drude_distance = -1.0
    with pytest.raises(R.RefError, match="Distance cannot be negative"):
        R.ReferenceSim(s)
    s = synth.water_box(5, 1)
    s.pair_drude = s.pair_drude[:0]; s.pair_parent = s.pair_parent[:0]; s.k_spring = s.k_spring[:0]
    R.ReferenceSim.__init__  # (a System without pairs is legal for the integrator: one empty DrudeForce is still present)


def test_reference_serialization_roundtrip():
    """testSerialization (serialization/tests/TestSerializeDrudeTGNHIntegrator.cpp:45-67) through the reference's proxy."""
    out, xml = R.serialization_roundtrip(301.1, 0.1, 10.5, 0.005, 0.001)
    np.testing.assert_array_equal(out, [301.1, 0.1, 10.5, 0.005, 0.001, 20, 1, 0, 1e-5])
    assert 'type="DrudeTGNHIntegrator"' in xml and 'version="1"' in xml
    for prop in ("stepSize", "constraintTolerance", "temperature", "couplingTime", "drudeTemperature", "drudeCouplingTime",
                 "drudeStepsPerRealStep", "numNHChains", "useDrudeNHChains"):
        assert prop + '="' in xml
    out, _ = R.serialization_roundtrip(300.0, 0.2, 1.0, 0.01, 0.002, 10, 4, True, 1e-6)
    np.testing.assert_array_equal(out, [300.0, 0.2, 1.0, 0.01, 0.002, 10, 4, 1, 1e-6])
