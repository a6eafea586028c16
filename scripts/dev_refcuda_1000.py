"""Diagnostic (GPU): per-variable differences after 1000 steps between the reference's CUDA platform and (a) this repo's plugin in the
single layout, (b) the plugin in the mixed layout, (c) the oracle's TG layer; and the reference against itself in mixed vs double."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openmm_drudenose_b200 import synth
from oracle import oracle as O, refcuda as R

np.set_printoptions(linewidth=200, precision=3)
for drude_chain in (False, True):
    s = synth.water_box(12500, 4, quantize_masses=True, cold_drudes=True, drude_sigma=1.4e-4, force_sigma=2.0, max_drude_distance=2.0,
                        use_drude_nh_chains=drude_chain)
    s.forces = np.rint(s.forces * 4294967296.0) / 4294967296.0
    p = s.positions.astype(np.float32).astype(np.float64); v = s.velocities.astype(np.float32).astype(np.float64)
    sims = {"ref_double": R.CudaSim(s, "reference", "double"), "ref_mixed": R.CudaSim(s, "reference", "mixed"),
            "b200_single": R.CudaSim(s, "b200", "single"), "b200_mixed": R.CudaSim(s, "b200", "mixed")}
    out = {}
    for k, sim in sims.items():
        sim.set_state(p, v, s.forces); sim.step(1000)
        out[k] = sim.thermostat()
        sim.close()
    o = O.Oracle(s, O.TG); po, vo, fo = p.copy(), v.copy(), s.forces.copy(); o.step(po, vo, fo, 1000)
    out["oracle"] = (*o.chain_state(), o.vscale)
    ref = out["ref_double"]
    print(f"==== drude_chain={drude_chain}; reference (double) eta_dot:\n{ref[1]}\nvscale {ref[3]}")
    for k in ("ref_mixed", "oracle", "b200_mixed", "b200_single"):
        eta, ed, edd, vs = out[k]
        with np.errstate(divide="ignore", invalid="ignore"):
            rel_ed = np.abs(ed - ref[1]) / np.abs(ref[1]); rel_eta = np.abs(eta - ref[0]) / np.abs(ref[0])
        print(f"-- {k}: vscale rel {np.max(np.abs(vs/ref[3]-1)):.2e}; eta_dot abs/max {np.max(np.abs(ed-ref[1]))/np.abs(ref[1]).max():.2e}; eta abs/max {np.max(np.abs(eta-ref[0]))/np.abs(ref[0]).max():.2e}")
        print("   per-variable rel eta_dot:\n", np.nan_to_num(rel_ed), "\n   per-variable rel eta:\n", np.nan_to_num(rel_eta))
