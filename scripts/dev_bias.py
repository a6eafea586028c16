"""Diagnostic (GPU): systematic (radial) velocity error of one step from identical state, per temperature group:
delta_g = sum m (v_dev - v_ref) . v_ref / sum m |v_ref|^2 over the particles of group g (a scale bias of 1e-9 per step moves the
chain velocities by 1e-6 relative, scripts notes in DESIGN.md)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from openmm_drudenose_b200 import capi, synth
from oracle import oracle as O
from util import DeviceState
np.set_printoptions(linewidth=200, precision=3)
dev = torch.device("cuda:0")
s = synth.water_box(12500, 4, quantize_masses=True, cold_drudes=True, drude_sigma=1.4e-4, force_sigma=2.0, max_drude_distance=2.0, use_drude_nh_chains=False)
m = s.masses[:, None]
def bias(vd, vr):
    out = []
    for g in range(4):
        sel = s.temp_group == g
        out.append(float(np.sum(m[sel] * (vd[sel] - vr[sel]) * vr[sel]) / np.sum(m[sel] * vr[sel] ** 2)))
    return np.array(out)
for gen in (2, 1):
    os.environ["TGNH_V2"] = "1" if gen == 2 else "0"
    st = DeviceState(s, dev); h = capi.Handle(s); o = O.Oracle(s, O.TG)
    h.step(*st.ptrs, nsteps=50); torch.cuda.synchronize()
    n = s.num_particles
    res = {k: [] for k in ("after half1", "after half2 (kick+scale)", "full step via tgnh_step")}
    for i in range(8):
        p, v, f = st.pos().copy(), st.vel().copy(), s.forces.copy()
        o.set_chain_state(*h.chain_state())
        # first half on both sides
        h.invalidate(); h.half1(*st.ptrs); torch.cuda.synchronize()
        o.propagate_nh_chain(v); o.half_kick(v, f); o.drift(p, v); o.hard_wall(p, v)
        res["after half1"].append(bias(st.vel(), v))
        # resync, second half
        v = st.vel().copy(); o.set_chain_state(*h.chain_state())
        h.half2(st.velm.data_ptr(), st.force.data_ptr()); torch.cuda.synchronize()
        o.half_kick(v, f); o.propagate_nh_chain(v)
        res["after half2 (kick+scale)"].append(bias(st.vel(), v))
        p, v = st.pos().copy(), st.vel().copy(); o.set_chain_state(*h.chain_state())
        h.invalidate(); h.step(*st.ptrs, nsteps=1); torch.cuda.synchronize()
        o.step(p, v, f, 1)
        res["full step via tgnh_step"].append(bias(st.vel(), v))
    print(f"generation {h.kernel_generation}")
    for k, a in res.items():
        a = np.array(a); print(f"  {k}: mean per group {a.mean(0)}  (std {a.std(0)})")
    h.close()
