"""Diagnostic (GPU): where does the per-step radial velocity bias come from?  One tgnh_step from identical state vs the oracle."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from openmm_drudenose_b200 import capi, synth
from oracle import oracle as O
from util import DeviceState
np.set_printoptions(linewidth=220, precision=3)
dev = torch.device("cuda:0")
f32 = lambda a: a.astype(np.float32).astype(np.float64)
def run(label, **kw):
    s = synth.water_box(12500, 4, quantize_masses=True, cold_drudes=True, drude_sigma=1.4e-4, max_drude_distance=2.0, use_drude_nh_chains=False, **kw)
    m = s.masses[:, None]; tg = s.temp_group; n = s.num_particles
    res = s.res_id
    Mres = np.bincount(res, weights=s.masses)
    def parts(v):
        V = np.stack([np.bincount(res, weights=s.masses * v[:, c]) for c in range(3)], 1) / Mres[:, None]
        return v - V[res], V[res]
    def bias(vd, vr, sel):
        return float(np.sum(m[sel] * (vd[sel] - vr[sel]) * vr[sel]) / np.sum(m[sel] * vr[sel] ** 2))
    st = DeviceState(s, dev); h = capi.Handle(s); o = O.Oracle(s, O.TG)
    h.step(*st.ptrs, nsteps=50); torch.cuda.synchronize()
    rows = []
    for i in range(12):
        p, v, f = st.pos().copy(), st.vel().copy(), s.forces.copy()
        o.set_chain_state(*h.chain_state())
        h.invalidate(); h.step(*st.ptrs, nsteps=1); torch.cuda.synchronize()
        o.step(p, v, f, 1)
        vd = st.vel(); vr32 = f32(v)
        rd, Vd = parts(vd); rr, Vr = parts(v)
        g = 2
        sel = tg == g
        rows.append([bias(vd, v, sel), bias(vd, vr32, sel), bias(rd, rr, sel), bias(Vd, Vr, sel),
                     bias(vd, v, sel & (s.masses > 10)), bias(vd, v, sel & (s.masses < 0.5)), bias(vd, v, sel & (s.masses == 1.0)),
                     h.vscale()[g] - 1, o.vscale[g] - 1])
    a = np.array(rows)
    print(f"== {label}: group 2, 12 steps; columns: total | vs fp32-rounded oracle | COM-relative part | COM part | parents | Drudes | H | s-1 dev | s-1 oracle")
    print(a)
    print("mean", a.mean(0))
    h.close()
run("standard (force sigma 2)", force_sigma=2.0)
run("no forces", force_sigma=0.0)
run("weak thermostat (tau 1e3 ps)", force_sigma=2.0, coupling_time=1e3, drude_coupling_time=1e3)
