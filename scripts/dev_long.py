import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R+'/tests')
import numpy as np, torch
from openmm_drudenose_b200 import synth, capi
from oracle import oracle as O
from util import DeviceState, group_temperatures
dev=torch.device('cuda:0')
s = synth.water_box(25000, 4, quantize_masses=True, cold_drudes=True, drude_sigma=1.4e-4, force_sigma=2.0, max_drude_distance=0.5)
st = DeviceState(s, dev); h = capi.Handle(s); o = O.Oracle(s, O.TG)
p, v, f = s.positions.copy(), s.velocities.copy(), s.forces.copy()
dof = o.thermostat_params()[0]
for blk in range(12):
    n = 1 if blk < 4 else (10 if blk < 8 else 240)
    h.step(*st.ptrs, nsteps=n); o.step(p, v, f, n)
    tg = group_temperatures(h.kinetic_energies(), dof); tr = group_temperatures(o.ke2, dof)
    vg = st.vel()
    relg = vg[s.pair_parent]-vg[s.pair_drude]; relr = v[s.pair_parent]-v[s.pair_drude]
    print(blk, n, "T_D gpu %.6f ref %.6f | T0 %.6f %.6f | vscaleD %.8f %.8f | rel rms diff %.3e rel rms %.4f" % (tg[-1], tr[-1], tg[0], tr[0], h.vscale()[-1], o.vscale[-1], np.sqrt(np.mean((relg-relr)**2)), np.sqrt(np.mean(relr**2))))
