import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R+'/tests')
import numpy as np, torch
from openmm_drudenose_b200 import synth, capi
from oracle import oracle as O
from util import DeviceState, group_temperatures
dev=torch.device('cuda:0')
np.set_printoptions(linewidth=220, precision=3)
s = synth.water_box(25000, 4, quantize_masses=True, cold_drudes=True, drude_sigma=1.4e-4, force_sigma=2.0, max_drude_distance=2.0, use_drude_nh_chains=False)
st = DeviceState(s, dev); h = capi.Handle(s); o = O.Oracle(s, O.TG)
p, v, f = s.positions.copy(), s.velocities.copy(), s.forces.copy()
tot=0
for n in (1, 1, 8, 90, 400, 500):
    h.step(*st.ptrs, nsteps=n); o.step(p, v, f, n); tot+=n
    eg=h.chain_state()[1]; er=o.chain_state()[1]
    print(tot, "ke rel", (h.kinetic_energies()-o.ke2)/o.ke2)
    print("    etadot rel err rows0,4,5:", ((eg-er)/np.where(er!=0,er,1))[[0,4,5],:3].ravel())
