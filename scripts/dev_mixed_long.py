"""Growth of the mixed-precision path's deviation from the fp64 oracle over a free run (rounding-level differences
amplified by the Drude chain's own dynamics)."""
import sys
import numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from openmm_drudenose_b200 import capi, synth
from oracle import oracle as O
from util import DeviceState, group_temperatures

cuda = torch.device("cuda:0")
s = synth.water_box(25000, 4, cold_drudes=True, drude_sigma=1.4e-4, force_sigma=2.0, max_drude_distance=2.0)
s.forces = s.forces.astype(np.float32).astype(np.float64)
st = DeviceState(s, cuda, precision=1)
h = capi.Handle(s, precision=capi.PRECISION_MIXED, padded=st.padded); h.set_posq_correction(st.corr.data_ptr())
o = O.Oracle(s, O.TG)
p, v, f = s.positions.copy(), s.velocities.copy(), s.forces.copy()
dof = o.thermostat_params()[0]
for blk in range(8):
    h.step(*st.ptrs, nsteps=125); o.step(p, v, f, 125)
    tg, tr = group_temperatures(h.kinetic_energies(), dof), group_temperatures(o.ke2, dof)
    eg, er = h.chain_state(), o.chain_state()
    rel = [np.abs(a - b) / np.maximum(np.abs(b), 1e-300) for a, b in zip(eg[:2], er[:2])]
    print((blk + 1) * 125, "T rel", np.array2string(np.abs(tg - tr) / tr, precision=1), "eta rel max/thermostat", np.array2string(rel[0].max(axis=1), precision=1),
          "etaDot", np.array2string(rel[1].max(axis=1), precision=1), flush=True)
