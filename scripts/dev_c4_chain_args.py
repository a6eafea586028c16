"""Largest exp argument dtc * |etaDot| in the chain for the C4 generator (1M-particle sample) along the bench's step range."""
import sys
import numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from openmm_drudenose_b200 import capi, synth
from util import DeviceState
dev = torch.device("cuda:0")
s = synth.water_box(250000, 4, box_molecules=2500000)
st = DeviceState(s, dev)
h = capi.Handle(s)
dtc = s.step_size / s.drude_steps
done = 0
for n in (1, 2, 7, 40, 50, 100, 300, 500):
    h.step(*st.ptrs, nsteps=n); done += n
    ed = h.chain_state()[1]
    dof = h.thermostat_params()[0]
    print(done, "max dtc*|etaDot| per thermostat", np.array2string(dtc * np.abs(ed).max(axis=1), precision=3), "T", np.array2string(h.kinetic_energies() / np.maximum(dof, 1) / 0.0083144626, precision=3))
