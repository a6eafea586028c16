"""Diagnostic (GPU): (1) error of the device's kinetic-energy sums against the fp64 oracle on the SAME fp32 velocities, per thermostat;
(2) per-step drift of the thermostat scale factors between device and oracle when both start every step from the same fp32 state."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from openmm_drudenose_b200 import capi, synth
from oracle import oracle as O
from util import DeviceState
np.set_printoptions(linewidth=200, precision=3)
dev = torch.device("cuda:0")
s = synth.water_box(12500, 4, quantize_masses=True, cold_drudes=True, drude_sigma=1.4e-4, force_sigma=2.0, max_drude_distance=2.0, use_drude_nh_chains=False)
for gen in (2, 1):
    os.environ["TGNH_V2"] = "1" if gen == 2 else "0"
    st = DeviceState(s, dev)
    h = capi.Handle(s)
    o = O.Oracle(s, O.TG)
    errs = []
    for i in range(10):
        h.step(*st.ptrs, nsteps=7)
        torch.cuda.synchronize()
        ke_d = h.compute_kinetic_energies(st.velm.data_ptr())
        ke_o = o.compute_ke2(np.ascontiguousarray(st.vel()))
        errs.append(ke_d / ke_o - 1)
    print(f"generation {h.kernel_generation}: KE(device)/KE(oracle on the same fp32 velocities) - 1, 10 states:\n", np.array(errs))
    # per-step: scale factors from identical state
    p, v, f = st.pos().copy(), st.vel().copy(), s.forces.copy()
    o2 = O.Oracle(s, O.TG); o2.set_chain_state(*h.chain_state())
    n = s.num_particles
    d = []
    for i in range(10):
        h.invalidate()
        h.step(*st.ptrs, nsteps=1)
        o2.step(p, v, f, 1)
        d.append(h.vscale() / o2.vscale - 1)
        # resync oracle to the device state
        p, v = st.pos().copy(), st.vel().copy(); o2.set_chain_state(*h.chain_state())
    print(" vscale(device)/vscale(oracle) - 1 per step from identical state:\n", np.array(d))
    h.close()
