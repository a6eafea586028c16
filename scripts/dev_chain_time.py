"""Per-sub-step cost of the device Nose-Hoover chain: step time of a tiny system against S (drude_steps) and M."""
import sys, time
import numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from openmm_drudenose_b200 import capi, synth
from util import DeviceState

dev = torch.device("cuda:0")
for M in (1, 3):
    for dr in (True, False):
        res = []
        for S in (1, 20, 80):
            s = synth.water_box(64, 2, drude_steps=S, num_nh_chains=M, use_drude_nh_chains=dr)
            st = DeviceState(s, dev)
            h = capi.Handle(s)
            h.step(*st.ptrs, nsteps=20)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); h.step(*st.ptrs, nsteps=400); e1.record(); torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / 400 * 1e3)
            h.close()
        print(f"M={M} drude_chain={dr}: us/step S=1 {res[0]:.2f}  S=20 {res[1]:.2f}  S=80 {res[2]:.2f}  -> per sub-step pair {(res[2]-res[1])/60*1000:.0f} ns = {(res[2]-res[1])/60*1965/2:.0f} cycles per sub-step")
