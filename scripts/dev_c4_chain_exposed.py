"""How much of the Nose-Hoover chain's latency is visible in the C4 step?  Step time of the 10M-particle system against S (sub-steps
per chain update): the slope is the exposed cost per sub-step pair."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np, torch
import bench
from openmm_drudenose_b200 import capi, synth
dev = torch.device("cuda:0")
res = {}
for S in [int(x) for x in os.environ.get("S_LIST", "1,20,60").split(",")]:
    s = synth.water_box(2_500_000, 4, drude_steps=S, **bench.C4_STATE["c4"])
    padded, _, (velm, posq, force) = bench.device_buffers(torch, s, dev)
    h = capi.Handle(s, padded=padded)
    ptrs = [velm.data_ptr(), posq.data_ptr(), force.data_ptr()]
    h.step(*ptrs, 5); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); h.step(*ptrs, 50); e1.record(); torch.cuda.synchronize()
    res[S] = e0.elapsed_time(e1) / 50 * 1e3
    print(f"S={S}: {res[S]:.1f} us/step  hint/ke2: {h.kinetic_energies()[-2:]}", flush=True)
    h.close(); del velm, posq, force
if 60 in res and 20 in res and 1 in res: print(f"exposed chain cost: {(res[60]-res[20])/40*1000:.0f} ns per sub-step pair; S=20 costs {(res[20]-res[1]):.1f} us over S=1")
