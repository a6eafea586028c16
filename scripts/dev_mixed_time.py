"""C4 in the OpenMM mixed / double layouts (double4 velm, int64 forces): ms per step and algorithmic bandwidth
(240 B per particle-step for mixed = 2*(32+24+32) + (32+32))."""
import sys
import numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from openmm_drudenose_b200 import capi, synth
from util import DeviceState
dev = torch.device("cuda:0")
s = synth.water_box(2_500_000, 4)
for prec, name, alg in ((0, "single + int64 forces", 144), (1, "mixed", 240), (2, "double", 240)):
    st = DeviceState(s, dev, force_format=capi.FORCE_I64_SOA, precision=prec)
    h = capi.Handle(s, force_format=capi.FORCE_I64_SOA, precision=prec, padded=st.padded)
    if prec == 1:
        h.set_posq_correction(st.corr.data_ptr())
    h.step(*st.ptrs, nsteps=5); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); h.step(*st.ptrs, nsteps=30); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    print(f"{name}: {ms:.4f} ms/step, {s.num_particles / ms / 1e6:.2f}e9 particle-steps/s, {alg * s.num_particles / ms / 1e6:.0f} GB/s algorithmic ({alg} B)")
    h.close(); del st
    torch.cuda.empty_cache()
