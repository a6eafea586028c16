"""Residue-per-lane reduction on/off (TGNH_RPL) for a K = 4 and a K = 5 water box of ~10 M particles: per-kind launch times."""
import os, sys
import numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from openmm_drudenose_b200 import capi, synth
from util import DeviceState
dev = torch.device("cuda:0")
which = sys.argv[1]
ffmt = 1 if which.endswith("i") else 0          # k4i: OpenMM's int64 fixed-point forces
s = synth.water_box(2_500_000, 4, pair_force="common", cold_drudes=True, force_sigma=2.0) if which.startswith("k4") else synth.swm4_box(2_000_000, pair_force="common", cold_drudes=True, force_sigma=2.0)
st = DeviceState(s, dev, force_format=ffmt)
h = capi.Handle(s, force_format=ffmt)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); sp = stream.cuda_stream
h.step(*st.ptrs, nsteps=5, stream=sp); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); h.step(*st.ptrs, nsteps=40, stream=sp); e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 40 * 1e3
h.set_profiling(True); h.step(*st.ptrs, nsteps=20, stream=sp); torch.cuda.synchronize()
print(which, "RPL env", os.environ.get("TGNH_RPL"), "rpl", h.residue_per_lane, "N", s.num_particles, "us/step %.1f" % t,
      {k: (round(v[0] / max(v[1], 1) * 1e3, 1), v[1]) for k, v in h.profile().items()}, "ke2", np.round(h.kinetic_energies(), 3)[:3])
