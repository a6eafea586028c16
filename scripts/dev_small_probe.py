"""Small systems: per-step time of tgnh_step and per-kind launch times (profiling mode) for C1/C2/C3, both kernel
generations where both apply (TGNH_V2=0 forces the first generation)."""
import os, sys
import torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from openmm_drudenose_b200 import capi, synth
from util import DeviceState
dev = torch.device("cuda:0")
which = sys.argv[1]
s = {"c1": synth.nacl_box, "c2": lambda: synth.swm4_box(10000), "c3": lambda: synth.ionic_liquid(1000)}[which]()
st = DeviceState(s, dev)
h = capi.Handle(s)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); sp = stream.cuda_stream
def timed(n=200):
    h.step(*st.ptrs, nsteps=20, stream=sp); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); h.step(*st.ptrs, nsteps=n, stream=sp); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
t = [round(timed(), 2) for _ in range(3)]
h.set_profiling(True); h.step(*st.ptrs, nsteps=50, stream=sp); torch.cuda.synchronize()
print(which, "V2=%s FUSE=%s" % (os.environ.get("TGNH_V2"), os.environ.get("TGNH_FUSE_CHAIN")), "gen", h.kernel_generation, "N", s.num_particles,
      "T", len(h.vscale()), "us/step", t, {k: (round(v[0] / max(v[1], 1) * 1e3, 2), v[1]) for k, v in h.profile().items()})
